"""The reference-side binding is committed as SOURCE (no JDK in the image: it cannot be compiled here).  This test keeps it
honest against the C ABI: every entry point include/ngsep_pileup.h exports has exactly one JNI function in
jni/ngsep_pileup_jni.c that calls it and one native method of the same name in NgsepPileupNative.java, with matching
parameter counts; the struct offsets GpuAlignmentsPileupGenerator.toCalledSNV reads are the ones of ngsep_snv_call; the Java
generator overrides the public entry points of AlignmentsPileupGenerator it must intercept."""
import os
import re

import numpy as np

from ngsepcore_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(*p):
    return open(os.path.join(ROOT, *p)).read()


def _camel(c_name):          # ngsep_tile_fetch_pop -> tileFetchPop
    parts = c_name[len("ngsep_"):].split("_")
    return parts[0] + "".join(x.capitalize() for x in parts[1:])


def test_every_export_has_a_jni_function_and_a_native_method():
    header = _read("include", "ngsep_pileup.h")
    exports = re.findall(r"^(?:int|const char\*)\s+(ngsep_\w+)\s*\(", header, re.M)
    assert sorted(exports) == sorted(capi.EXPORTS)
    shim = _read("jni", "ngsep_pileup_jni.c")
    java = _read("java", "ngsep", "discovery", "gpu", "NgsepPileupNative.java")
    jni_fns = dict((m.group(1), m.group(2)) for m in re.finditer(r"JNI_FN\((\w+)\)\(([^)]*)\)", shim))
    natives = dict((m.group(1), m.group(2)) for m in re.finditer(r"native\s+[\w\[\]]+\s+(\w+)\(([^)]*)\)", java))
    for c in exports:
        name = _camel(c)
        assert name in jni_fns, f"no JNI function for {c}"
        assert name in natives, f"no native method for {c}"
        assert re.search(r"\b%s\s*\(" % c, shim), f"the shim never calls {c}"
        n_jni = len([a for a in jni_fns[name].split(",") if a.strip()]) - 2          # JNIEnv*, jclass
        n_java = len([a for a in natives[name].split(",") if a.strip()])
        assert n_jni == n_java, (c, n_jni, n_java)
    assert set(jni_fns) == set(natives) == {_camel(c) for c in exports}


def test_java_reads_the_call_record_at_the_c_offsets():
    gen = _read("java", "ngsep", "discovery", "gpu", "GpuAlignmentsPileupGenerator.java")
    off = {n: capi.CALL_DTYPE.fields[n][1] for n in capi.CALL_DTYPE.names}
    assert (off["pos"], off["count"], off["depth"], off["logcond"], off["gq"], off["qual"], off["a0"], off["a1"], off["ref"], off["kind"]) == \
           (4, 8, 24, 64, 208, 210, 212, 213, 214, 215)
    for expr in ("b.getInt(o + 4)", "b.getInt(o + 8)", "b.getInt(o + 24)", "b.getDouble(o + 64 + 8 * (4 * i + j))", "b.getShort(o + 208)",
                 "b.getShort(o + 210)", "b.get(o + 212)", "b.get(o + 213)", "b.get(o + 214)", "b.get(o + 215)"):
        assert expr in gen, expr
    native = _read("java", "ngsep", "discovery", "gpu", "NgsepPileupNative.java")
    assert f"CALL_BYTES = {capi.CALL_DTYPE.itemsize}" in native and f"SITE_BYTES = {capi.SITE_DTYPE.itemsize}" in native
    assert f"SAMPLE_CALL_BYTES = {capi.SAMPLE_CALL_DTYPE.itemsize}" in native and f"HDR_BYTES = {capi.HDR_DTYPE.itemsize}" in native
    for flag, val in (("M_NEG", capi.M_NEG), ("M_NMASK", capi.M_NMASK), ("M_WINDOW", capi.M_WINDOW), ("M_SIMPLE", capi.M_SIMPLE)):
        assert f"{flag} = 0x{val:08x}" in native


def test_generator_mirrors_the_reference_surface():
    gen = _read("java", "ngsep", "discovery", "gpu", "GpuAlignmentsPileupGenerator.java")
    assert "extends AlignmentsPileupGenerator" in gen
    for method in ("public void addListener(PileupListener listener)", "public void processAlignment(ReadAlignment aln)", "public void notifyEndOfAlignments()"):
        assert method in gen, method
    # everything else of APG:71-193, 201-361 is inherited: the setters feed the getters this class reads
    for getter in ("getMaxAlnsPerStartPos()", "getBasesToIgnore5P()", "getBasesToIgnore3P()", "getQuerySeq()", "getQueryFirst()", "getQueryLast()", "getGenome()",
                   "getSequencesMetadata()"):
        assert getter in gen, getter
