"""ctypes binding of libngsep_host.so (include/ngsep_host.h): the host pipeline around the device C ABI.

``Detector`` mirrors the public surface of the reference's SingleSampleVariantsDetector for this path
(reference: src/ngsep/discovery/SingleSampleVariantsDetector.java — option setters :141-560, run() :589-656):
same option names, same defaults, VCF text out.  There is no CPU fallback: without the built libraries or a B200
every call raises NgsepError.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .capi import HDR_DTYPE, NgsepError
from .simulate import AlignmentBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libngsep_host.so")
EXPORTS = ["ngsep_detector_options_default", "ngsep_ssvd_create", "ngsep_msvd_create", "ngsep_ssvd_destroy", "ngsep_ssvd_error", "ngsep_ssvd_set_genome",
           "ngsep_ssvd_run", "ngsep_ssvd_vcf", "ngsep_ssvd_stats", "ngsep_ssvd_pack", "ngsep_host_pack", "ngsep_host_free",
           "ngsep_host_merge_order", "ngsep_host_pack_quality", "ngsep_coverage_stats", "ngsep_relative_allele_counts", "ngsep_relative_allele_counts_regions", "ngsep_host_pack_layout"]


class HostBatch(C.Structure):
    _fields_ = [("n", C.c_int64), ("seq_idx", C.c_void_p), ("first", C.c_void_p), ("flags", C.c_void_p), ("mapq", C.c_void_p),
                ("rg", C.c_void_p), ("nh", C.c_void_p), ("cigar_off", C.c_void_p), ("cigar", C.c_void_p), ("base_off", C.c_void_p),
                ("bases", C.c_void_p), ("quals", C.c_void_p), ("name_hash", C.c_void_p)]


class DetectorOptions(C.Structure):
    _fields_ = [("sample_id", C.c_char_p), ("ploidy", C.c_int32), ("print_sample_ploidy", C.c_int32), ("min_mq", C.c_int32),
                ("query_seq", C.c_char_p), ("query_first", C.c_int32), ("query_last", C.c_int32), ("ignore_lower_case_ref", C.c_int32),
                ("max_alns_per_start_pos", C.c_int32), ("process_non_unique_primary", C.c_int32), ("process_secondary", C.c_int32),
                ("ignore5", C.c_int32), ("ignore3", C.c_int32), ("het_rate", C.c_double), ("het_rate_set", C.c_int32),
                ("max_base_qs", C.c_int32), ("min_quality", C.c_int32), ("call_embedded_snvs", C.c_int32), ("calc_strand_bias", C.c_int32),
                ("device", C.c_int32), ("tile_len", C.c_int32), ("shard", C.c_int32), ("n_shards", C.c_int32), ("host_only", C.c_int32),
                ("synchronous", C.c_int32), ("n_devices", C.c_int32), ("known_strs", C.c_char_p), ("known_variants", C.c_char_p)]


class RunStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("records", "admitted", "versions", "events", "indel_calls", "snv_calls", "covered_positions",
                                          "candidate_positions", "aligned_bases", "tiles", "kernel_launches")] + \
               [(n, C.c_double) for n in ("kernel_ms", "prepass_s", "pack_s", "device_s", "total_s")]


_lib = None


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NgsepError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(LIB_PATH)
    lib.ngsep_detector_options_default.argtypes = [C.POINTER(DetectorOptions)]
    lib.ngsep_ssvd_create.argtypes = [C.POINTER(DetectorOptions), C.POINTER(C.c_void_p)]
    lib.ngsep_msvd_create.argtypes = [C.POINTER(DetectorOptions), C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_double,
                                      C.POINTER(C.c_void_p)]
    lib.ngsep_ssvd_destroy.argtypes = [C.c_void_p]
    lib.ngsep_ssvd_error.argtypes = [C.c_void_p]
    lib.ngsep_ssvd_error.restype = C.c_char_p
    lib.ngsep_ssvd_set_genome.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(C.c_void_p)]
    lib.ngsep_ssvd_run.argtypes = [C.c_void_p, C.POINTER(HostBatch), C.c_int]
    lib.ngsep_ssvd_vcf.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
    lib.ngsep_ssvd_vcf.restype = C.c_void_p
    lib.ngsep_ssvd_stats.argtypes = [C.c_void_p, C.POINTER(RunStats)]
    lib.ngsep_host_pack.argtypes = [C.POINTER(HostBatch), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                    C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
    lib.ngsep_ssvd_pack.argtypes = [C.c_void_p, C.POINTER(HostBatch), C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_int64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                    C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
    lib.ngsep_host_free.argtypes = [C.c_void_p]
    lib.ngsep_host_pack_quality.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)]
    lib.ngsep_relative_allele_counts.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(HostBatch), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                 C.POINTER(C.c_void_p)]
    lib.ngsep_relative_allele_counts_regions.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(HostBatch), C.c_int, C.c_int, C.c_int, C.c_int,
                                                         C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.ngsep_coverage_stats.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(HostBatch), C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.ngsep_host_merge_order.argtypes = [C.c_int, C.POINTER(HostBatch), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                           C.POINTER(C.c_int64)]
    _lib = lib
    return lib


def _host_batch(b: AlignmentBatch):
    keep = [np.ascontiguousarray(b.seq_idx, np.int32), np.ascontiguousarray(b.first, np.int32), np.ascontiguousarray(b.flags, np.uint16),
            np.ascontiguousarray(b.mapq, np.uint8), np.ascontiguousarray(b.rg, np.int16), np.ascontiguousarray(b.nh, np.int16),
            np.ascontiguousarray(b.cigar_off, np.int64), np.ascontiguousarray(b.cigar, np.uint32), np.ascontiguousarray(b.base_off, np.int64),
            np.ascontiguousarray(b.bases, np.uint8), np.ascontiguousarray(b.quals, np.uint8)]
    hb = HostBatch(b.n, *[a.ctypes.data for a in keep], None)
    return hb, keep


def last_pack_quality():
    """(bits, dict16) of the quality array the last pack call of this thread produced (ngsep_tile.qual_bits / qual_dict)."""
    bits = C.c_uint32(0)
    d = (C.c_uint8 * 16)()
    load_library().ngsep_host_pack_quality(C.byref(bits), d)
    return int(bits.value), bytes(d)


def last_pack_layout():
    """(layout, n_words_full) of the payload the last pack call of this thread produced (ngsep_tile.layout / n_words)."""
    layout, full = C.c_uint32(0), C.c_uint64(0)
    load_library().ngsep_host_pack_layout(C.byref(layout), C.byref(full))
    return int(layout.value), int(full.value)


def decode_quality(qual: np.ndarray, hdr: np.ndarray, payload: np.ndarray, quality) -> np.ndarray:
    """The phred+33 bytes a packed quality array stands for, in the layout of the 8-bit mode (L bytes per read padded to 4 with
    '!'; reads in order of their first record) — for tests that compare packers."""
    bits, qdict = quality
    if bits == 8:
        return qual
    lut = np.frombuffer(qdict, np.uint8)
    out = []
    seen = set()
    nxt = np.concatenate([hdr["off"][1:], [payload.size]]).astype(np.int64)
    per = 8 // bits
    for i in range(hdr.shape[0]):
        qoff = int(payload[nxt[i] - 1])
        if qoff in seen:
            continue
        seen.add(qoff)
        L = int(hdr["meta"][i] & 0xFFFF)
        raw = qual[qoff * 4: qoff * 4 + (L * bits + 7) // 8]
        codes = np.stack([(raw >> (bits * k)) & ((1 << bits) - 1) for k in range(per)], axis=1).reshape(-1)[:L]
        run = lut[codes]
        out.append(np.concatenate([run, np.full((-L) % 4, ord("!"), np.uint8)]))
    return np.concatenate(out) if out else np.zeros(0, np.uint8)


def _take(ptr, n, dtype, lib):
    """Copies n items of a malloc'ed native buffer into a numpy array and frees the buffer."""
    if n:
        nbytes = n * np.dtype(dtype).itemsize
        arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(nbytes,)).view(dtype).copy()
    else:
        arr = np.zeros(0, dtype)
    lib.ngsep_host_free(ptr)
    return arr


def pack_sequence(batch: AlignmentBatch, seq_idx: int, min_mq: int = 20, max_alns_per_start: int = 5, ignore5: int = 0, ignore3: int = 0):
    """C++ admission + packing of one sequence (no event pre-pass) -> (hdr, payload, qual, max_span, aligned_bases)."""
    lib = load_library()
    hb, keep = _host_batch(batch)
    hdr_p, pay_p, q_p = C.c_void_p(), C.c_void_p(), C.c_void_p()
    n_reads, n_words, n_qual, span, bases = C.c_int64(), C.c_uint64(), C.c_uint64(), C.c_int(), C.c_uint64()
    rc = lib.ngsep_host_pack(C.byref(hb), seq_idx, min_mq, max_alns_per_start, ignore5, ignore3, C.byref(hdr_p), C.byref(pay_p), C.byref(q_p),
                             C.byref(n_reads), C.byref(n_words), C.byref(n_qual), C.byref(span), C.byref(bases))
    if rc != 0:
        raise NgsepError(f"ngsep_host_pack failed ({rc})")
    pack_sequence.quality = last_pack_quality()      # how the returned quality array is coded (see decode_quality)
    return (_take(hdr_p, n_reads.value, HDR_DTYPE, lib), _take(pay_p, n_words.value, np.uint32, lib), _take(q_p, n_qual.value, np.uint8, lib),
            span.value, bases.value)


def merge_batches(batches, minMQ: int = 20, p: bool = False, s: bool = False) -> AlignmentBatch:
    """AlignmentsPileupGenerator.processFiles (APG:212-289) over one AlignmentBatch per input file: ONE batch with the records in
    the order the reference consumes them (per-input reader filter, k-way merge by sequence / first / last, ties to the lowest
    input).  Read groups are unified by name; the contig lists must agree."""
    lib = load_library()
    n = len(batches)
    arr = (HostBatch * n)()
    keep = []
    for i, b in enumerate(batches):
        hb, k = _host_batch(b)
        arr[i] = hb
        keep.append(k)
    in_p, rec_p, n_out = C.c_void_p(), C.c_void_p(), C.c_int64()
    rc = lib.ngsep_host_merge_order(n, arr, minMQ, int(p), int(s), C.byref(in_p), C.byref(rec_p), C.byref(n_out))
    if rc != 0:
        raise NgsepError(f"ngsep_host_merge_order failed ({rc})")
    which = _take(in_p, n_out.value, np.int32, lib)
    rec = _take(rec_p, n_out.value, np.int64, lib)
    names, samples, index = [], [], {}
    remap = []
    for b in batches:
        m = np.full(len(b.rg_names) + 1, -1, np.int16)       # last entry: "no read group" (-1) stays -1
        for j, (rid, sm) in enumerate(zip(b.rg_names, list(b.rg_samples) + [None] * len(b.rg_names))):
            if rid not in index:
                index[rid] = len(names); names.append(rid); samples.append(sm if sm is not None else rid)
            m[j] = index[rid]
        remap.append(m)
    parts = [b.select(rec[which == i]) for i, b in enumerate(batches)]
    pos_in_part = np.zeros(len(which), np.int64)
    for i in range(n):
        sel = which == i
        pos_in_part[sel] = np.arange(int(sel.sum()))
    # concatenate the parts, then gather in consumption order
    off = np.concatenate([[0], np.cumsum([pt.n for pt in parts])])
    cat_idx = off[which] + pos_in_part
    cl = np.concatenate([np.diff(pt.cigar_off) for pt in parts]); bl = np.concatenate([np.diff(pt.base_off) for pt in parts])
    c_start = np.concatenate([[0], np.cumsum(cl)])[:-1]; b_start = np.concatenate([[0], np.cumsum(bl)])[:-1]
    from .simulate import _gather_ranges
    cigar = np.concatenate([pt.cigar for pt in parts]); bases = np.concatenate([pt.bases for pt in parts]); quals = np.concatenate([pt.quals for pt in parts])
    cl_o, bl_o = cl[cat_idx], bl[cat_idx]
    co = np.zeros(len(cat_idx) + 1, np.int64); np.cumsum(cl_o, out=co[1:])
    bo = np.zeros(len(cat_idx) + 1, np.int64); np.cumsum(bl_o, out=bo[1:])
    rg = np.concatenate([remap[i][pt.rg] for i, pt in enumerate(parts)])
    cat = lambda name: np.concatenate([getattr(pt, name) for pt in parts])[cat_idx]
    b0 = batches[0]
    return AlignmentBatch(b0.contig_names, b0.contig_lengths, cat("seq_idx"), cat("first"), cat("flags"), cat("mapq"), rg[cat_idx].astype(np.int16),
                          cat("nh"), co, _gather_ranges(cigar, c_start[cat_idx], cl_o), bo, _gather_ranges(bases, b_start[cat_idx], bl_o),
                          _gather_ranges(quals, b_start[cat_idx], bl_o), names, samples)


def relative_allele_counts(batch: AlignmentBatch, m: int = 10, M: int = 1000, q: int = 20, s: bool = False, device: int = 0, frs: str | None = None,
                           srs: str | None = None) -> str:
    """GPU drop-in for the RelativeAlleleCountsCalculator module (reference: src/ngsep/discovery/RelativeAlleleCountsCalculator.java; keyword
    names = its option ids: -m minRD, -M maxRD, -q minBaseQualityScore, -s secondary alignments, -frs / -srs files with the regions to
    filter / to select): the report text."""
    lib = load_library()
    hb, keep = _host_batch(batch)
    n = len(batch.contig_names)
    names = (C.c_char_p * n)(*[x.encode() for x in batch.contig_names])
    lens = (C.c_int64 * n)(*[int(l) for l in batch.contig_lengths])
    text = C.c_void_p()
    rc = lib.ngsep_relative_allele_counts_regions(n, names, lens, C.byref(hb), m, M, q, int(s), frs.encode() if frs else None, srs.encode() if srs else None, device,
                                                  C.byref(text))
    if rc != 0:
        raise NgsepError(f"ngsep_relative_allele_counts failed ({rc})")
    out = C.string_at(text).decode()
    lib.ngsep_host_free(text)
    return out


def coverage_stats(batch: AlignmentBatch, minMQ: int = 20, device: int = 0) -> str:
    """GPU drop-in for the CoverageStats module (reference: src/ngsep/discovery/CoverageStatisticsCalculator.java): the table
    `depth<TAB>positions<TAB>positions by unique alignments` for depths 1..299 and a final `More` line."""
    lib = load_library()
    hb, keep = _host_batch(batch)
    n = len(batch.contig_names)
    names = (C.c_char_p * n)(*[s.encode() for s in batch.contig_names])
    lens = (C.c_int64 * n)(*[int(l) for l in batch.contig_lengths])
    text = C.c_void_p()
    rc = lib.ngsep_coverage_stats(n, names, lens, C.byref(hb), minMQ, device, C.byref(text))
    if rc != 0:
        raise NgsepError(f"ngsep_coverage_stats failed ({rc})")
    out = C.string_at(text).decode()
    lib.ngsep_host_free(text)
    return out


class Detector:
    """GPU drop-in for SingleSampleVariantsDetector (discovery mode).  Keyword names are the reference's option ids."""

    def __init__(self, sampleId: str = "Sample", ploidy: int = 2, psp: bool = False, minMQ: int = 20, querySeq: str | None = None,
                 first: int = 0, last: int = 1000000000, ignoreLowerCaseRef: bool = False, maxAlnsPerStartPos: int = 5, p: bool = False,
                 s: bool = False, ignore5: int = 0, ignore3: int = 0, h: float | None = None, maxBaseQS: int = 30, minQuality: int = 40,
                 embeddedSNVs: bool = False, csb: bool = False, device: int = 0, tile_len: int = 0, shard: int = 0, n_shards: int = 1, host_only: bool = False,
                 synchronous: bool = False, n_devices: int = 1, knownSTRs: str | None = None, knownVariants: str | None = None):
        self.lib = load_library()
        o = DetectorOptions()
        self.lib.ngsep_detector_options_default(C.byref(o))
        self._sid = sampleId.encode()
        self._qs = querySeq.encode() if querySeq else None
        o.sample_id, o.ploidy, o.print_sample_ploidy, o.min_mq = self._sid, ploidy, int(psp), minMQ
        o.query_seq, o.query_first, o.query_last = self._qs, first, last
        o.ignore_lower_case_ref, o.max_alns_per_start_pos = int(ignoreLowerCaseRef), maxAlnsPerStartPos
        o.process_non_unique_primary, o.process_secondary, o.ignore5, o.ignore3 = int(p), int(s), ignore5, ignore3
        if h is not None:
            o.het_rate, o.het_rate_set = h, 1
        o.max_base_qs, o.min_quality, o.call_embedded_snvs, o.calc_strand_bias = maxBaseQS, minQuality, int(embeddedSNVs), int(csb)
        o.device, o.tile_len, o.shard, o.n_shards, o.host_only = device, tile_len, shard, n_shards, int(host_only)
        o.synchronous = int(synchronous)
        o.n_devices = int(n_devices)
        self._strs = knownSTRs.encode() if knownSTRs else None
        o.known_strs = self._strs
        self._known = knownVariants.encode() if knownVariants else None
        o.known_variants = self._known
        self.h = C.c_void_p()
        self._create(o)
        self._keep = []

    def _create(self, o: DetectorOptions) -> None:
        if self.lib.ngsep_ssvd_create(C.byref(o), C.byref(self.h)) != 0:
            raise NgsepError("ngsep_ssvd_create failed")

    def set_genome(self, genome: dict) -> None:
        names = [n.encode() for n in genome]
        arrs = [np.ascontiguousarray(v, np.uint8) for v in genome.values()]
        n = len(names)
        c_names = (C.c_char_p * n)(*names)
        c_len = (C.c_int64 * n)(*[a.size for a in arrs])
        c_bases = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        rc = self.lib.ngsep_ssvd_set_genome(self.h, n, c_names, c_len, c_bases)
        if rc != 0:
            raise NgsepError(f"ngsep_ssvd_set_genome failed ({rc})")

    def run(self, batch: AlignmentBatch, with_header: bool = True) -> str:
        hb, keep = _host_batch(batch)
        rc = self.lib.ngsep_ssvd_run(self.h, C.byref(hb), int(with_header))
        if rc != 0:
            raise NgsepError(f"ngsep_ssvd_run failed ({rc}): {self.lib.ngsep_ssvd_error(self.h).decode()}")
        n = C.c_size_t()
        p = self.lib.ngsep_ssvd_vcf(self.h, C.byref(n))
        return C.string_at(p, n.value).decode()

    def pack(self, batch: AlignmentBatch, seq_idx: int):
        """The tile the detector submits for sequence seq_idx -> dict(hdr, payload, qual, max_span, max_back, first, last, aligned_bases)."""
        hb, keep = _host_batch(batch)
        hdr_p, pay_p, q_p = C.c_void_p(), C.c_void_p(), C.c_void_p()
        n_reads, n_words, n_qual, bases = C.c_int64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        span, back, tf, tl = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        rc = self.lib.ngsep_ssvd_pack(self.h, C.byref(hb), seq_idx, C.byref(hdr_p), C.byref(pay_p), C.byref(q_p), C.byref(n_reads), C.byref(n_words),
                                      C.byref(n_qual), C.byref(span), C.byref(back), C.byref(tf), C.byref(tl), C.byref(bases))
        if rc != 0:
            raise NgsepError(f"ngsep_ssvd_pack failed ({rc}): {self.lib.ngsep_ssvd_error(self.h).decode()}")
        return dict(hdr=_take(hdr_p, n_reads.value, HDR_DTYPE, self.lib), payload=_take(pay_p, n_words.value, np.uint32, self.lib),
                    qual=_take(q_p, n_qual.value, np.uint8, self.lib), max_span=span.value, max_back=back.value, first=tf.value, last=tl.value,
                    aligned_bases=bases.value, quality=last_pack_quality(), layout=last_pack_layout())

    def stats(self) -> RunStats:
        s = RunStats()
        self.lib.ngsep_ssvd_stats(self.h, C.byref(s))
        return s

    def close(self) -> None:
        if self.h:
            self.lib.ngsep_ssvd_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultisampleDetector(Detector):
    """GPU drop-in for MultisampleVariantsDetector (reference: src/ngsep/discovery/MultisampleVariantsDetector.java, run() :421-462;
    option ids of main/CommandsDescriptor.xml:480-563).  ``read_groups`` is the @RG table of the input(s) as (ID, SM or None)
    pairs; ``AlignmentBatch.rg`` indexes it.  Samples and their VCF column order follow MSVD.loadSamplesFromAlignmentHeaders."""

    def __init__(self, read_groups, minAlleleDepthFrequency: float = 0.0, **options):
        self._rg = [(i.encode(), (s or "").encode()) for i, s in read_groups]
        self._madf = float(minAlleleDepthFrequency)
        for k in ("sampleId", "csb"):
            if k in options:
                raise TypeError(f"MultisampleVariantsDetector has no option -{k}")
        super().__init__(**options)

    def _create(self, o: DetectorOptions) -> None:
        n = len(self._rg)
        ids = (C.c_char_p * n)(*[r[0] for r in self._rg])
        sms = (C.c_char_p * n)(*[r[1] for r in self._rg])
        if self.lib.ngsep_msvd_create(C.byref(o), n, ids, sms, self._madf, C.byref(self.h)) != 0:
            raise NgsepError("ngsep_msvd_create failed")
