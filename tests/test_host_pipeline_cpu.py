"""CPU-side checks of the C++ host pipeline (libngsep_host.so) that need no GPU: packer equality with the numpy
packer, and the event pre-pass (indel realigner + host indel/STR caller) against the oracle's indel records."""
import numpy as np
import pytest

from ngsepcore_b200 import hostapi, pypack, simulate
from tests import oracle_util as ou


def test_host_symbols():
    import os, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "ngsep_host.h")).read()
    declared = set(re.findall(r"\b(ngsep_[a-z_]+)\s*\(", hdr))
    assert declared == set(hostapi.EXPORTS)
    lib = hostapi.load_library()
    for name in declared:
        assert hasattr(lib, name)


def test_cpp_packer_equals_numpy_packer(monkeypatch):
    genome, batch, _ = simulate.simulate(contigs=[("a", 30000), ("b", 20000)], depth=12, seeds=(3, 4, 5), n_frac=0.02, low_mapq_frac=0.1)
    monkeypatch.setenv("NGSEP_QUAL_BITS", "8")       # quality bytes as they are: the layout of the numpy packer
    for ci in range(2):
        idx = pypack.admit(batch, max_alns_per_start=2)
        idx = idx[batch.seq_idx[idx] == ci]
        h1, p1, q1, s1 = pypack.pack_tile(batch, idx, ignore5=2, ignore3=4)
        h2, p2, q2, s2, nb = hostapi.pack_sequence(batch, ci, max_alns_per_start=2, ignore5=2, ignore3=4)
        assert hostapi.pack_sequence.quality[0] == 8
        assert s1 == s2 and nb == int((h1["meta"] & 0xFFFF).sum())
        assert np.array_equal(h1, h2)
        assert np.array_equal(p1, p2)
        assert np.array_equal(q1, q2)


def _quality_class(q):      # what the path can tell apart: every q <= 3 alike (CH:214), min(30, q) otherwise (CH:91)
    v = q.astype(np.int32) - 33
    return np.where(v <= 3, 0, np.minimum(v, 30))


@pytest.mark.parametrize("alphabet,bits", [(None, 2), ((2, 11, 15, 22, 27, 33, 37, 40), 4), (tuple(range(2, 42)), 8)])
def test_quality_dictionary_modes(monkeypatch, alphabet, bits):
    """ngsep_tile.qual_bits: a tile whose qualities fall into at most 4 / 16 classes travels as 2- / 4-bit codes.  Decoding the
    packed array through the dictionary gives, base by base, the class of the original quality; everything else of the tile
    (headers, CIGARs, bases, masks) is what the 8-bit packing produces."""
    monkeypatch.delenv("NGSEP_QUAL_BITS", raising=False)
    kw = {} if alphabet is None else dict(qual_vals=alphabet)
    genome, batch, _ = simulate.simulate(contigs=[("a", 40000)], depth=15, seeds=(13, 14, 15), n_frac=0.01, read_len=101, **kw)
    h2, p2, q2, s2, nb = hostapi.pack_sequence(batch, 0)
    quality = hostapi.pack_sequence.quality
    assert quality[0] == bits
    monkeypatch.setenv("NGSEP_QUAL_BITS", "8")
    h8, p8, q8, s8, _ = hostapi.pack_sequence(batch, 0)
    assert np.array_equal(h2, h8) and s2 == s8
    nxt = np.concatenate([h2["off"][1:], [p2.size]]).astype(np.int64) - 1       # the qoff word of every record
    mask = np.ones(p2.size, bool)
    mask[nxt] = False
    assert np.array_equal(p2[mask], p8[mask])
    if bits == 8:
        assert np.array_equal(q2, q8) and np.array_equal(p2, p8)
        return
    assert q2.size * 8 <= q8.size * bits + 4 * 8 * h2.shape[0]                   # the array really shrank (runs are padded to 4 bytes)
    dec = hostapi.decode_quality(q2, h2, p2, quality)
    assert dec.size == q8.size
    L = 101
    rows = dec.reshape(-1, 104)[:, :L], q8.reshape(-1, 104)[:, :L]
    assert np.array_equal(_quality_class(rows[0]), _quality_class(rows[1]))


@pytest.mark.parametrize("seed,kw", [
    (1, dict(indel_rate=1e-3, depth=30)),
    (2, dict(indel_rate=2e-3, depth=12, indel_jitter=0.6)),
    (3, dict(indel_rate=1e-3, depth=45, ungapped_end_bp=12)),
])
def test_event_prepass_indel_records_match_oracle(oracle, tmp_path, seed, kw):
    genome, batch, _ = simulate.simulate(contigs=[("chrA", 120000)], seeds=(seed * 10, seed * 10 + 1, seed * 10 + 2), **kw)
    fa, sam = str(tmp_path / "r.fa"), str(tmp_path / "r.sam")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    want, _, _, _ = ou.run_ssvd(oracle, fa, sam)
    want_lines = [l for l in want.splitlines() if "TYPE=INDEL" in l or "TYPE=STR" in l]
    det = hostapi.Detector(host_only=True)
    det.set_genome(genome)
    got_lines = det.run(batch, with_header=False).splitlines()
    assert len(want_lines) > 20
    assert got_lines == want_lines
    assert det.stats().events >= len(want_lines)


def test_header_text_matches_oracle(oracle, tmp_path):
    genome, batch, _ = simulate.simulate(contigs=[("c", 3000)], depth=3)
    fa, sam = str(tmp_path / "r.fa"), str(tmp_path / "r.sam")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    want, _, _, _ = ou.run_ssvd(oracle, fa, sam, args=("-sampleId", "S1", "-psp"))
    det = hostapi.Detector(host_only=True, sampleId="S1", psp=True)
    det.set_genome(genome)
    got = det.run(batch)
    hdr = lambda t: [l for l in t.splitlines() if l.startswith("#")]
    assert hdr(got) == hdr(want)
    # fixed parts of the header are the reference's own (training/yeastDemo_...vcf.gz carries the same lines)
    assert hdr(got)[0] == "##fileformat=VCFv4.2"
    assert '##FORMAT=<ID=BSDP,Number=4,Type=Integer,Description="Number of base calls (depth) for the 4 nucleotides in called SNVs sorted as A,C,G,T">' in hdr(got)


@pytest.mark.parametrize("seed,kw,det_kw,oargs", [
    (1, dict(n_samples=8, depth=10, mode="wgs", indel_rate=1.5e-3, rgs_per_sample=2, no_rg_frac=0.03), {}, ()),
    (2, dict(n_samples=12, depth=8, mode="gbs", n_tags=150, indel_rate=2e-3), dict(maxAlnsPerStartPos=100), ("-maxAlnsPerStartPos", "100")),
    (3, dict(n_samples=5, depth=14, mode="wgs", indel_rate=2e-3, indel_jitter=0.6), dict(minQuality=60, h=0.01), ("-minQuality", "60", "-h", "0.01")),
])
def test_population_indel_records_match_oracle(oracle, tmp_path, seed, kw, det_kw, oargs):
    """MultisampleVariantsDetector, host side only (no GPU): the population indel / STR records — pooled allele clustering,
    the allele refinement loop, per-sample VD.callIndel(variant), DiversityStatistics INFO — against the oracle's."""
    genome, batch, _ = simulate.simulate_population([("chrA", 60000), ("chrB", 30000)], seeds=(seed * 100, seed * 100 + 1, seed * 100 + 2), **kw)
    fa, sam = str(tmp_path / "r.fa"), str(tmp_path / "r.sam")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    want, _, _, _ = ou.run_msvd(oracle, fa, [sam], args=oargs)
    want_lines = [l for l in want.splitlines() if "TYPE=INDEL" in l or "TYPE=STR" in l]
    det = hostapi.MultisampleDetector(list(zip(batch.rg_names, batch.rg_samples)), host_only=True, **det_kw)
    det.set_genome(genome)
    got = det.run(batch)
    got_lines = [l for l in got.splitlines() if not l.startswith("#")]
    assert len(want_lines) >= 8
    assert got_lines == want_lines
    hdr = lambda t: [l for l in t.splitlines() if l.startswith("#")]
    assert hdr(got) == hdr(want)      # sample columns in TreeMap order, MSVD:494-515


def test_threaded_packer_equals_single_thread(monkeypatch):
    """Tiles above 20 k reads are packed by several host threads (contiguous read ranges, concatenated with payload / quality
    offset fix-ups): byte-identical to the single-thread packing, read versions of the realigner pre-pass included."""
    genome, batch, _ = simulate.simulate(contigs=[("chrA", 250000)], depth=30, seeds=(81, 82, 83), indel_rate=1e-3, n_frac=0.01)
    out = {}
    for threads in ("1", "5", "8"):
        monkeypatch.setenv("NGSEP_HOST_THREADS", threads)
        det = hostapi.Detector(host_only=True)
        det.set_genome(genome)
        out[threads] = det.pack(batch, 0)
    ref = out["1"]
    assert ref["hdr"].shape[0] > 70000 and (ref["hdr"]["meta"] & 0x04000000).any()      # realigner versions are present
    for threads in ("5", "8"):
        t = out[threads]
        for k in ("hdr", "payload", "qual"):
            assert np.array_equal(ref[k], t[k]), (threads, k)
        assert (ref["max_span"], ref["max_back"], ref["aligned_bases"]) == (t["max_span"], t["max_back"], t["aligned_bases"])


def test_plan_without_materialised_reads_packs_the_same_tile(monkeypatch):
    """The detectors keep an admitted read as 16 bytes (LiteRead) and build its HostRead only when the read has any CIGAR but one
    match block or lies under a realigner event; untouched reads are packed straight from the batch.  On an input without indels
    the tile must be, byte for byte, what admit_sequence + pack_reads (every read materialised) give — with the per-start cap,
    ignored ends, low MAPQ records, bases outside ACGT and a query window."""
    monkeypatch.setenv("NGSEP_WIRE_COMPACT", "0")
    genome, batch, _ = simulate.simulate(contigs=[("a", 60000), ("b", 30000)], depth=25, seeds=(33, 34, 35), n_frac=0.02, low_mapq_frac=0.1, indel_rate=0.0,
                                         read_len=93, n_read_groups=3)
    det = hostapi.Detector(host_only=True, maxAlnsPerStartPos=2, ignore5=2, ignore3=4)
    det.set_genome(genome)
    for ci in range(2):
        h1, p1, q1, s1, nb = hostapi.pack_sequence(batch, ci, max_alns_per_start=2, ignore5=2, ignore3=4)
        t = det.pack(batch, ci)
        assert h1.shape[0] > 5000
        assert np.array_equal(h1, t["hdr"]) and np.array_equal(p1, t["payload"]) and np.array_equal(q1, t["qual"])
        assert (s1, nb) == (t["max_span"], t["aligned_bases"])
    # the admission of one long sequence (> 400 k records) on several threads, cut where the start position changes
    genome, batch, _ = simulate.simulate(contigs=[("long", 1500000)], depth=30, seeds=(36, 37, 38), low_mapq_frac=0.05, indel_rate=0.0)
    assert batch.n > 400000
    want = hostapi.pack_sequence(batch, 0, max_alns_per_start=2, ignore5=2, ignore3=4)
    det.set_genome(genome)
    for threads in ("1", "4"):
        monkeypatch.setenv("NGSEP_HOST_THREADS", threads)
        t = det.pack(batch, 0)
        assert np.array_equal(want[0], t["hdr"]) and np.array_equal(want[1], t["payload"]) and np.array_equal(want[2], t["qual"])


def test_quality_coder_taken_over_from_the_previous_tile(monkeypatch):
    """run_device hands the quality coding of a tile on to the next one (QualHint): the packer then skips its pass over the
    qualities and notices a class without a code while it packs.  Same alphabet: the same bytes as a fresh scan.  A tile with a
    new class: packed again the long way (4-bit codes here), decoding to the classes of the original qualities.  A tile that
    uses fewer classes than the coder knows: still decodes right."""
    monkeypatch.setenv("NGSEP_PACK_HINT", "1")
    mk = lambda seeds, **kw: simulate.simulate(contigs=[("a", 40000)], depth=15, seeds=seeds, n_frac=0.01, read_len=101, **kw)
    genome, b1, _ = mk((13, 14, 15))
    _, b2, _ = mk((13, 14, 16), genome=genome)
    _, b3, _ = mk((13, 14, 17), genome=genome, qual_vals=(2, 11, 15, 22, 27, 33, 37, 40))
    _, b4, _ = mk((13, 14, 18), genome=genome, qual_vals=(15, 37))
    det = hostapi.Detector(host_only=True); det.set_genome(genome)
    fresh = lambda b: (lambda d: (d.set_genome(genome), d.pack(b, 0))[1])(hostapi.Detector(host_only=True))
    t1, t2 = det.pack(b1, 0), det.pack(b2, 0)          # t2 is packed with t1's coder
    f2 = fresh(b2)
    assert t1["quality"][0] == 2 and t2["quality"] == f2["quality"]
    for k in ("hdr", "payload", "qual"):
        assert np.array_equal(t2[k], f2[k]), k
    t3, f3 = det.pack(b3, 0), fresh(b3)                # new classes: the long way, like a fresh detector
    assert t3["quality"][0] == 4 and t3["quality"] == f3["quality"] and np.array_equal(t3["qual"], f3["qual"])
    t4 = det.pack(b4, 0)                               # two of the eight classes: t3's coder still fits
    assert t4["quality"] == t3["quality"]
    monkeypatch.setenv("NGSEP_QUAL_BITS", "8")
    monkeypatch.delenv("NGSEP_PACK_HINT")
    t8 = fresh(b4)
    got = hostapi.decode_quality(t4["qual"], t4["hdr"], t4["payload"], t4["quality"])
    want = hostapi.decode_quality(t8["qual"], t8["hdr"], t8["payload"], t8["quality"])
    assert np.array_equal(_quality_class(got), _quality_class(want))


def test_vector_packer_equals_portable_packer(tmp_path):
    """The packer codes 32 bases / qualities at a time where the host has AVX2 (chosen at run time); NGSEP_NO_AVX2=1 keeps the
    portable 8-at-a-time code.  Both give the same bytes: odd read lengths, bases outside ACGT, 2- / 4- / 8-bit quality modes."""
    import hashlib, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "pack_digest.py"
    script.write_text(
        "import sys, hashlib\n"
        f"sys.path.insert(0, {root!r})\n"
        "from ngsepcore_b200 import hostapi, simulate\n"
        "for L, qv in ((101, None), (77, (2, 11, 15, 22, 27, 33, 37, 40)), (150, tuple(range(2, 42))), (31, None), (64, None)):\n"
        "    kw = {} if qv is None else dict(qual_vals=qv)\n"
        "    genome, batch, _ = simulate.simulate(contigs=[('a', 30000)], depth=12, seeds=(23, 24, 25), n_frac=0.02, read_len=L, indel_rate=1e-3, **kw)\n"
        "    det = hostapi.Detector(host_only=True); det.set_genome(genome)\n"
        "    t = det.pack(batch, 0)\n"
        "    print(L, t['quality'][0], t['layout'], *[hashlib.sha1(t[k].tobytes()).hexdigest() for k in ('hdr', 'payload', 'qual')])\n")
    outs = []
    for no_avx2 in ("0", "1"):
        env = dict(os.environ, NGSEP_NO_AVX2=no_avx2)
        outs.append(subprocess.run([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, text=True, check=True).stdout)
    assert outs[0] == outs[1] and len(outs[0].splitlines()) == 5
    assert {l.split()[1] for l in outs[0].splitlines()} == {"2", "4", "8"}


def test_merge_batches_follows_process_files(oracle, tmp_path):
    """hostapi.merge_batches = AlignmentsPileupGenerator.processFiles (APG:212-289): one batch per input file, consumed in the
    reference's order (k-way merge by sequence, first, last; ties to the lowest input).  The oracle reads the three SAM files;
    the host pipeline gets the merged batch; their population indel records must agree — and the merged order must be the
    oracle-independent rule itself."""
    genome, batch, _ = simulate.simulate_population([("chrA", 50000), ("chrB", 20000)], n_samples=6, depth=10, mode="wgs", indel_rate=2.5e-3,
                                                    seeds=(301, 302, 303))
    fa = str(tmp_path / "r.fa")
    simulate.write_fasta(fa, genome)
    parts, files = [], []
    for k in range(3):
        part = batch.select(np.nonzero((batch.rg // 2) == k)[0])
        parts.append(part)
        path = str(tmp_path / f"part{k}.sam")
        simulate.write_sam(path, part)
        files.append(path)
    merged = hostapi.merge_batches(parts)
    assert merged.n == batch.n - int((batch.mapq < 20).sum())          # each input's reader drops its non-unique records first
    # the rule itself: non-decreasing (sequence, first); within equal (sequence, first, last) the lower input comes first
    key = merged.seq_idx.astype(np.int64) * (1 << 32) + merged.first
    assert (np.diff(key) >= 0).all()
    want, _, _, _ = ou.run_msvd(oracle, fa, files)
    want_lines = [l for l in want.splitlines() if "TYPE=INDEL" in l or "TYPE=STR" in l]
    det = hostapi.MultisampleDetector(list(zip(merged.rg_names, merged.rg_samples)), host_only=True)
    det.set_genome(genome)
    got_lines = [l for l in det.run(merged).splitlines() if not l.startswith("#")]
    assert len(want_lines) > 15
    assert got_lines == want_lines


def _write_str_loci(path, truth, contigs, rng, per_contig_random=40):
    """A -knownSTRs file (sequence, first, last + extra columns as in the tutorial's TRF table): loci around about half of the planted
    indels, loci without any indel, neighbouring loci that makeNonRedundantSTRs merges (gap <= 2) or keeps apart (gap > 5)."""
    rows = []
    for name, length in contigs:
        for v in truth[name]:
            if v.kind != "snv" and rng.random() < 0.5:
                f = max(2, v.pos - int(rng.integers(1, 6)))
                l = min(length - 2, v.pos + v.length + int(rng.integers(2, 14)))
                rows.append((name, f, l))
                if rng.random() < 0.3:
                    rows.append((name, l + int(rng.integers(0, 3)), min(length - 2, l + 12)))     # merged with the previous one
                elif rng.random() < 0.3:
                    rows.append((name, l + 7, min(length - 2, l + 20)))                              # a separate locus right behind
        for _ in range(per_contig_random):
            f = int(rng.integers(50, length - 100))
            rows.append((name, f, f + int(rng.integers(8, 40))))
    order = {n: i for i, (n, _) in enumerate(contigs)}
    rows.sort(key=lambda r: (order[r[0]], r[1], r[2]))
    with open(path, "w") as fh:
        for k, (n, f, l) in enumerate(rows):
            fh.write(f"{n} {f} {l} 2 {l - f + 1}.0 2 90 5 40 25 25 25 25 1.9 AC ACACAC\n" if k % 2 else f"{n}\t{f}\t{l}\n")
        fh.write("chrUnknown 5 20\nchrA x y\n")     # an unknown sequence and a malformed line are skipped
    return len(rows)


@pytest.mark.parametrize("seed,kw,opts", [
    (4, dict(indel_rate=2e-3, depth=30), ()),
    (5, dict(indel_rate=3e-3, depth=18, indel_jitter=0.6), ("-minQuality", "20")),
])
def test_known_strs_indel_records_match_oracle(oracle, tmp_path, seed, kw, opts):
    """-knownSTRs (SSVD:906-913, IRPL:89-99): input loci are fixed events — every indel inside moves to the locus start, the call
    spans the locus, TYPE=STR, a locus without a length change is still a candidate (VD:307) — and positions inside are embedded.
    Host stages only (no GPU): the indel / STR records of the C++ host pipeline against the oracle."""
    contigs = [("chrA", 90000), ("chrB", 40000)]
    genome, batch, truth = simulate.simulate(contigs=contigs, seeds=(seed * 10, seed * 10 + 1, seed * 10 + 2), **kw)
    fa, sam, strs = str(tmp_path / "r.fa"), str(tmp_path / "r.sam"), str(tmp_path / "strs.txt")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    assert _write_str_loci(strs, truth, contigs, np.random.default_rng(seed)) > 100
    want, _, _, _ = ou.run_ssvd(oracle, fa, sam, args=("-knownSTRs", strs) + opts)
    want_lines = [l for l in want.splitlines() if "TYPE=INDEL" in l or "TYPE=STR" in l]
    plain, _, _, _ = ou.run_ssvd(oracle, fa, sam, args=opts)
    assert want != plain and sum("TYPE=STR" in l for l in want_lines) > 10
    det_kw = {"minQuality": int(opts[1])} if opts else {}
    det = hostapi.Detector(host_only=True, knownSTRs=strs, **det_kw)
    det.set_genome(genome)
    got_lines = [l for l in det.run(batch).splitlines() if not l.startswith("#")]
    assert got_lines == want_lines


def _is_snv_record(line):
    t = line.split("\t")
    return all(len(a) == 1 and a in "ACGT" for a in [t[3]] + t[4].split(","))


@pytest.mark.parametrize("seed,kw,opts,det_kw", [
    (4, dict(indel_rate=1e-3, depth=12), (), {}),
    (5, dict(indel_rate=3e-3, depth=2), ("-minQuality", "20"), dict(minQuality=20)),                 # coverage gaps: variants without a pileup, empty pileups
    (6, dict(indel_rate=2e-3, depth=20, indel_jitter=0.6), ("-querySeq", "chrA", "-first", "5000", "-last", "41000"), dict(querySeq="chrA", first=5000, last=41000)),
    (7, dict(indel_rate=1e-3, depth=30), ("-ploidy", "4"), dict(ploidy=4)),      # the pool branch of genotypeVariantSample (SSVPL:376-377)
])
def test_known_variants_host_records_match_oracle(oracle, tmp_path, seed, kw, opts, det_kw):
    """-knownVariants (SSVD:897-905): the input variants are the realigner's fixed events (IRPL:89-99, 141-157) and every one of them
    that meets a pileup is genotyped (SSVPL:162-176) — homozygous-reference and undecided results are records too, with the ID and
    QUAL of the input.  Host stages only (no GPU): the records of the variants that are NOT SNVs (VD.callIndel with the variant given)
    against the oracle; the SNV sites are device work (tests/test_gpu_detector_parity.py::test_known_variants_mode)."""
    from tests.known_util import make_known_vcf
    contigs = [("chrA", 60000), ("chrB", 30000)]
    genome, batch, truth = simulate.simulate(contigs=contigs, seeds=(seed * 10, seed * 10 + 1, seed * 10 + 2), **kw)
    fa, sam, kv = str(tmp_path / "r.fa"), str(tmp_path / "r.sam"), str(tmp_path / "known.vcf")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    discovered, _, _, _ = ou.run_ssvd(oracle, fa, sam, args=("-minQuality", "0"))
    assert make_known_vcf(kv, discovered, genome, contigs, np.random.default_rng(seed)) > 200
    want, _, _, _ = ou.run_ssvd(oracle, fa, sam, args=("-knownVariants", kv) + opts)
    want_lines = [l for l in want.splitlines() if not l.startswith("#") and not _is_snv_record(l)]
    gts = {l.split("\t")[9].split(":")[0] for l in want_lines}
    assert len(want_lines) > 40 and {"0/0", "./."} <= gts and len(gts) >= 3
    det = hostapi.Detector(host_only=True, knownVariants=kv, **det_kw)
    det.set_genome(genome)
    got_lines = [l for l in det.run(batch).splitlines() if not l.startswith("#")]
    assert got_lines == want_lines


def test_known_variants_loader_and_empty_list(oracle, tmp_path):
    """An input VCF without a usable record means discovery (SSVPL:147: inputVariants.size() == 0); a file without the #CHROM line is
    an error (VCFFileHeader.java:165-167)."""
    contigs = [("chrA", 30000)]
    genome, batch, _ = simulate.simulate(contigs=contigs, seeds=(71, 72, 73), depth=15, indel_rate=2e-3)
    fa, sam, kv = str(tmp_path / "r.fa"), str(tmp_path / "r.sam"), str(tmp_path / "known.vcf")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    open(kv, "w").write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\nchrA\t100\t.\tA\t<DEL>\t.\t.\tSVTYPE=DEL\nchrA\t200\t.\tA\t.\t.\t.\t.\n")
    want, _, _, _ = ou.run_ssvd(oracle, fa, sam, args=("-knownVariants", kv))
    plain, _, _, _ = ou.run_ssvd(oracle, fa, sam)
    assert want == plain
    det = hostapi.Detector(host_only=True, knownVariants=kv)
    det.set_genome(genome)
    assert [l for l in det.run(batch).splitlines() if not l.startswith("#")] == [l for l in plain.splitlines() if "TYPE=INDEL" in l or "TYPE=STR" in l]
    open(kv, "w").write("chrA\t100\t.\tA\tC\t.\t.\t.\n")
    bad = hostapi.Detector(host_only=True, knownVariants=kv)
    bad.set_genome(genome)
    with pytest.raises(Exception):
        bad.run(batch)


def test_admission_inside_one_sequence_on_several_threads(oracle, tmp_path, monkeypatch):
    """A run with fewer sequences than host threads admits a long sequence piece by piece (cut where the start position changes:
    the duplicate-record rule and the per-start cap never look across such a cut).  Same host-made records on one thread, on
    eight, and from the oracle; with a cap of one alignment per start, thousands of start groups lose reads."""
    contigs = [("chrL", 700000)]
    genome, batch, _ = simulate.simulate(contigs=contigs, seeds=(91, 92, 93), depth=32, indel_rate=1e-3)
    assert batch.n > 200000
    fa, sam = str(tmp_path / "r.fa"), str(tmp_path / "r.sam")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    want, _, _, _ = ou.run_ssvd(oracle, fa, sam, args=("-maxAlnsPerStartPos", "1"))
    want_lines = [l for l in want.splitlines() if "TYPE=INDEL" in l or "TYPE=STR" in l]
    outs = []
    for n in ("1", "8"):
        monkeypatch.setenv("NGSEP_HOST_THREADS", n)
        det = hostapi.Detector(host_only=True, maxAlnsPerStartPos=1)
        det.set_genome(genome)
        outs.append([l for l in det.run(batch).splitlines() if not l.startswith("#")])
        assert det.stats().admitted < batch.n
    assert outs[0] == outs[1] == want_lines and len(want_lines) > 300


def test_compact_wire_layout_is_the_full_layout_minus_what_the_device_rebuilds(monkeypatch):
    """NGSEP_LAYOUT_COMPACT (what the detectors submit): the payload of a record flagged SIMPLE is [bases2, qoff] — its CIGAR word
    (len*8+3) and low-quality mask stay on the host —, every other record's payload is unchanged, hdr.off index the compact stream,
    and the reported n_words is the length of the full layout.  Dropping those words from the full packing of the same reads gives
    the compact stream word for word."""
    from ngsepcore_b200 import capi
    genome, batch, _ = simulate.simulate(contigs=[("a", 60000)], depth=20, seeds=(23, 24, 25), indel_rate=2e-3, n_frac=0.01)
    det = hostapi.Detector(host_only=True)
    det.set_genome(genome)
    monkeypatch.setenv("NGSEP_WIRE_COMPACT", "0")
    full = det.pack(batch, 0)
    monkeypatch.delenv("NGSEP_WIRE_COMPACT")
    comp = det.pack(batch, 0)
    assert full["layout"] == (0, full["payload"].size) and comp["layout"] == (1, full["payload"].size)
    assert np.array_equal(full["qual"], comp["qual"]) and full["quality"] == comp["quality"]
    hf, hc = full["hdr"], comp["hdr"]
    for name in ("first", "meta", "ign_start", "ign_end"):
        assert np.array_equal(hf[name], hc[name])
    simple = (hf["meta"] & capi.M_SIMPLE) != 0
    assert 0.5 < simple.mean() < 1.0                      # indel reads, realigner versions and reads with N keep their payload
    L = (hf["meta"] & 0xFFFF).astype(np.int64)
    nb, nw = (L + 15) // 16, (L + 31) // 32
    end_f = np.concatenate([hf["off"][1:], [full["payload"].size]]).astype(np.int64)
    size_f = end_f - hf["off"]
    assert np.array_equal(size_f[simple], (1 + nb + nw + 1)[simple])
    size_c = np.where(simple, nb + 1, size_f)
    off_c = np.concatenate([[0], np.cumsum(size_c)[:-1]])
    assert np.array_equal(off_c, hc["off"]) and int(size_c.sum()) == comp["payload"].size
    keep = np.ones(full["payload"].size, bool)
    for i in np.nonzero(simple)[0]:
        o = int(hf["off"][i])
        assert full["payload"][o] == (L[i] << 3 | 3)
        keep[o] = False
        keep[o + 1 + nb[i]: o + 1 + nb[i] + nw[i]] = False
    assert np.array_equal(full["payload"][keep], comp["payload"])
    assert comp["payload"].nbytes < 0.75 * full["payload"].nbytes


@pytest.mark.parametrize("seed,depth,ploidy,opts,det_kw", [
    (61, 40, 4, (), {}),
    (71, 25, 3, ("-minQuality", "20"), dict(minQuality=20)),
    (81, 60, 6, ("-maxBaseQS", "25", "-h", "0.01"), dict(maxBaseQS=25, h=0.01)),
])
def test_pool_indel_records_match_oracle(oracle, tmp_path, seed, depth, ploidy, opts, det_kw):
    """Pools / polyploids (-ploidy >= 3): discoverIndel's pool branch (SSVPL:281-295) — createIndelVariantPool, genotypeVariantPool
    (SSVPL:402-503) with one CountsHelper per allele-frequency hypothesis k / ploidy, the second round with the reduced allele list —
    on the host, from the per-call allele tallies.  Host stages only: the indel records (no TYPE: a pool variant is a plain
    GenomicVariantImpl) against the oracle's; the SNV half needs the device (test_gpu_detector_parity.py::test_pools_and_polyploids)."""
    genome, batch, _ = simulate.simulate(contigs=[("chrA", 80000), ("chrB", 30000)], depth=depth, seeds=(seed, seed + 1, seed + 2), indel_rate=2e-3, snv_rate=3e-3)
    fa, sam = str(tmp_path / "r.fa"), str(tmp_path / "r.sam")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    want, _, _, _ = ou.run_ssvd(oracle, fa, sam, args=("-ploidy", str(ploidy)) + opts)
    want_lines = [l for l in want.splitlines() if not l.startswith("#") and not _is_snv_record(l)]
    assert len(want_lines) > 100 and all("TYPE=" not in l and l.split("\t")[5] == "0" for l in want_lines)
    det = hostapi.Detector(host_only=True, ploidy=ploidy, **det_kw)
    det.set_genome(genome)
    assert [l for l in det.run(batch).splitlines() if not l.startswith("#")] == want_lines
