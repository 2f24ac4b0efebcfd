"""GPU parity, whole detector: host pipeline + CUDA kernels (C ABI) vs the oracle, VCF text byte for byte.

Inputs carry planted SNVs and indels (placement jitter, indels absorbed ungapped at read ends), so the indel realigner
fires, reads get versioned, indel records come from the host caller and SNV records from the device."""
import os
import subprocess

import numpy as np
import pytest

from ngsepcore_b200 import hostapi, simulate
from tests import oracle_util as ou

pytestmark = pytest.mark.gpu


def _inputs(tmp_path, contigs, **kw):
    genome, batch, truth = simulate.simulate(contigs=contigs, **kw)
    fa, sam = str(tmp_path / "ref.fa"), str(tmp_path / "reads.sam")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    return genome, batch, fa, sam


def _body(text):
    return [l for l in text.splitlines() if not l.startswith("#")]


def test_default_options(oracle, tmp_path):
    genome, batch, fa, sam = _inputs(tmp_path, [("chrA", 300000), ("chrB", 150000)], depth=30, seeds=(41, 42, 43))
    want, _, gp, _ = ou.run_ssvd(oracle, fa, sam)
    det = hostapi.Detector()
    det.set_genome(genome)
    got = det.run(batch)
    st = det.stats()
    assert got == want
    assert st.covered_positions == gp
    assert st.indel_calls > 20 and st.snv_calls > 300 and st.versions > 50 and st.kernel_launches > 0


@pytest.mark.parametrize("seed,sim_kw,det_kw,oargs", [
    (5, dict(indel_rate=1e-3, depth=25), dict(embeddedSNVs=True, csb=True), ("-embeddedSNVs", "-csb")),
    (6, dict(indel_rate=2e-3, depth=12, indel_jitter=0.6, snv_rate=3e-3), dict(minQuality=20, maxBaseQS=25, h=0.01), ("-minQuality", "20", "-maxBaseQS", "25", "-h", "0.01")),
    (7, dict(indel_rate=5e-4, depth=60, ungapped_end_bp=12, n_frac=0.01, lower_frac=0.03), dict(ignoreLowerCaseRef=True, maxAlnsPerStartPos=2, ignore5=2, ignore3=5),
     ("-ignoreLowerCaseRef", "-maxAlnsPerStartPos", "2", "-ignore5", "2", "-ignore3", "5")),
    (8, dict(indel_rate=1e-3, depth=20), dict(ploidy=1), ("-ploidy", "1")),
])
def test_option_matrix(oracle, tmp_path, seed, sim_kw, det_kw, oargs):
    genome, batch, fa, sam = _inputs(tmp_path, [("chrA", 200000)], seeds=(seed * 10, seed * 10 + 1, seed * 10 + 2), **sim_kw)
    want, _, _, _ = ou.run_ssvd(oracle, fa, sam, args=oargs)
    det = hostapi.Detector(**det_kw)
    det.set_genome(genome)
    got = det.run(batch)
    assert len(_body(want)) > 100
    assert got == want


def test_region_tiles_and_shards_are_invariant(oracle, tmp_path):
    """Region tiles with a read-length halo, dealt round-robin to 1/2/3 shards, give the same records."""
    genome, batch, fa, sam = _inputs(tmp_path, [("chrA", 260000), ("chrB", 90000)], depth=30, seeds=(91, 92, 93), indel_rate=5e-4)
    want, _, _, _ = ou.run_ssvd(oracle, fa, sam)
    for n_shards in (1, 2, 3):
        lines = []
        for shard in range(n_shards):
            det = hostapi.Detector(tile_len=37000, shard=shard, n_shards=n_shards)
            det.set_genome(genome)
            lines += _body(det.run(batch, with_header=False))
        order = {n: i for i, n in enumerate(genome)}
        lines.sort(key=lambda l: (order[l.split("\t")[0]], int(l.split("\t")[1])))
        assert lines == _body(want)


def test_query_region(oracle, tmp_path):
    genome, batch, fa, sam = _inputs(tmp_path, [("chrA", 120000), ("chrB", 120000)], depth=20, seeds=(61, 62, 63), indel_rate=1e-3)
    args = ("-querySeq", "chrB", "-first", "30000", "-last", "80000")
    want, _, _, _ = ou.run_ssvd(oracle, fa, sam, args=args)
    det = hostapi.Detector(querySeq="chrB", first=30000, last=80000)
    det.set_genome(genome)
    assert det.run(batch) == want
    assert len(_body(want)) > 30


def test_cli_sam_to_vcf(oracle, tmp_path):
    genome, batch, fa, sam = _inputs(tmp_path, [("chrA", 80000)], depth=20, seeds=(71, 72, 73))
    want, _, _, _ = ou.run_ssvd(oracle, fa, sam, args=("-sampleId", "NA1"))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "ngsepcore_b200", "ngsep_b200")
    out = str(tmp_path / "out")
    subprocess.check_call([cli, "SingleSampleVariantsDetector", "-i", sam, "-r", fa, "-o", out, "-sampleId", "NA1"])
    assert open(out + ".vcf").read() == want
    # error behaviour of the reference module: missing required parameter / unknown option
    assert subprocess.call([cli, "SingleSampleVariantsDetector", "-r", fa, "-o", out], stderr=subprocess.DEVNULL) == 2
    assert subprocess.call([cli, "SingleSampleVariantsDetector", "-i", sam, "-r", fa, "-o", out, "-bogus"], stderr=subprocess.DEVNULL) == 1


def test_pipelined_run_equals_synchronous_run(oracle, tmp_path):
    """The product path — sequences planned in parallel on the host threads, region tiles packed into page-locked ring buffers
    with up to three tiles in flight — gives the bytes of the synchronous path (one tile at a time from pageable buffers, one host
    thread) and of the oracle; a second run of the same detector object (buffers reused) as well."""
    genome, batch, fa, sam = _inputs(tmp_path, [("chrA", 420000), ("chrB", 180000), ("chrC", 90000), ("chrD", 40000)], depth=30, seeds=(191, 192, 193),
                                     indel_rate=5e-4)
    want, _, gp, _ = ou.run_ssvd(oracle, fa, sam)
    sync = hostapi.Detector(tile_len=50000, synchronous=True)
    sync.set_genome(genome)
    ref = sync.run(batch)
    assert ref == want
    for tile_len in (50000, 0):
        det = hostapi.Detector(tile_len=tile_len)
        det.set_genome(genome)
        assert det.run(batch) == want
        s1 = det.stats()
        assert det.run(batch) == want
        s2 = det.stats()
        assert s1.covered_positions == gp and s2.covered_positions == 2 * gp
        assert s1.tiles == (16 if tile_len else 4)


def test_results_behind_the_kernels_equal_copies_at_fetch(oracle, tmp_path, monkeypatch):
    """A pipelined tile's counters, staging report and first 2048 call records come back behind its kernels into page-locked memory
    (ngsep_tile_fetch reads host memory); NGSEP_PREFETCH=0 copies them when they are fetched, as before.  Same records either way —
    with a tile that holds more than 2048 calls (the rest is copied at fetch time), with many small tiles, and for the oracle."""
    genome, batch, fa, sam = _inputs(tmp_path, [("chrA", 260000), ("chrB", 60000)], depth=25, seeds=(461, 462, 463), snv_rate=1.2e-2, indel_rate=3e-4)
    want, _, gp, _ = ou.run_ssvd(oracle, fa, sam)
    assert sum(1 for l in _body(want) if l.startswith("chrA\t")) > 2500
    for prefetch in ("1", "0"):
        monkeypatch.setenv("NGSEP_PREFETCH", prefetch)       # (read when the context is created)
        for tile_len in (0, 20000):
            det = hostapi.Detector(tile_len=tile_len)
            det.set_genome(genome)
            assert det.run(batch) == want, (prefetch, tile_len)
            assert det.stats().covered_positions == gp
            det.close()


@pytest.mark.parametrize("alphabet,bits", [((2, 11, 15, 22, 27, 33, 37, 40), 4), (tuple(range(2, 42)), 8), (None, 2)])
def test_quality_code_modes_give_the_oracle_records(oracle, tmp_path, monkeypatch, alphabet, bits):
    """Qualities travel host -> device as 2- / 4-bit codes when the tile's qualities fall into at most 4 / 16 classes the path can
    tell apart (ngsep_tile.qual_bits), as bytes otherwise: the records are the oracle's in every mode, and forcing bytes
    (NGSEP_QUAL_BITS=8) changes nothing."""
    kw = {} if alphabet is None else dict(qual_vals=alphabet)
    genome, batch, fa, sam = _inputs(tmp_path, [("chrA", 150000)], depth=25, seeds=(201, 202, 203), indel_rate=5e-4, **kw)
    want, _, gp, _ = ou.run_ssvd(oracle, fa, sam, args=("-maxBaseQS", "27"))
    monkeypatch.delenv("NGSEP_QUAL_BITS", raising=False)
    det = hostapi.Detector(maxBaseQS=27)
    det.set_genome(genome)
    assert det.pack(batch, 0)["quality"][0] == bits
    assert det.run(batch) == want
    monkeypatch.setenv("NGSEP_QUAL_BITS", "8")
    raw = hostapi.Detector(maxBaseQS=27)
    raw.set_genome(genome)
    assert raw.pack(batch, 0)["quality"][0] == 8
    assert raw.run(batch) == want
    assert len(_body(want)) > 100


def test_region_tiles_dealt_to_several_gpus_of_one_process(oracle, tmp_path):
    """SURVEY 8e / VERDICT r01 item 5: ONE detector process, admission + event pre-pass once, region tiles dealt in turn to the
    contexts of N GPUs (each with its own ring of page-locked buffers and tiles in flight), ordered merge on the host: the VCF
    of the N-GPU run is the 1-GPU run's and the oracle's.  Needs at least two GPUs (gpurun --gpus 2); skipped on one."""
    from ngsepcore_b200 import capi
    try:
        capi.Context(device=1).close()
    except capi.NgsepError:
        pytest.skip("one GPU only")
    genome, batch, fa, sam = _inputs(tmp_path, [("chrA", 400000), ("chrB", 130000)], depth=30, seeds=(211, 212, 213), indel_rate=5e-4)
    want, _, gp, _ = ou.run_ssvd(oracle, fa, sam)
    one = hostapi.Detector(tile_len=40000)
    one.set_genome(genome)
    assert one.run(batch) == want
    for n in (2, 4, 8):
        try:
            capi.Context(device=n - 1).close()
        except capi.NgsepError:
            break
        det = hostapi.Detector(tile_len=40000, n_devices=n)
        det.set_genome(genome)
        assert det.run(batch) == want
        assert det.stats().covered_positions == gp and det.stats().tiles == one.stats().tiles


@pytest.mark.parametrize("opts,det_kw", [((), {}), (("-embeddedSNVs", "-minQuality", "30"), dict(embeddedSNVs=True, minQuality=30))])
def test_known_strs_mode(oracle, tmp_path, opts, det_kw):
    """-knownSTRs (SSVD:906-913): the loci of the reference's yeast tutorial (first 400 rows of training/Saccharomyces_cerevisiae_STRs.txt,
    tests/golden/yeast_strs_head.txt) plus loci laid over planted indels, on a chrI-sized sequence.  Input loci are fixed realigner
    events (IRPL:89-99), their calls span the locus as TYPE=STR, no SNV is tried at a locus start, SNVs inside a locus are embedded
    (dropped, or TYPE=EMBEDDED with -embeddedSNVs).  Whole detector and the CLI against the oracle, byte for byte."""
    import subprocess
    from tests.test_host_pipeline_cpu import _write_str_loci
    contigs = [("chrI", 230218)]
    genome, batch, truth = simulate.simulate(contigs=contigs, depth=30, seeds=(241, 242, 243), indel_rate=1e-3, snv_rate=3e-3)
    fa, sam, strs = str(tmp_path / "ref.fa"), str(tmp_path / "reads.sam"), str(tmp_path / "strs.txt")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    _write_str_loci(strs, truth, contigs, np.random.default_rng(17), per_contig_random=0)
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "yeast_strs_head.txt")
    rows = sorted([l.split()[:3] for l in open(strs) if l.split()[0] == "chrI"] + [l.split()[:3] for l in open(golden)], key=lambda r: (int(r[1]), int(r[2])))
    open(strs, "w").writelines(f"{r[0]}\t{r[1]}\t{r[2]}\n" for r in rows)
    want, _, gp, _ = ou.run_ssvd(oracle, fa, sam, args=("-knownSTRs", strs) + opts)
    plain, _, _, _ = ou.run_ssvd(oracle, fa, sam, args=opts)
    body = _body(want)
    assert sum("TYPE=STR" in l for l in body) > 40 and len(body) != len(_body(plain))
    if "-embeddedSNVs" in opts:
        assert sum("TYPE=EMBEDDED" in l for l in body) > 20
    det = hostapi.Detector(knownSTRs=strs, **det_kw)
    det.set_genome(genome)
    assert det.run(batch) == want
    assert det.stats().covered_positions == gp
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "out")
    subprocess.check_call([os.path.join(root, "ngsepcore_b200", "ngsep_b200"), "SingleSampleVariantsDetector", "-i", sam, "-r", fa, "-o", out, "-knownSTRs", strs] + list(opts))
    assert open(out + ".vcf").read() == want


@pytest.mark.parametrize("depth,opts,det_kw", [
    (25, (), {}),
    (3, ("-minQuality", "10", "-maxBaseQS", "25"), dict(minQuality=10, maxBaseQS=25)),     # gaps in the cover, shallow columns, undecided calls
    (40, ("-ploidy", "4"), dict(ploidy=4)),               # the pool branch of genotypeVariantSample (SSVPL:366-377): sites come back as evidence columns
    (170, ("-ploidy", "6", "-minQuality", "30", "-maxAlnsPerStartPos", "0"), dict(ploidy=6, minQuality=30, maxAlnsPerStartPos=0)),   # ... deeper than one call slot
])
def test_known_variants_mode(oracle, tmp_path, depth, opts, det_kw):
    """-knownVariants (SSVD:897-905; the genotyping half of the reference's per-sample pipeline, training/runGenotyping): the detector
    genotypes the variants of an input VCF — SNV sites on the device (ngsep_tile_sites: every covering record counted, VD.genotypeSNV
    for the site's alleles, no screen and no hom-ref proof), the other variants by the host at their realigner event — and reports
    every one that meets a pileup, 0/0 and ./. included.  Input = the records of a discovery run plus hom-ref sites, multi-allelic
    SNVs, SNVs inside indel spans, variants sharing a position, input STRs, uncovered positions and records the reader drops.
    Whole detector (one tile, many tiles, synchronous) and the CLI against the oracle, byte for byte."""
    from tests.known_util import make_known_vcf
    contigs = [("chrA", 200000), ("chrB", 90000)] if depth < 100 else [("chrA", 60000), ("chrB", 20000)]
    genome, batch, truth = simulate.simulate(contigs=contigs, depth=depth, seeds=(341, 342, 343), indel_rate=1e-3, snv_rate=2e-3)
    fa, sam, kv = str(tmp_path / "ref.fa"), str(tmp_path / "reads.sam"), str(tmp_path / "known.vcf")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    discovered, _, _, _ = ou.run_ssvd(oracle, fa, sam, args=("-minQuality", "0"))
    assert make_known_vcf(kv, discovered, genome, contigs, np.random.default_rng(depth), extra_per_contig=400) > (1000 if depth < 100 else 500)
    want, _, gp, _ = ou.run_ssvd(oracle, fa, sam, args=("-knownVariants", kv) + opts)
    body = _body(want)
    gts = [l.split("\t")[9].split(":")[0] for l in body]
    assert len(body) > (1000 if depth < 100 else 500) and gts.count("0/0") > 50 and gts.count("./.") > 10 and sum("," in l.split("\t")[4] for l in body) > 50
    det = hostapi.Detector(knownVariants=kv, **det_kw)
    det.set_genome(genome)
    assert det.run(batch) == want
    tiled = hostapi.Detector(knownVariants=kv, tile_len=30000, **det_kw)
    tiled.set_genome(genome)
    assert tiled.run(batch) == want
    sync = hostapi.Detector(knownVariants=kv, tile_len=70000, synchronous=True, **det_kw)
    sync.set_genome(genome)
    assert sync.run(batch) == want
    parts = []
    for shard in range(2):
        d = hostapi.Detector(knownVariants=kv, tile_len=25000, shard=shard, n_shards=2, **det_kw)
        d.set_genome(genome)
        parts += _body(d.run(batch))
    key = {n: i for i, (n, _) in enumerate(contigs)}
    assert sorted(parts, key=lambda l: (key[l.split("\t")[0]], int(l.split("\t")[1]))) == sorted(body, key=lambda l: (key[l.split("\t")[0]], int(l.split("\t")[1])))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "out")
    subprocess.check_call([os.path.join(root, "ngsepcore_b200", "ngsep_b200"), "SingleSampleVariantsDetector", "-i", sam, "-r", fa, "-o", out, "-knownVariants", kv] + list(opts))
    assert open(out + ".vcf").read() == want


def test_compact_wire_layout_gives_the_full_layout_results(oracle, tmp_path, monkeypatch):
    """NGSEP_LAYOUT_COMPACT (the detectors' default): single-block records cross the bus without their CIGAR word and low-quality mask,
    the device rebuilds them at staging (one scan + one pass).  Same VCF bytes as the full layout and as the oracle; through the C ABI
    the same calls from the same sequence packed both ways, fewer bytes copied; a compact tile whose offsets do not add up is refused."""
    from ngsepcore_b200 import capi
    genome, batch, fa, sam = _inputs(tmp_path, [("chrA", 250000), ("chrB", 60000)], depth=30, seeds=(451, 452, 453), indel_rate=1e-3, n_frac=0.01)
    want, _, gp, _ = ou.run_ssvd(oracle, fa, sam)
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("NGSEP_WIRE_COMPACT", mode)
        det = hostapi.Detector(tile_len=70000)
        det.set_genome(genome)
        outs[mode] = det.run(batch)
        assert det.stats().covered_positions == gp
    assert outs["1"] == outs["0"] == want
    packs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("NGSEP_WIRE_COMPACT", mode)
        d = hostapi.Detector(host_only=True)
        d.set_genome(genome)
        packs[mode] = d.pack(batch, 0)
    res = {}
    with capi.Context() as ctx:
        ctx.ref_upload(0, genome["chrA"])
        for mode, t in packs.items():
            ctx.submit(0, t["first"], t["last"], t["hdr"], t["payload"], t["qual"], t["max_span"], t["max_back"], quality=t["quality"], layout=t["layout"])
            calls, _, st = ctx.fetch()
            res[mode] = (calls, st.h2d_bytes, st.covered_positions)
        assert np.array_equal(res["1"][0], res["0"][0]) and res["1"][2] == res["0"][2] and len(res["1"][0]) > 100
        assert res["1"][1] < 0.85 * res["0"][1]
        t = packs["1"]
        bad = t["hdr"].copy()
        bad["off"][len(bad) // 2] += 1                       # a single-block record whose compact payload is one word short
        ctx.submit(0, t["first"], t["last"], bad, t["payload"], t["qual"], t["max_span"], t["max_back"], quality=t["quality"], layout=t["layout"])
        with pytest.raises(capi.NgsepError):
            ctx.fetch()
        with pytest.raises(capi.NgsepError):                 # n_words (full) smaller than the stream it is expanded from
            ctx.submit(0, t["first"], t["last"], t["hdr"], t["payload"], t["qual"], t["max_span"], t["max_back"], quality=t["quality"], layout=(1, t["payload"].size - 5))


@pytest.mark.parametrize("ploidy,depth,opts,det_kw", [
    (4, 40, (), {}),
    (3, 25, ("-minQuality", "20", "-embeddedSNVs"), dict(minQuality=20, embeddedSNVs=True)),
    (8, 220, ("-maxBaseQS", "25", "-maxAlnsPerStartPos", "0"), dict(maxBaseQS=25, maxAlnsPerStartPos=0)),   # columns deeper than one call slot holds
])
def test_pools_and_polyploids(oracle, tmp_path, ploidy, depth, opts, det_kw):
    """-ploidy >= 3, the pool / polyploid algorithm of SingleSampleVariantPileupListener (SSVPL:240-254, 281-295, 402-503).  No hom-ref
    proof holds for it: every candidate of the screen reaches k_pool_columns, which counts the column, applies createSNVVariantPool's
    support rule and hands the surviving columns over with their counted bases in pileup order (NGSEP_CALL_POOL); the host builds the
    helpers of every allele-frequency hypothesis from them and genotypes the pool.  Whole detector (one tile, many tiles) and the CLI
    against the oracle, byte for byte; records are ADP-format CalledGenomicVariantImpl records without variant quality."""
    contigs = [("chrA", 160000), ("chrB", 50000)] if depth < 100 else [("chrA", 30000)]
    genome, batch, fa, sam = _inputs(tmp_path, contigs, depth=depth, seeds=(511 + ploidy, 512, 513), indel_rate=1e-3, snv_rate=3e-3)
    want, _, gp, _ = ou.run_ssvd(oracle, fa, sam, args=("-ploidy", str(ploidy)) + opts)
    body = _body(want)
    assert len(body) > (300 if depth < 100 else 60) and all(l.split("\t")[8] == "GT:PL:GQ:DP:ADP:ACN" and l.split("\t")[5] == "0" for l in body)
    det = hostapi.Detector(ploidy=ploidy, **det_kw)
    det.set_genome(genome)
    assert det.run(batch) == want
    assert det.stats().covered_positions == gp
    tiled = hostapi.Detector(ploidy=ploidy, tile_len=20000, **det_kw)
    tiled.set_genome(genome)
    assert tiled.run(batch) == want
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "out")
    subprocess.check_call([os.path.join(root, "ngsepcore_b200", "ngsep_b200"), "SingleSampleVariantsDetector", "-i", sam, "-r", fa, "-o", out, "-ploidy", str(ploidy)] + list(opts))
    assert open(out + ".vcf").read() == want
