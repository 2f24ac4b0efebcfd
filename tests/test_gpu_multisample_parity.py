"""GPU parity, MultisampleVariantsDetector: host pipeline + k_pileup_windows + k_genotype_columns_multi (C ABI) vs the
oracle's restatement of discovery/MultisampleVariantsDetector.java, VCF text byte for byte (BASELINE configs[3]).

Populations share planted SNV / indel sites with per-sample genotypes; samples may own several read groups and some
reads carry no RG tag (pooled discovery only, SSVPL:365).  Every record checks, per sample, the integer counts (BSDP, DP,
bit-exact), the genotype, GQ and the PL values (rounded -10 * fp64 log-likelihoods summed in the sample's pileup order)."""
import numpy as np
import pytest

from ngsepcore_b200 import hostapi, simulate
from tests import oracle_util as ou

pytestmark = pytest.mark.gpu


def _population(tmp_path, contigs, **kw):
    genome, batch, _ = simulate.simulate_population(contigs, **kw)
    fa, sam = str(tmp_path / "ref.fa"), str(tmp_path / "pop.sam")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    return genome, batch, fa, sam


def _body(text):
    return [l for l in text.splitlines() if not l.startswith("#")]


def _detector(batch, **kw):
    return hostapi.MultisampleDetector(list(zip(batch.rg_names, batch.rg_samples)), **kw)


def test_wgs_population_default_options(oracle, tmp_path):
    genome, batch, fa, sam = _population(tmp_path, [("chrA", 150000), ("chrB", 60000)], n_samples=12, depth=8, mode="wgs",
                                         rgs_per_sample=2, no_rg_frac=0.02, snv_rate=3e-3, indel_rate=5e-4, seeds=(21, 22, 23))
    want, _, gp, _ = ou.run_msvd(oracle, fa, [sam])
    det = _detector(batch)
    det.set_genome(genome)
    got = det.run(batch)
    st = det.stats()
    body = _body(want)
    assert len(body) > 400 and sum(1 for l in body if "TYPE=INDEL" in l or "TYPE=STR" in l) > 30
    assert any("," in l.split("\t")[4] and len(l.split("\t")[3]) == 1 for l in body)     # multi-allelic SNV sites are present
    assert got == want
    assert st.covered_positions == gp and st.kernel_launches > 0 and st.snv_calls > 300 and st.indel_calls > 30


def test_gbs_population_deep_identical_starts(oracle, tmp_path):
    """GBS-style tags: every sample's reads of a tag start at the same position (hundreds of records per start), run with
    -maxAlnsPerStartPos 100 as the reference's documentation asks for this data type (README.txt:588-596)."""
    genome, batch, fa, sam = _population(tmp_path, [("chrA", 120000)], n_samples=24, depth=10, mode="gbs", n_tags=400,
                                         snv_rate=5e-3, indel_rate=4e-4, seeds=(31, 32, 33))
    want, _, gp, _ = ou.run_msvd(oracle, fa, [sam], args=("-maxAlnsPerStartPos", "100"))
    det = _detector(batch, maxAlnsPerStartPos=100)
    det.set_genome(genome)
    got = det.run(batch)
    assert len(_body(want)) > 100
    assert got == want
    assert det.stats().covered_positions == gp


@pytest.mark.parametrize("seed,det_kw,oargs", [
    (4, dict(embeddedSNVs=True, minQuality=60, psp=True), ("-embeddedSNVs", "-minQuality", "60", "-psp")),
    (5, dict(h=0.01, maxBaseQS=25, ignore5=3, ignore3=4, minAlleleDepthFrequency=0.05), ("-h", "0.01", "-maxBaseQS", "25", "-ignore5", "3", "-ignore3", "4",
                                                                                        "--minAlleleDepthFrequency", "0.05")),
    (6, dict(ignoreLowerCaseRef=True, maxAlnsPerStartPos=2, querySeq="chrA", first=20000, last=70000),
     ("-ignoreLowerCaseRef", "-maxAlnsPerStartPos", "2", "-querySeq", "chrA", "-first", "20000", "-last", "70000")),
])
def test_option_matrix(oracle, tmp_path, seed, det_kw, oargs):
    genome, batch, fa, sam = _population(tmp_path, [("chrA", 90000), ("chrB", 30000)], n_samples=9, depth=10, mode="wgs", snv_rate=4e-3,
                                         indel_rate=1e-3, lower_frac=0.03, n_frac=0.01, seeds=(seed * 10, seed * 10 + 1, seed * 10 + 2))
    want, _, _, _ = ou.run_msvd(oracle, fa, [sam], args=oargs)
    det = _detector(batch, **det_kw)
    det.set_genome(genome)
    got = det.run(batch)
    assert len(_body(want)) > 60
    assert got == want


def test_region_tiles_and_shards_are_invariant(oracle, tmp_path):
    genome, batch, fa, sam = _population(tmp_path, [("chrA", 110000), ("chrB", 40000)], n_samples=6, depth=10, mode="wgs", snv_rate=3e-3,
                                         indel_rate=5e-4, seeds=(81, 82, 83))
    want, _, _, _ = ou.run_msvd(oracle, fa, [sam])
    for n_shards in (1, 3):
        lines = []
        for shard in range(n_shards):
            det = _detector(batch, tile_len=23000, shard=shard, n_shards=n_shards)
            det.set_genome(genome)
            lines += _body(det.run(batch, with_header=False))
        order = {n: i for i, n in enumerate(genome)}
        lines.sort(key=lambda l: (order[l.split("\t")[0]], int(l.split("\t")[1])))
        assert lines == _body(want)


def test_many_samples_more_than_one_cta_batch(oracle, tmp_path):
    """More samples than threads in a CTA of k_genotype_columns_multi (128): the sample axis is walked in batches."""
    genome, batch, fa, sam = _population(tmp_path, [("chrA", 20000)], n_samples=150, depth=3, mode="wgs", snv_rate=4e-3, indel_rate=0.0,
                                         seeds=(51, 52, 53))
    want, _, _, _ = ou.run_msvd(oracle, fa, [sam])
    det = _detector(batch)
    det.set_genome(genome)
    assert det.run(batch) == want
    assert len(_body(want)) > 40


def test_cli_several_input_files(oracle, tmp_path):
    """ngsep_b200 MultisampleVariantsDetector over one SAM per sample group: the k-way merge of APG:212-289 (order by
    sequence, first, last; ties to the lowest file index) happens on the host before admission."""
    import os
    import subprocess
    genome, batch, fa, sam = _population(tmp_path, [("chrA", 50000), ("chrB", 20000)], n_samples=6, depth=10, mode="wgs", snv_rate=4e-3,
                                         indel_rate=1.5e-3, seeds=(71, 72, 73))
    # some records carry a read group that no header declares: their own group for the per-start cap, no sample (pooled only)
    batch.rg_names = batch.rg_names + ["ghost"]
    ghost = np.random.default_rng(9).random(batch.n) < 0.04
    keep_rg = batch.rg.copy()
    batch.rg = np.where(ghost, len(batch.rg_names) - 1, batch.rg).astype(np.int16)
    simulate.write_sam(sam, batch)      # header: the six declared groups (write_sam zips ids with SM values); records: seven
    files = []
    for k in range(3):   # three inputs with two samples each
        idx = np.nonzero((keep_rg // 2) == k)[0]
        path = str(tmp_path / f"part{k}.sam")
        simulate.write_sam(path, batch.select(idx))
        files.append(path)
    want, _, _, _ = ou.run_msvd(oracle, fa, files, args=("-minQuality", "50"))
    assert len(_body(want)) > 60
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "ngsepcore_b200", "ngsep_b200")
    out = str(tmp_path / "multi.vcf")
    subprocess.check_call([cli, "MultisampleVariantsDetector", "-r", fa, "-o", out, "-minQuality", "50"] + files)
    assert open(out).read() == want
    # one input with every read group: same samples, no merge
    want1, _, _, _ = ou.run_msvd(oracle, fa, [sam])
    subprocess.check_call([cli, "MultisampleVariantsDetector", "-r", fa, "-o", out, sam])
    assert open(out).read() == want1
    assert subprocess.call([cli, "MultisampleVariantsDetector", "-o", out, sam], stderr=subprocess.DEVNULL) == 2      # MSVD:423


def test_c_abi_sample_axis_directly(oracle, tmp_path):
    """The sample axis through the C ABI itself (ctypes, numpy packer): pooled counts of EVERY position (EMIT_ALL hook)
    against the oracle's, and the device-resident batched launch (ngsep_resident_run / ngsep_resident_fetch_pop) against the
    submit / fetch_pop pipeline — same sites, same per-sample records."""
    from ngsepcore_b200 import capi, pypack
    contigs = [("chrA", 40000), ("chrB", 25000)]
    genome, batch, _ = simulate.simulate_population(contigs, n_samples=5, depth=9, mode="wgs", snv_rate=4e-3, indel_rate=0.0, no_rg_frac=0.05,
                                                    seeds=(91, 92, 93))
    fa, sam = str(tmp_path / "ref.fa"), str(tmp_path / "pop.sam")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    want, ocounts, gp, _ = ou.run_msvd(oracle, fa, [sam], want_counts=True)
    idx_all = pypack.admit(batch)
    groups = [[s] for s in range(5)]      # one read group per sample; records without RG carry -1 -> 0xFFFF
    tiles = []
    for ci, (name, length) in enumerate(contigs):
        idx = idx_all[batch.seq_idx[idx_all] == ci]
        hdr, pay, qual, span = pypack.pack_tile(batch, idx)
        tiles.append((ci, length, hdr, pay, qual, span, batch.rg[idx].astype(np.uint16)))     # int16 -1 -> 0xFFFF
    # (1) pooled counts of every position
    with capi.Context(flags=capi.F_EMIT_ALL, n_samples=5) as ctx:
        ctx.samples_set(groups, 5)
        opos = 0
        for ci, length, hdr, pay, qual, span, grp in tiles:
            ctx.ref_upload(ci, genome[contigs[ci][0]])
            ctx.submit(ci, 1, length, hdr, pay, qual, span, group=grp)
            sites, calls, counts, st = ctx.fetch_pop(want_counts=True)
            oc = ocounts[opos:]
            nxt = np.nonzero(np.diff(oc[:, 0]) < 0)[0]
            end = (nxt[0] + 1) if nxt.size else oc.shape[0]
            oc = oc[:end]
            opos += end
            dense = np.zeros((length + 1, 5), np.int32)
            dense[oc[:, 0]] = oc[:, 1:]
            assert np.array_equal(counts, dense[1:]), f"pooled counts differ on {contigs[ci][0]}"
    # (2) pipeline vs device-resident batch
    with capi.Context(n_samples=5) as ctx:
        ctx.samples_set(groups, 5)
        piped, slots = [], []
        for ci, length, hdr, pay, qual, span, grp in tiles:
            ctx.ref_upload(ci, genome[contigs[ci][0]])
            ctx.submit(ci, 1, length, hdr, pay, qual, span, group=grp)
            piped.append(ctx.fetch_pop()[:2])
        n_records = sum(len(s) for s, _ in piped)
        assert n_records == len(_body(want)) and n_records > 60      # no indels planted: every record is a device SNV site
        for ci, length, hdr, pay, qual, span, grp in tiles:
            slots.append(ctx.resident_add(ci, 1, length, hdr, pay, qual, span, group=grp))
        ctx.resident_run(-1)
        for (s_p, c_p), slot in zip(piped, slots):
            s_r, c_r, _, _ = ctx.fetch_pop(slot=slot)
            for name in ("pos", "count", "depth", "qs", "ref", "n_alleles", "allele", "n_alleles0"):
                assert np.array_equal(s_p[name], s_r[name]), name
            for name in ("count", "depth", "gq", "g0", "g1", "logcond"):
                assert np.array_equal(c_p[name], c_r[name]), name
        # the single-sample entry point refuses, and leaves the tile in flight
        ci, length, hdr, pay, qual, span, grp = tiles[0]
        ctx.submit(ci, 1, length, hdr, pay, qual, span, group=grp)
        with pytest.raises(capi.NgsepError):
            ctx.fetch()
        assert len(ctx.fetch_pop()[0]) == len(piped[0][0])


def test_population_tiles_dealt_to_several_gpus(oracle, tmp_path):
    """MultisampleVariantsDetector with region tiles dealt to the contexts of several GPUs of one process (every context gets the
    sample table, ngsep_samples_set): the VCF is the one-GPU run's and the oracle's.  Needs two GPUs; skipped on one."""
    from ngsepcore_b200 import capi
    try:
        capi.Context(device=1).close()
    except capi.NgsepError:
        pytest.skip("one GPU only")
    genome, batch, fa, sam = _population(tmp_path, [("chrA", 160000), ("chrB", 50000)], n_samples=8, depth=8, mode="wgs", snv_rate=3e-3, indel_rate=5e-4,
                                         seeds=(51, 52, 53))
    want, _, gp, _ = ou.run_msvd(oracle, fa, [sam])
    for n in (1, 2):
        det = _detector(batch, tile_len=30000, n_devices=n)
        det.set_genome(genome)
        assert det.run(batch) == want
        assert det.stats().covered_positions == gp
