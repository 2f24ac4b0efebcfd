"""GPU parity, SNV path: CUDA kernels (through the C ABI) vs the CPU oracle on the same seeded alignments.

Inputs here carry no indels, so the indel realigner never fires and the numpy packer (no valid windows) is a
complete host side.  Bar: per-position counts and DP bit-exact, call set / alleles / GT / PL / GQ / QUAL identical.
"""
import os

import numpy as np
import pytest

from ngsepcore_b200 import calls as callfmt
from ngsepcore_b200 import capi, pypack, simulate
from tests import oracle_util as ou

pytestmark = pytest.mark.gpu


def _run_case(oracle, tmp_path, contigs, depth, sim_kw=None, ctx_kw=None, oracle_args=(), admit_kw=None, pack_kw=None, env=None, emit_all=True):
    sim_kw = dict(sim_kw or {})
    genome, batch, _ = simulate.simulate(contigs=contigs, depth=depth, indel_rate=0.0, **sim_kw)
    fa, sam = str(tmp_path / "ref.fa"), str(tmp_path / "reads.sam")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    vcf_text, ocounts, gp, _ = ou.run_ssvd(oracle, fa, sam, args=oracle_args, want_counts=True)
    orecs = ou.parse_vcf(vcf_text)
    idx_all = pypack.admit(batch, **(admit_kw or {}))
    got_lines, n_cov = [], 0
    ctx_kw = dict(ctx_kw or {})
    flags = ctx_kw.pop("flags", 0) | (capi.F_EMIT_ALL if emit_all else 0)
    with capi.Context(flags=flags, **ctx_kw) as ctx:
        opos = 0
        for ci, (name, length) in enumerate(contigs):
            ctx.ref_upload(ci, genome[name])
            idx = idx_all[batch.seq_idx[idx_all] == ci]
            hdr, pay, qual, span = pypack.pack_tile(batch, idx, **(pack_kw or {}))
            ctx.submit(ci, 1, length, hdr, pay, qual, span)
            calls, counts, stats = ctx.fetch(want_counts=emit_all)
            n_cov += stats.covered_positions
            if not emit_all:      # production path: screen + hom-ref proofs decide which columns are evaluated at all
                assert 0 < stats.open_columns <= length
                for c in calls:
                    got_lines.append(callfmt.snv_record_line(c, name))
                continue
            # counts parity: oracle dumps pos,A,C,G,T,DP for each position where discoverSNV ran, contig after contig
            oc = ocounts[opos:]
            nxt = np.nonzero(np.diff(oc[:, 0]) < 0)[0]
            end = (nxt[0] + 1) if nxt.size else oc.shape[0]
            oc = oc[:end]
            opos += end
            dense = np.zeros((length + 1, 5), np.int32)
            dense[oc[:, 0]] = oc[:, 1:]
            assert np.array_equal(counts, dense[1:]), f"counts differ on {name}: first at {np.nonzero((counts != dense[1:]).any(1))[0][:5] + 1}"
            for c in calls:
                got_lines.append(callfmt.snv_record_line(c, name))
    assert n_cov == gp  # genotyped positions = non-empty pileups (APG:497)
    want_lines = [r["line"] for r in orecs]
    assert len(got_lines) == len(want_lines)
    assert got_lines == want_lines
    return len(want_lines)


def test_small_two_contigs(oracle, tmp_path):
    n = _run_case(oracle, tmp_path, [("chrA", 200000), ("chrB", 100000)], 30)
    assert n > 200


def test_low_depth_and_n_runs(oracle, tmp_path):
    _run_case(oracle, tmp_path, [("c1", 120000)], 6, sim_kw=dict(n_frac=0.02, seeds=(11, 12, 13)))


def test_high_depth_multichunk(oracle, tmp_path):
    # 400x on a short contig: windows shrink and payload chunks roll over
    _run_case(oracle, tmp_path, [("hot", 6000)], 400, sim_kw=dict(seeds=(21, 22, 23)), admit_kw=dict(max_alns_per_start=0),
              oracle_args=("-maxAlnsPerStartPos", "0"))


def test_options_quality_ignore(oracle, tmp_path):
    _run_case(oracle, tmp_path, [("c1", 80000)], 25, sim_kw=dict(seeds=(31, 32, 33), lower_frac=0.05),
              ctx_kw=dict(max_base_qs=20, min_quality=20, het_rate=0.01, flags=capi.F_IGNORE_LOWER_REF),
              oracle_args=("-maxBaseQS", "20", "-minQuality", "20", "-h", "0.01", "-ignoreLowerCaseRef", "-ignore5", "3", "-ignore3", "7"),
              pack_kw=dict(ignore5=3, ignore3=7))


def test_panel_depth_1000x(oracle, tmp_path):
    """BASELINE configs[4] depth stress: a 1000x targeted panel with the per-start cap lifted (-maxAlnsPerStartPos 0, as a
    panel needs: APG:425); columns of ~1000 records stream through the stage chunk by chunk and the genotype kernel sums
    a thousand terms per likelihood in pileup order."""
    n = _run_case(oracle, tmp_path, [("panel", 3000)], 1000, sim_kw=dict(seeds=(41, 42, 43), snv_rate=4e-3), admit_kw=dict(max_alns_per_start=0),
                  oracle_args=("-maxAlnsPerStartPos", "0"))
    assert n >= 5


def test_quality_read_in_place_matches_copy(tmp_path):
    """A single-sample tile whose quality buffer is page-locked is not copied: k_genotype_columns reads the bytes it needs
    from the caller's buffer.  Same calls, bit for bit, as the copying path."""
    genome, batch, _ = simulate.simulate(contigs=[("c1", 150000)], depth=30, indel_rate=0.0, seeds=(61, 62, 63))
    idx = pypack.admit(batch)
    hdr, pay, qual, span = pypack.pack_tile(batch, idx)
    with capi.Context(flags=capi.F_QUALITY_IN_PLACE) as ctx:
        ctx.ref_upload(0, genome["c1"])
        ctx.submit(0, 1, 150000, hdr, pay, qual, span)
        a, _, st_a = ctx.fetch()
        assert st_a.quality_in_place == 0 and st_a.h2d_bytes == hdr.nbytes + pay.nbytes + qual.nbytes
        capi.host_register(qual)
        try:
            ctx.submit(0, 1, 150000, hdr, pay, qual, span)
            b, _, st_b = ctx.fetch()
        finally:
            capi.host_unregister(qual)
        assert st_b.quality_in_place == 1 and st_b.h2d_bytes == hdr.nbytes + pay.nbytes
        assert len(a) > 100 and len(a) == len(b)
        for name in a.dtype.names:      # field by field: the struct's padding bytes are not written by the device
            assert np.array_equal(a[name], b[name]), name
        assert st_a.covered_positions == st_b.covered_positions   # (open_columns may differ: which record flags a position first is a race)


# The same inputs WITHOUT the EMIT_ALL hook: the production path (screen, hom-ref proofs, open columns only).  The calls
# and the genotyped-position count must still be the oracle's.
_CASES = {
    "two_contigs": dict(contigs=[("chrA", 200000), ("chrB", 100000)], depth=30),
    "low_depth_n_runs": dict(contigs=[("c1", 120000)], depth=6, sim_kw=dict(n_frac=0.02, seeds=(11, 12, 13))),
    "high_depth_multichunk": dict(contigs=[("hot", 6000)], depth=400, sim_kw=dict(seeds=(21, 22, 23)), admit_kw=dict(max_alns_per_start=0),
                                  oracle_args=("-maxAlnsPerStartPos", "0")),
    "options_quality_ignore": dict(contigs=[("c1", 80000)], depth=25, sim_kw=dict(seeds=(31, 32, 33), lower_frac=0.05),
                                   ctx_kw=dict(max_base_qs=20, min_quality=20, het_rate=0.01, flags=capi.F_IGNORE_LOWER_REF),
                                   oracle_args=("-maxBaseQS", "20", "-minQuality", "20", "-h", "0.01", "-ignoreLowerCaseRef", "-ignore5", "3", "-ignore3", "7"),
                                   pack_kw=dict(ignore5=3, ignore3=7)),
    "panel_1000x": dict(contigs=[("panel", 3000)], depth=1000, sim_kw=dict(seeds=(41, 42, 43), snv_rate=4e-3), admit_kw=dict(max_alns_per_start=0),
                        oracle_args=("-maxAlnsPerStartPos", "0")),
    "haploid_prior": dict(contigs=[("c1", 100000)], depth=18, sim_kw=dict(seeds=(51, 52, 53), snv_rate=2e-3), ctx_kw=dict(het_rate=1e-6), oracle_args=("-h", "0.000001")),
}


@pytest.mark.parametrize("case", sorted(_CASES))
def test_production_path_without_emit_all(oracle, tmp_path, case):
    kw = dict(_CASES[case])
    n = _run_case(oracle, tmp_path, kw.pop("contigs"), kw.pop("depth"), emit_all=False, **kw)
    assert n >= 5
