"""CoverageStats (discovery/CoverageStatisticsCalculator.java), one of the other PileupListener reducers of SURVEY.md 8f-4, on the
device: admission on the host (secondary alignments processed, 100 alignments per start position), read-depth histograms per
region tile in ngsep_coverage_histogram, the printed table byte for byte against the oracle's restatement of the listener on the
reference's own walk — plain and mutated records (clips, N skips, secondary / NH / low-MAPQ / unmapped / duplicate records),
several -minMQ values, through the Python binding and through the CLI."""
import os
import subprocess

import numpy as np
import pytest

from ngsepcore_b200 import capi, hostapi, simulate
from tests import oracle_util as ou

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "ngsepcore_b200", "ngsep_b200")


def test_plain_alignments_match_the_oracle(oracle, tmp_path):
    genome, batch, _ = simulate.simulate(contigs=[("chrA", 300000), ("chrB", 90000)], depth=35, seeds=(221, 222, 223), indel_rate=1e-3, low_mapq_frac=0.1)
    fa, sam = str(tmp_path / "r.fa"), str(tmp_path / "r.sam")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    for mq in (20, 0, 30):
        want, _ = ou.run_coverage_stats(oracle, fa, sam, min_mq=mq)
        got = hostapi.coverage_stats(batch, minMQ=mq)
        assert got == want
    rows = [l.split("\t") for l in want.splitlines()]
    assert len(rows) == 300 and rows[-1][0] == "More"
    assert sum(int(r[1]) for r in rows) > 380000 and any(int(r[1]) != int(r[2]) for r in rows)     # low-MAPQ reads are not unique


def test_mutated_records_through_the_cli(oracle, tmp_path):
    from tests.test_gpu_sam_torture import _mutate
    genome, batch, _ = simulate.simulate(contigs=[("chrA", 80000), ("chrB", 30000)], depth=25, seeds=(231, 232, 233), indel_rate=1e-3, n_read_groups=2)
    fa, sam0, sam = (str(tmp_path / n) for n in ("r.fa", "plain.sam", "m.sam"))
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam0, batch)
    open(sam, "w").writelines(_mutate(open(sam0).readlines(), np.random.default_rng(3)))
    for mq in (20, 5):
        want, _ = ou.run_coverage_stats(oracle, fa, sam, min_mq=mq)
        out = str(tmp_path / "cov.txt")
        subprocess.check_call([CLI, "CoverageStats", "-i", sam, "-r", fa, "-o", out, "-minMQ", str(mq)], stderr=subprocess.DEVNULL)
        assert open(out).read() == want
    assert subprocess.call([CLI, "CoverageStats", "-r", fa], stderr=subprocess.DEVNULL) == 2       # the input file is required


def test_histogram_entry_point_directly():
    """ngsep_coverage_histogram on hand-made intervals: depth d = reads with first <= p <= last, clipped to the tile; [max] = more."""
    first = np.array([1, 1, 5, 5, 5, 20, 90], np.int32)
    last = np.array([10, 4, 9, 30, 6, 25, 140], np.int32)
    unique = np.array([1, 0, 1, 1, 0, 1, 1], np.uint8)
    with capi.Context() as ctx:
        ha, hu = ctx.coverage_histogram(1, 100, first, last, unique, max_coverage=3)
    depth = np.zeros(101, int); du = np.zeros(101, int)
    for f, l, u in zip(first, last, unique):
        depth[f:min(l, 100) + 1] += 1
        du[f:min(l, 100) + 1] += int(u)
    want_a = np.bincount(np.minimum(depth[1:], 3), minlength=4)
    want_u = np.bincount(np.minimum(du[1:], 3), minlength=4)
    assert ha.tolist() == want_a.tolist() and hu.tolist() == want_u.tolist()
