"""RelativeAlleleCountsCalculator (discovery/RelativeAlleleCountsCalculator.java), one of the other PileupListener reducers of
SURVEY.md 8f-4, on the device: admission on the host (maxRD alignments per start position), the allele counts of EVERY position of
a region tile in ngsep_tile_allele_counts (shared-memory position x allele count tiles, the detector's own callable-base rules),
the three distributions of the report fed in position order — the report byte for byte against the oracle's restatement of the
listener on the reference's own walk: plain and mutated records, several option sets, the Python binding and the CLI."""
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
    # a sequence above and one below the 100 000 bp that earn a column of their own (:310), reads with N bases, indels, low MAPQ
    genome, batch, _ = simulate.simulate(contigs=[("chrA", 260000), ("chrB", 90000), ("chrC", 120000)], depth=25, seeds=(421, 422, 423), indel_rate=1e-3,
                                         snv_rate=4e-3, n_frac=0.01, low_mapq_frac=0.05)
    fa, sam = str(tmp_path / "r.fa"), str(tmp_path / "r.sam")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    for kw in (dict(), dict(m=5, M=2, q=30), dict(m=20, q=11, s=True), dict(m=1, q=0), dict(m=12, q=35)):
        want, _ = ou.run_relative_allele_counts(oracle, fa, sam, min_rd=kw.get("m", 10), max_rd=kw.get("M", 1000), min_bq=kw.get("q", 20), secondary=kw.get("s", False))
        got = hostapi.relative_allele_counts(batch, **kw)
        assert got == want, kw
    want, _ = ou.run_relative_allele_counts(oracle, fa, sam)
    lines = want.splitlines()
    assert lines[0] == "Distribution of allele proportions" and "Distribution of number of alleles" in lines
    assert lines[lines.index("Distribution of allele proportions per sequence") + 1] == "Proportion\tchrA\tchrC"
    count = int([l for l in lines if l.startswith("Count")][0].split("\t")[1])
    assert count > 400000 and float(lines[1].split("\t")[1]) < count        # positions with a second allele exist


def test_regions_to_filter_and_to_select(oracle, tmp_path):
    """-frs / -srs (RelativeAlleleCountsCalculator.java:131-161, 247-254): positions inside a region to filter are skipped, with regions
    to select only the positions inside one of them count.  Region files as SimpleGenomicRegionFileHandler reads them: tabs or single
    blanks, extra columns, overlapping and nested regions, a sequence the genome does not have, lines that do not parse (two blanks in
    a row, words for numbers, too few fields), a region with first > last.  Python binding and CLI, each against the oracle."""
    genome, batch, _ = simulate.simulate(contigs=[("chrA", 180000), ("chrB", 60000)], depth=22, seeds=(451, 452, 453), indel_rate=1e-3, snv_rate=5e-3)
    fa, sam, frs, srs = (str(tmp_path / n) for n in ("r.fa", "r.sam", "filter.txt", "select.txt"))
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    open(frs, "w").write("chrA\t1000\t30000\textra column\nchrA 20000 45000\nchrA\t21000\t22000\nchrB\t59000\t70000\nchrZ\t1\t100000\n"
                         "chrA  50000 60000\nchrA\tfifty\t60000\nchrB\t100\nchrB\t9000\t8000\n\nchrA\t+170000\t175000\r\n")
    open(srs, "w").write("chrA\t10000\t120000\nchrB\t1\t30000\nchrB\t25000\t31000\n")
    base, _ = ou.run_relative_allele_counts(oracle, fa, sam)
    count = lambda text: int([l for l in text.splitlines() if l.startswith("Count")][0].split("\t")[1])
    seen = set()
    for kw in (dict(frs=frs), dict(srs=srs), dict(frs=frs, srs=srs, m=5, q=25)):
        want, _ = ou.run_relative_allele_counts(oracle, fa, sam, min_rd=kw.get("m", 10), min_bq=kw.get("q", 20), frs=kw.get("frs"), srs=kw.get("srs"))
        assert hostapi.relative_allele_counts(batch, **kw) == want, kw
        assert 1000 < count(want) < count(base)
        seen.add(count(want))
        out = str(tmp_path / "rac.txt")
        opts = [x for k, v in kw.items() for x in ("-" + k, str(v))]
        subprocess.check_call([CLI, "RelativeAlleleCountsCalculator", "-i", sam, "-r", fa, "-o", out] + opts, stderr=subprocess.DEVNULL)
        assert open(out).read() == want, opts
    assert len(seen) == 3


def test_mutated_records_through_the_cli(oracle, tmp_path):
    from tests.test_gpu_sam_torture import _mutate
    genome, batch, _ = simulate.simulate(contigs=[("chrA", 120000), ("chrB", 30000)], depth=25, seeds=(431, 432, 433), indel_rate=1e-3, n_read_groups=2)
    fa, sam0, sam = (str(tmp_path / n) for n in ("r.fa", "plain.sam", "m.sam"))
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam0, batch)
    open(sam, "w").writelines(_mutate(open(sam0).readlines(), np.random.default_rng(5)))
    for opts in ((), ("-m", "4", "-q", "15", "-s"), ("-M", "1")):
        o = dict(zip(opts[::2], opts[1::2])) if "-s" not in opts else {"-m": "4", "-q": "15"}
        want, _ = ou.run_relative_allele_counts(oracle, fa, sam, min_rd=int(o.get("-m", 10)), max_rd=int(o.get("-M", 1000)), min_bq=int(o.get("-q", 20)),
                                                secondary="-s" in opts)
        out = str(tmp_path / "rac.txt")
        subprocess.check_call([CLI, "RelativeAlleleCountsCalculator", "-i", sam, "-r", fa, "-o", out] + list(opts), stderr=subprocess.DEVNULL)
        assert open(out).read() == want, opts
    assert subprocess.call([CLI, "RelativeAlleleCountsCalculator", "-r", fa], stderr=subprocess.DEVNULL) == 2       # the input file is required


def test_entry_point_directly_against_the_count_hook(oracle, tmp_path):
    """ngsep_tile_allele_counts on a packed sequence against the per-position A, C, G, T counts the detector's EMIT_ALL hook returns
    for the same tile (CountsHelper counts bases of quality > 3, capped at 30: with -q 4 and reads without N the two agree)."""
    genome, batch, _ = simulate.simulate(contigs=[("chrA", 50000)], depth=18, seeds=(441, 442, 443), indel_rate=2e-3)
    h, p, q, span, _ = hostapi.pack_sequence(batch, 0, max_alns_per_start=1000)
    quality = hostapi.pack_sequence.quality
    with capi.Context(flags=capi.F_EMIT_ALL) as ctx:
        ctx.ref_upload(0, genome["chrA"])
        out = ctx.tile_allele_counts(0, 1, 50000, h, p, q, span, min_rd=1, min_base_qs=4, quality=quality)
        ctx.submit(0, 1, 50000, h, p, q, span, quality=quality)
        _, counts, _ = ctx.fetch(want_counts=True)
    acgt = np.sort(counts[:, :4], axis=1)
    have = counts[:, 4] >= 1
    n_alleles = (counts[:, :4] > 0).sum(axis=1)
    ok = have & (n_alleles > 0)
    assert ok.sum() > 45000
    assert np.array_equal(out[ok, 0] & 0x0FFFFFFF, acgt[ok, 3]) and np.array_equal(out[ok, 1], acgt[ok, 2]) and np.array_equal(out[ok, 0] >> 28, n_alleles[ok])
    assert not out[~ok].any()
