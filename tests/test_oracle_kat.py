"""Pins the oracle's SNV arithmetic against the reference's own demo call set."""
import gzip
import os

import pytest

from tests import oracle_util as ou

GOLD = os.path.join(os.path.dirname(__file__), "golden", "yeast_demo_kat.tsv.gz")


def _rows():
    with gzip.open(GOLD, "rt") as f:
        for line in f:
            if line[0] == "#":
                continue
            yield line.rstrip("\n").split("\t")


def test_survey_known_answers(oracle):
    # training/yeastDemo_...vcf.gz chrI:30309 C>T 0,0,0,21 DP 21 -> 1/1 PL 730,63,0 GQ 90
    r = ou.snv_column(oracle, [0, 0, 0, 21], ref="C", alt="T")
    assert (r["gt"], r["pl"], r["gq"], r["dp"], r["alt"]) == ((1, 1), (730, 63, 0), 90, 21, "T")
    # chrI:183063 G>C 0,12,60,0 -> 0/1 418,217,2086 GQ 255
    r = ou.snv_column(oracle, [0, 12, 60, 0], ref="G", alt="C")
    assert (r["gt"], r["pl"], r["gq"]) == ((0, 1), (418, 217, 2086), 255)
    # chrI:183214 G>A 16,0,30,0 -> 0/1 556,139,1043 GQ 255
    r = ou.snv_column(oracle, [16, 0, 30, 0], ref="G", alt="A")
    assert (r["gt"], r["pl"], r["gq"]) == ((0, 1), (556, 139, 1043), 255)
    # derived: ref A 15,15,0,0 -> het A/C PL 522,90,522 GQ 255 QUAL 255
    r = ou.snv_column(oracle, [15, 15, 0, 0], ref="A", alt="C")
    assert (r["gt"], r["pl"], r["gq"], r["qual"]) == ((0, 1), (522, 90, 522), 255, 255)
    # ref A 28,2,0,0 -> hom-ref wins with GQ 52 (record dropped upstream)
    r = ou.snv_column(oracle, [28, 2, 0, 0], ref="A", alt="C")
    assert r["gt"] == (0, 0) and r["gq"] == 52


def test_demo_vcf_tuples(oracle):
    """Every sample field of the reference's demo VCF is a (BSDP -> GT, PL, GQ) tuple.  Base qualities are not
    recorded, so only the all-q30 columns can match PL exactly (about a quarter of them); a wrong formula matches
    essentially none.  GT must match everywhere (it does not depend on a few sub-q30 bases)."""
    tot = exact = gt_ok = gq_ok = 0
    for c, pos, ref, alt, gt, pl, gq, dp, bsdp, acn in _rows():
        cnt = [int(x) for x in bsdp.split(",")]
        r = ou.snv_column(oracle, cnt, n_lowq=int(dp) - sum(cnt), ref=ref, alt=alt, discovery=False)
        mine_gt = f"{r['gt'][0]}/{r['gt'][1]}"
        tot += 1
        gt_ok += mine_gt == gt
        gq_ok += str(r["gq"]) == gq
        exact += mine_gt == gt and ",".join(map(str, r["pl"])) == pl and str(r["gq"]) == gq and r["dp"] == int(dp)
    assert tot > 40000
    assert gt_ok / tot > 0.999
    assert gq_ok / tot > 0.98
    assert exact / tot > 0.25


def test_tables_closed_form(oracle):
    import math
    assert oracle.oracle_table_gt(250, 30, 0) == pytest.approx(-0.00043451177, abs=1e-10)
    assert oracle.oracle_table_error(30, 4) == pytest.approx(-3.47712125, abs=1e-8)
    assert oracle.oracle_table_gt(250, 30, 4) == pytest.approx(-0.30131962, abs=1e-8)
    from ngsepcore_b200.capi import host_tables
    t = host_tables(0.001)
    for q in range(4, 31):  # the host tables handed to the device are bit-identical to the oracle's
        assert t[3 * q] == oracle.oracle_table_gt(250, q, 0)
        assert t[3 * q + 1] == oracle.oracle_table_error(q, 4)
        assert t[3 * q + 2] == oracle.oracle_table_gt(250, q, 4)
    assert t[93] == math.log10(0.999 / 4) and t[94] == math.log10(0.001 / 12)


def test_phred_and_fisher(oracle):
    assert oracle.oracle_phred(0.0) == 255
    assert oracle.oracle_phred(1e-30) == 255
    assert oracle.oracle_phred(1e-4) == 40
    assert oracle.oracle_phred(0.5) == 3
    assert oracle.oracle_fisher(3, 1, 1, 3) == pytest.approx(17.0 / 70.0, rel=1e-12)   # one-sided tail 16/70 + 1/70, FisherExactTest.java:80-97
    assert oracle.oracle_mononucleotide(b"AAAAAAAACGT") == 8
    assert oracle.oracle_mononucleotide(b"ACGTACGT") == 0
