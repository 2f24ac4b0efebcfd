"""CPU checks of the oracle's MultisampleVariantsDetector restatement (oracle/ngsep_oracle_multi.cpp).

The reference holds no golden vector for the multi-sample logic (parity unpinned for this row), so the
checks are (a) facts of the Java runtime the restatement depends on (HashSet iteration order, DecimalFormat), (b) a
hand-built two-sample column whose per-sample fields must equal the PINNED single-sample arithmetic
(VD.genotypeSNV, tests/test_oracle_kat.py) and whose INFO follows DiversityStatistics by hand, and (c) the MSVD rules
as invariants over a simulated population."""
import ctypes as C

import numpy as np

from ngsepcore_b200 import simulate
from tests import oracle_util as ou


def test_java_runtime_facts(oracle):
    # DecimalFormat("##0.0#"), HALF_EVEN (main/io/ParseUtils.java:29)
    for v, s in [(0.0, "0.0"), (1.0, "1.0"), (0.5, "0.5"), (0.25, "0.25"), (0.125, "0.12"), (0.375, "0.38"), (1 / 3, "0.33"), (2 / 3, "0.67"), (0.1, "0.1")]:
        assert ou.english_fmt(oracle, v) == s
    # HashSet<String>: bucket = (h ^ h >>> 16) & 15 — single characters iterate in code order whatever the insertion order;
    # "RG0" (hash 81051 -> bucket 10) comes before "RG1" (81052 -> bucket 13)
    def order(keys):
        arr = (C.c_char_p * len(keys))(*[k.encode() for k in keys])
        out = (C.c_int * len(keys))()
        oracle.oracle_string_hashset_order(len(keys), arr, out)
        return [keys[i] for i in out]
    assert order(["b", "a", "c"]) == ["a", "b", "c"]
    assert order(["RG1", "RG0"]) == ["RG0", "RG1"]
    assert order(["Aa", "BB"]) == ["Aa", "BB"]   # equal hash codes (2112): bin keeps the insertion order


def _handmade(tmp_path):
    rng = np.random.default_rng(5)
    ref = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 400)].copy()
    ref[149] = ord("A")
    fa, sam = str(tmp_path / "h.fa"), str(tmp_path / "h.sam")
    simulate.write_fasta(fa, {"c": ref})
    lines = ["@HD\tVN:1.6\tSO:coordinate", "@SQ\tSN:c\tLN:400", "@RG\tID:g1\tSM:alpha", "@RG\tID:g2\tSM:beta", "@RG\tID:g3\tSM:gamma"]
    k = 0
    for start in range(91, 111):          # 20 start positions, one read per sample and start (cap of 5 per start and read group)
        for rg, alt in (("g1", True), ("g2", False)):
            if rg == "g1" and start % 2:  # alpha: 10 reads, all with C at 150
                continue
            seq = ref[start - 1:start + 79].copy()
            if alt:
                seq[150 - start] = ord("C")
            lines.append(f"r{k}\t0\tc\t{start}\t60\t80M\t*\t0\t0\t{seq.tobytes().decode()}\t{'?' * 80}\tRG:Z:{rg}")
            k += 1
    open(sam, "w").write("\n".join(lines) + "\n")
    return fa, sam


def test_handmade_two_sample_column(oracle, tmp_path):
    fa, sam = _handmade(tmp_path)
    text, _, gp, _ = ou.run_msvd(oracle, fa, [sam])
    rec = [l for l in text.splitlines() if not l.startswith("#")]
    hdr = [l for l in text.splitlines() if l.startswith("#CHROM")][0].split("\t")
    assert hdr[9:] == ["alpha", "beta", "gamma"]       # TreeMap order of the sample ids (MSVD:494,515)
    assert len(rec) == 1
    f = rec[0].split("\t")
    a = ou.snv_column(oracle, [0, 10, 0, 0], ref="A", alt="C", discovery=False)     # pinned VD.genotypeSNV arithmetic
    b = ou.snv_column(oracle, [20, 0, 0, 0], ref="A", alt="C", discovery=False)
    assert a["gt"] == (1, 1) and b["gt"] == (0, 0)
    assert f[:7] == ["c", "150", ".", "A", "C", str(a["gq"]), "."]                  # QUAL = best GQ of the non-hom-ref samples (MSVD:685-691)
    # DiversityStatistics: alpha 1/1 (copy numbers 0,2), beta 0/0 (2,0), gamma undecided -> NS=2, AN=2, AFS=2,2, OH=0, MAF=0.5
    assert f[7] == "NS=2;AN=2;AFS=2,2;OH=0.0;MAF=0.5"
    assert f[8] == "GT:PL:GQ:DP:BSDP:ACN"
    assert f[9] == f"1/1:{','.join(map(str, a['pl']))}:{a['gq']}:10:0,10,0,0:0,2"
    assert f[10] == f"0/0:{','.join(map(str, b['pl']))}:{b['gq']}:20:20,0,0,0:2,0"
    assert f[11] == "./.:0,0,0:0:0:0,0,0,0:2,0"                                    # no reads: undecided call without a report (VD:23-27)
    assert gp == 99   # covered positions 91..189


def test_population_invariants(oracle, tmp_path):
    contigs = [("c1", 40000), ("c2", 25000)]
    genome, batch, _ = simulate.simulate_population(contigs, n_samples=7, depth=9, mode="wgs", rgs_per_sample=2, no_rg_frac=0.03,
                                                    snv_rate=4e-3, indel_rate=6e-4, seeds=(11, 12, 13))
    fa, sam = str(tmp_path / "p.fa"), str(tmp_path / "p.sam")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    text, counts, gp, _ = ou.run_msvd(oracle, fa, [sam], want_counts=True)
    recs = [l.split("\t") for l in text.splitlines() if not l.startswith("#")]
    assert len(recs) > 150 and sum(1 for r in recs if "TYPE=INDEL" in r[7] or "TYPE=STR" in r[7]) > 10
    last = {}
    for r in recs:
        fmt = r[8].split(":")
        smp = [dict(zip(fmt, s.split(":"))) for s in r[9:]]
        assert len(smp) == 7
        info = dict(kv.split("=") for kv in r[7].split(";"))
        decided = [s for s in smp if s["GT"] != "./."]
        assert int(info["NS"]) == len(decided)
        assert all(int(s["GQ"]) >= 40 for s in decided) and all(int(s["GQ"]) == 0 for s in smp if s["GT"] == "./.")   # SSVPL:388
        non_ref = [int(s["GQ"]) for s in decided if s["GT"] != "0/0"]
        assert non_ref and int(r[5]) == max(non_ref) and int(r[5]) >= 40                                              # MSVD:533,685-691
        called = {a for s in decided for a in s["GT"].split("/")}
        assert int(info["AN"]) == len(called)
        n_alleles = 1 + len(r[4].split(","))
        assert ("MAF" in info) == (n_alleles == 2)
        afs = [int(x) for x in info["AFS"].split(",")]
        assert len(afs) == n_alleles and sum(afs) == 2 * len(decided)
        assert int(r[1]) > last.get(r[0], 0)          # one record per position, ascending (indel suppression, MSVD:536-538)
        last[r[0]] = int(r[1]) if len(r[3]) == 1 else int(r[1]) + len(r[3]) - 1
    assert gp > 60000 and counts.shape[0] > 60000


def _expected_impl_sample(oracle, counts, alleles, h=0.001):
    """Independent numpy restatement of CH.updateCounts + CH.getPosteriorProbabilities + VD.genotypeSNV (multi-allelic branch)
    for a sample whose bases all have quality 30: returns (GT, PL list, GQ)."""
    import math
    a30, b30, c30 = oracle.oracle_table_gt(250, 30, 0), oracle.oracle_table_error(30, 4), oracle.oracle_table_gt(250, 30, 4)
    L = np.zeros((4, 4))
    for b, n in enumerate(counts):                      # the order of summation does not matter here: equal terms per base
        for _ in range(n):
            for x in range(4):
                L[x][x] += a30 if x == b else b30
                for y in range(4):
                    if x != y:
                        L[x][y] += c30 if (x == b or y == b) else b30
    ev = []
    for i in range(4):
        ev.append(L[i][i] + math.log10((1 - h) / 4))
        for j in range(4):
            if j != i:
                ev.append(L[i][j] + math.log10(h / 12))
    mx = max(ev)
    e = [0.0 if v - mx < -20 else 10 ** (v - mx) for v in ev]
    tot = sum(e)
    P = np.zeros((4, 4))
    k = 0
    for i in range(4):
        P[i][i] = e[k] / tot; k += 1
        for j in range(4):
            if j != i:
                P[i][j] = e[k] / tot; k += 1
    idx = ["ACGT".index(a) for a in alleles]
    best, bi, bj = P[idx[0]][idx[0]], 0, 0
    for i in range(len(idx)):
        for j in range(i, len(idx)):
            gp = P[idx[i]][idx[j]] + (P[idx[j]][idx[i]] if i != j else 0.0)
            if gp > best + 0.01:
                best, bi, bj = gp, i, j
    maxp = P[idx[bi]][idx[bj]] + (P[idx[bj]][idx[bi]] if bi != bj else 0.0)
    gq = 255 if (1 - maxp) <= 0 or -10 * math.log10(1 - maxp) > 255 else int(math.floor(-10 * math.log10(1 - maxp) + 0.5))
    pl = [int(math.floor(-10 * L[idx[i]][idx[j]] + 0.5)) for j in range(len(idx)) for i in range(j + 1)]
    return (bi, bj), pl, gq


def test_handmade_multiallelic_site(oracle, tmp_path):
    """Three samples, three alleles (A reference, C and G): the record is a MULTISNV with CalledGenomicVariantImpl genotypes
    (VD:61-95): double-precision PL of the allele sub-matrix, BSDP from all counts, copy numbers from the counts, AFS over three
    alleles and no MAF (not biallelic)."""
    rng = np.random.default_rng(6)
    ref = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 400)].copy()
    ref[149] = ord("A")
    fa, sam = str(tmp_path / "h.fa"), str(tmp_path / "h.sam")
    simulate.write_fasta(fa, {"c": ref})
    lines = ["@HD\tVN:1.6\tSO:coordinate", "@SQ\tSN:c\tLN:400", "@RG\tID:g1\tSM:s1", "@RG\tID:g2\tSM:s2", "@RG\tID:g3\tSM:s3"]
    k = 0
    for start in range(91, 111):
        for rg, base, every in (("g1", "C", 2), ("g2", "G", 2), ("g3", None, 1)):
            if start % every:
                continue
            seq = ref[start - 1:start + 79].copy()
            if base:
                seq[150 - start] = ord(base)
            lines.append(f"r{k}\t0\tc\t{start}\t60\t80M\t*\t0\t0\t{seq.tobytes().decode()}\t{'?' * 80}\tRG:Z:{rg}")
            k += 1
    open(sam, "w").write("\n".join(lines) + "\n")
    text, _, _, _ = ou.run_msvd(oracle, fa, [sam])
    rec = [l for l in text.splitlines() if not l.startswith("#")]
    assert len(rec) == 1
    f = rec[0].split("\t")
    exp = [_expected_impl_sample(oracle, c, "ACG") for c in ([0, 10, 0, 0], [0, 0, 10, 0], [20, 0, 0, 0])]
    assert [e[0] for e in exp] == [(1, 1), (2, 2), (0, 0)]
    qual = max(exp[0][2], exp[1][2])
    assert f[:7] == ["c", "150", ".", "A", "C,G", str(qual), "."]
    assert f[7] == "NS=3;AN=3;AFS=2,2,2;OH=0.0;TYPE=MULTISNV"
    assert f[8] == "GT:PL:GQ:DP:BSDP:ACN"
    want = []
    for (gt, pl, gq), cnt, acn in zip(exp, ([0, 10, 0, 0], [0, 0, 10, 0], [20, 0, 0, 0]), ("0,2,0", "0,0,2", "2,0,0")):
        want.append(f"{gt[0]}/{gt[1]}:{','.join(map(str, pl))}:{gq}:{sum(cnt)}:{','.join(map(str, cnt))}:{acn}")
    assert f[9:] == want


def test_population_info_against_reference_demo_vcf(oracle):
    """PINNED: the NS / AN / AFS / MAF values of the reference's own two-sample demo VCF (20 456 records, 7 distinct genotype
    patterns; tests/golden/make_yeast_demo_diversity.py) are what the restated DiversityStatistics + DecimalFormat give for
    the file's genotypes and allele copy numbers."""
    import gzip
    import os
    gold = os.path.join(os.path.dirname(__file__), "golden", "yeast_demo_diversity.tsv.gz")
    n_rec = 0
    with gzip.open(gold, "rt") as f:
        for line in f:
            if line.startswith("#"):
                continue
            n_alleles, samples, ns, an, afs, maf, mult = line.rstrip("\n").split("\t")
            called, cns = [], []
            for s in samples.split("|"):
                gt, acn = s.split(":")
                called.append((-1, -1) if gt == "./." else tuple(int(x) for x in gt.split("/")))
                cns.append([int(x) for x in acn.split(",")])
            info = dict(kv.split("=") for kv in ou.diversity_info(oracle, int(n_alleles), called, cns).split(";"))
            assert (info["NS"], info["AN"], info["AFS"], info.get("MAF", ".")) == (ns, an, afs, maf), line
            n_rec += int(mult)
    assert n_rec == 20456
