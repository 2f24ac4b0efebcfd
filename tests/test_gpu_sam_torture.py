"""Record-level parity traps ("parity traps checklist") through the module CLI: a simulated SAM whose
records are mutated into the shapes htsjdk can hand over — soft / hard clips, '=' and 'X' operations, reference skips (N),
padding, secondary / NH-tagged / low-MAPQ / unmapped records, consecutive duplicate records, missing qualities, lower-case
bases, records without read group — called by `ngsep_b200 SingleSampleVariantsDetector` (host pipeline + CUDA kernels) and
by the oracle, VCF byte for byte, under several filter options."""
import os
import re
import subprocess

import numpy as np
import pytest

from ngsepcore_b200 import simulate
from tests import oracle_util as ou

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "ngsepcore_b200", "ngsep_b200")


def _mutate(lines, rng):
    out = []
    for line in lines:
        if line.startswith("@"):
            out.append(line)
            continue
        f = line.rstrip("\n").split("\t")
        flag, pos, cigar, seq, qual = int(f[1]), int(f[3]), f[5], f[9], f[10]
        ops = [(int(n), o) for n, o in re.findall(r"(\d+)([MIDNSHP=X])", cigar)]
        r = rng.random()
        if r < 0.06 and ops[0][1] == "M" and ops[0][0] > 20:           # soft clip at the start: POS moves, SEQ stays
            k = int(rng.integers(1, 9))
            ops = [(k, "S"), (ops[0][0] - k, "M")] + ops[1:]
            pos += k
        elif r < 0.10 and ops[-1][1] == "M" and ops[-1][0] > 20:        # soft clip at the end
            k = int(rng.integers(1, 9))
            ops = ops[:-1] + [(ops[-1][0] - k, "M"), (k, "S")]
        elif r < 0.13:                                                  # hard clips do not touch SEQ
            ops = [(5, "H")] + ops + [(3, "H")]
        elif r < 0.17 and ops[0][1] == "M" and ops[0][0] > 60:          # '=' and 'X' instead of M ('=' is M, X stays X)
            n = ops[0][0]
            ops = [(30, "="), (1, "X"), (n - 31, "=")] + ops[1:]
        elif r < 0.18 and len(ops) == 1 and ops[0][0] == 100:           # a reference skip: the tail aligns 150 bp further
            ops = [(60, "M"), (150, "N"), (40, "M")]
        elif r < 0.19 and len(ops) == 1 and ops[0][0] == 100:           # padding consumes nothing
            ops = [(50, "M"), (2, "P"), (50, "M")]
        elif r < 0.20 and len(ops) == 1:                                # equal neighbouring operations are merged (RA:1296-1312)
            ops = [(40, "M"), (60, "M")]
        tags = f[11:]
        r2 = rng.random()
        if r2 < 0.03: flag |= 0x100                                     # secondary
        elif r2 < 0.05: tags.append("NH:i:3")
        elif r2 < 0.08: tags.append("NH:i:1"); f[4] = "3"               # NH == 1 makes a low-MAPQ record unique (RAFR:284-291)
        elif r2 < 0.10: f[4] = "7"
        elif r2 < 0.11: flag |= 0x4
        elif r2 < 0.12: flag |= 0x400                                   # PCR duplicates are NOT dropped
        elif r2 < 0.13: qual = "*"
        elif r2 < 0.15: seq = seq[:20] + seq[20:60].lower() + seq[60:]
        elif r2 < 0.17: tags = [t for t in tags if not t.startswith("RG:Z:")]
        elif r2 < 0.18: seq = seq[:33] + "N" + seq[34:]
        elif r2 < 0.19: qual = qual[:70]                                # short QUAL is padded with '&' (RA:581-595)
        f[1], f[3], f[5], f[9], f[10] = str(flag), str(pos), "".join(f"{n}{o}" for n, o in ops), seq, qual
        rec = "\t".join(f[:11] + tags) + "\n"
        out.append(rec)
        if rng.random() < 0.02: out.append(rec)                          # the same record twice in a row (RAFR:292-306)
    # soft clips moved some starts: restore the coordinate order htsjdk insists on (stable)
    head = [l for l in out if l.startswith("@")]
    body = [l for l in out if not l.startswith("@")]
    order = {}
    for l in head:
        if l.startswith("@SQ"):
            order[l.split("\t")[1][3:]] = len(order)
    body.sort(key=lambda l: (order[l.split("\t")[2]], int(l.split("\t")[3])))
    return head + body


@pytest.mark.parametrize("opts", [
    (),
    ("-s", "-p"),
    ("-minMQ", "0", "-maxAlnsPerStartPos", "2", "-ignore5", "3", "-ignore3", "2"),
    ("-p", "-minQuality", "20", "-embeddedSNVs"),
])
def test_mutated_records(oracle, tmp_path, opts):
    genome, batch, _ = simulate.simulate(contigs=[("chrA", 70000), ("chrB", 30000)], depth=25, seeds=(131, 132, 133), indel_rate=1e-3, snv_rate=3e-3,
                                         n_read_groups=2, lower_frac=0.02)
    fa, sam0, sam = str(tmp_path / "r.fa"), str(tmp_path / "plain.sam"), str(tmp_path / "mutated.sam")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam0, batch)
    open(sam, "w").writelines(_mutate(open(sam0).readlines(), np.random.default_rng(7)))
    want, _, gp, _ = ou.run_ssvd(oracle, fa, sam, args=opts)
    out = str(tmp_path / "out")
    subprocess.check_call([CLI, "SingleSampleVariantsDetector", "-i", sam, "-r", fa, "-o", out] + list(opts), stderr=subprocess.DEVNULL)
    got = open(out + ".vcf").read()
    assert len([l for l in want.splitlines() if not l.startswith("#")]) > 150
    assert got == want


def test_mutated_records_multisample(oracle, tmp_path):
    """The same record shapes through `ngsep_b200 MultisampleVariantsDetector` (population of 6 samples, one input file)."""
    genome, batch, _ = simulate.simulate_population([("chrA", 50000), ("chrB", 20000)], n_samples=6, depth=10, mode="wgs", snv_rate=4e-3,
                                                    indel_rate=1e-3, seeds=(141, 142, 143), lower_frac=0.02)
    fa, sam0, sam = str(tmp_path / "r.fa"), str(tmp_path / "plain.sam"), str(tmp_path / "mutated.sam")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam0, batch)
    open(sam, "w").writelines(_mutate(open(sam0).readlines(), np.random.default_rng(11)))
    for opts in ((), ("-s", "-p", "-minQuality", "30"), ("-minMQ", "0", "-maxAlnsPerStartPos", "2", "-ignore5", "2")):
        want, _, _, _ = ou.run_msvd(oracle, fa, [sam], args=opts)
        out = str(tmp_path / "out.vcf")
        subprocess.check_call([CLI, "MultisampleVariantsDetector", "-r", fa, "-o", out] + list(opts) + [sam], stderr=subprocess.DEVNULL)
        assert len([l for l in want.splitlines() if not l.startswith("#")]) > 80
        assert open(out).read() == want


def test_reference_skip_longer_than_65535(oracle, tmp_path):
    """A spliced record whose reference span exceeds 65 535 bp (60M70000N40M): the window index packs the tile's largest span
    into 16 bits, so such a tile must hand its open columns over WITHOUT a record range and let the genotype kernel search
    with the full span (a capped span silently dropped the record from the columns of its far part).  The far
    40 bases of these records change DP / BSDP of the records called there, for one and for several samples."""
    genome, batch, _ = simulate.simulate(contigs=[("chrA", 160000)], depth=15, seeds=(151, 152, 153), indel_rate=0.0, snv_rate=5e-3)
    fa, sam0, sam = str(tmp_path / "r.fa"), str(tmp_path / "plain.sam"), str(tmp_path / "spliced.sam")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam0, batch)
    rng = np.random.default_rng(5)
    ref = genome["chrA"].tobytes().decode()
    out, n_spliced = [], 0
    for line in open(sam0):
        if not line.startswith("@"):
            f = line.split("\t")
            pos = int(f[3])
            if f[5] == "100M" and pos < 80000 and rng.random() < 0.02:
                far = pos + 60 + 70000 - 1                    # 0-based index of the first base after the skip
                tail = list(ref[far:far + 40].upper())
                for k in rng.integers(0, 40, size=2):          # two bases of the tail differ from the reference there
                    tail[k] = "ACGT"[("ACGT".index(tail[k]) + 1) % 4] if tail[k] in "ACGT" else "A"
                f[5] = "60M70000N40M"
                f[9] = f[9][:60] + "".join(tail)
                line = "\t".join(f)
                n_spliced += 1
        out.append(line)
    open(sam, "w").writelines(out)
    assert n_spliced > 100
    want, _, gp, _ = ou.run_ssvd(oracle, fa, sam)
    res = str(tmp_path / "out")
    subprocess.check_call([CLI, "SingleSampleVariantsDetector", "-i", sam, "-r", fa, "-o", res], stderr=subprocess.DEVNULL)
    body = [l for l in want.splitlines() if not l.startswith("#")]
    assert sum(1 for l in body if 70000 < int(l.split("\t")[1]) < 150100) > 100
    assert open(res + ".vcf").read() == want
