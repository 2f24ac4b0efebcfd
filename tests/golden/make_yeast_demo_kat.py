"""Extracts known-answer tuples for the SNV genotype arithmetic from the reference's own demo call set.

Run in the build container only (needs /root/reference); the output tests/golden/yeast_demo_kat.tsv is committed.
Source: /root/reference/training/yeastDemo_ann_q40_s_fi_I2_noREP_noCNV.vcf.gz, produced by the reference's tutorial
pipeline with -maxBaseQS 30 (training/Tutorial.txt:149-179).  Every biallelic-SNV sample field
GT:PL:GQ:DP:BSDP is one tuple: given BSDP (and the bases all at q>=30), the formulas of CountsHelper /
VariantDiscoverySNVQAlgorithm / VCFFileWriter must give back PL and GQ.
"""
import gzip

SRC = "/root/reference/training/yeastDemo_ann_q40_s_fi_I2_noREP_noCNV.vcf.gz"
OUT = "tests/golden/yeast_demo_kat.tsv.gz"

rows = []
with gzip.open(SRC, "rt") as f:
    for line in f:
        if line.startswith("#"):
            continue
        t = line.rstrip("\n").split("\t")
        ref, alt, fmt = t[3], t[4], t[8]
        if len(ref) != 1 or len(alt) != 1 or fmt != "GT:PL:GQ:DP:BSDP:ACN":
            continue
        for s in t[9:]:
            gt, pl, gq, dp, bsdp, acn = s.split(":")
            if gt == "./.":
                continue
            rows.append((t[0], t[1], ref, alt, gt, pl, gq, dp, bsdp, acn))
with gzip.open(OUT, "wt") as o:
    o.write("#chrom\tpos\tref\talt\tGT\tPL\tGQ\tDP\tBSDP\tACN\n")
    for r in rows:
        o.write("\t".join(r) + "\n")
print(len(rows), "tuples")
