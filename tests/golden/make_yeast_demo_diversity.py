"""Extracts known-answer tuples for the population INFO fields (NS, AN, AFS, MAF) from the reference's own demo call set.

Run in the build container only (needs /root/reference); the output tests/golden/yeast_demo_diversity.tsv.gz is committed.
Source: /root/reference/training/yeastDemo_ann_q40_s_fi_I2_noREP_noCNV.vcf.gz — a two-sample file whose NS / AN / MAF / AFS were
written by the reference's VCF tools through the same code MultisampleVariantsDetector uses: variants/DiversityStatistics.java:
123-218 (allele counts from the samples' allele copy numbers) and main/io/ParseUtils.java:29 (DecimalFormat "##0.0#").
One row per distinct (number of alleles, per-sample GT and ACN, INFO values) combination, with its multiplicity.
"""
import collections
import gzip

SRC = "/root/reference/training/yeastDemo_ann_q40_s_fi_I2_noREP_noCNV.vcf.gz"
OUT = "tests/golden/yeast_demo_diversity.tsv.gz"

rows = collections.Counter()
with gzip.open(SRC, "rt") as f:
    for line in f:
        if line.startswith("#"):
            continue
        t = line.rstrip("\n").split("\t")
        info = dict(kv.split("=", 1) for kv in t[7].split(";") if "=" in kv)
        fmt = t[8].split(":")
        n_alleles = 1 + len(t[4].split(","))
        samples = []
        for s in t[9:]:
            d = dict(zip(fmt, s.split(":")))
            samples.append(d["GT"] + ":" + d["ACN"])
        rows[(str(n_alleles), "|".join(samples), info.get("NS", "."), info.get("AN", "."), info.get("AFS", "."), info.get("MAF", "."))] += 1
with gzip.open(OUT, "wt") as o:
    o.write("#n_alleles\tsamples(GT:ACN|...)\tNS\tAN\tAFS\tMAF\tmultiplicity\n")
    for k, n in sorted(rows.items()):
        o.write("\t".join(k) + f"\t{n}\n")
print(len(rows), "distinct tuples,", sum(rows.values()), "records")
