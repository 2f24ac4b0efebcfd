"""Input VCFs for -knownVariants runs (test helper)."""
import numpy as np


def make_known_vcf(path, discovered_vcf, genome, contigs, rng, extra_per_contig=60):
    """A -knownVariants file: the records of a discovery run (all qualities) plus what a discovery run never holds — sites that are
    homozygous reference in the sample, multi-allelic SNVs, an SNV inside the span of an indel, two variants at one position, STR
    records, variants at positions no read covers, ID and fractional QUAL columns, records the reader drops (a structural variant,
    a reference-only gVCF site, a bad position, the wrong number of genotype columns)."""
    head = [l for l in discovered_vcf.splitlines() if l.startswith("#")]
    recs = []     # (contig index, pos, tie, line)
    order = {n: i for i, (n, _) in enumerate(contigs)}
    seqs = {n: bytes(np.asarray(genome[n], np.uint8)).decode() for n, _ in contigs}
    k = 0
    for l in discovered_vcf.splitlines():
        if l.startswith("#"):
            continue
        t = l.split("\t")
        if k % 7 == 0:
            t[2] = f"rs{k}"
        if k % 5 == 0:
            t[5] = f"{int(t[5])}.5" if k % 10 == 0 else "."
        recs.append((order[t[0]], int(t[1]), k, "\t".join(t)))
        k += 1
        if len(t[3]) > 3 and k % 3 == 0:      # an SNV inside the indel's span (hidden from the realigner by the indel, still genotyped)
            p = int(t[1]) + 2
            ref = t[3][2].upper()
            if ref in "ACGT":
                alt = "ACGT"[("ACGT".index(ref) + 1) % 4]
                recs.append((order[t[0]], p, k, f"{t[0]}\t{p}\t.\t{ref}\t{alt}\t50\t.\t.\tGT:PL:GQ:DP:BSDP:ACN\t0/1:0,0,0:50:0:0,0,0,0:1,1"))
                k += 1
    fmt = "GT:PL:GQ:DP:BSDP:ACN\t0/0:0,0,0:0:0:0,0,0,0:2,0"
    for name, length in contigs:
        s = seqs[name]
        for _ in range(extra_per_contig):
            p = int(rng.integers(2, length - 60))
            ref = s[p - 1].upper() if s else "A"
            if ref not in "ACGT":
                continue
            others = [b for b in "ACGT" if b != ref]
            kind = int(rng.integers(0, 6))
            if kind == 0:      # hom-ref site
                recs.append((order[name], p, k, f"{name}\t{p}\t.\t{ref}\t{others[0]}\t10\t.\t.\t{fmt}"))
            elif kind == 1:    # multi-allelic SNV with its type
                recs.append((order[name], p, k, f"{name}\t{p}\tm{k}\t{ref}\t{others[1]},{others[2]}\t99\t.\tTYPE=MULTISNV;NS=1\t{fmt}"))
            elif kind == 2:    # multi-allelic SNV, four alleles, no type
                recs.append((order[name], p, k, f"{name}\t{p}\t.\t{ref}\t{others[2]},{others[0]},{others[1]}\t.\t.\t.\t{fmt}"))
            elif kind == 3:    # an input STR over a stretch without indels, and a second variant at the same position
                span = int(rng.integers(4, 20))
                refa = s[p - 1:p - 1 + span].upper() if s else "A" * span
                if set(refa) <= set("ACGT"):
                    recs.append((order[name], p, k, f"{name}\t{p}\t.\t{refa}\t{refa[:2]}\t77\t.\tTYPE=STR\t{fmt}"))
                    k += 1
                    recs.append((order[name], p, k, f"{name}\t{p}\t.\t{ref}\t{others[0]}\t5\t.\t.\t{fmt}"))
            elif kind == 4:    # a deletion that is not in the sample
                refa = s[p - 1:p + 3].upper() if s else "AAAA"
                if set(refa) <= set("ACGT"):
                    recs.append((order[name], p, k, f"{name}\t{p}\t.\t{refa}\t{refa[0]}\t30\t.\tTYPE=INDEL\t{fmt}"))
            else:              # records the reader drops
                recs.append((order[name], p, k, f"{name}\t{p}\t.\t{ref}\t<DEL>\t30\t.\tSVTYPE=DEL;END={p + 500}\t{fmt}"))
                k += 1
                recs.append((order[name], p, k, f"{name}\t{p}\t.\t{ref}\t.\t30\t.\t.\t{fmt}"))
                k += 1
                recs.append((order[name], p, k, f"{name}\t-3\t.\t{ref}\t{others[0]}\t30\t.\t.\t{fmt}"))
                k += 1
                recs.append((order[name], p, k, f"{name}\t{p}\t.\t{ref}\t{others[0]}\t30\t.\t.\t{fmt}\t{fmt.split(chr(9))[1]}"))
            k += 1
    # the input need not be sorted: GenomicRegionSortedCollection sorts it (stable) — shuffle blocks of records
    recs.sort(key=lambda r: (r[0], r[1], r[2]))
    lines = [r[3] for r in recs]
    if len(lines) > 20:
        i = len(lines) // 3
        lines = lines[i:2 * i] + lines[:i] + lines[2 * i:]
    with open(path, "w") as fh:
        fh.write("\n".join(head) + "\n" + "\n".join(lines) + "\n")
    return len(lines)
