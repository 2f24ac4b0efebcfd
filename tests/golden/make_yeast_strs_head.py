"""tests/golden/yeast_strs_head.txt: the first 400 loci (all on chrI) of the reference's tutorial input for -knownSTRs,
training/Saccharomyces_cerevisiae_STRs.txt (a Tandem Repeats Finder table; the detector reads sequence, first, last from its first
three columns: genome/io/SimpleGenomicRegionFileHandler.java:57-80), cut to its first five columns.

    python tests/golden/make_yeast_strs_head.py /root/reference/training/Saccharomyces_cerevisiae_STRs.txt > tests/golden/yeast_strs_head.txt
"""
import sys

for k, line in enumerate(open(sys.argv[1])):
    if k >= 400:
        break
    print("\t".join(line.split()[:5]))
