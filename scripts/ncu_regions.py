#!/usr/bin/env python
"""Instruction share per code region (function / phase marker) from `ncu --page source --csv --print-source cuda,sass`."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
srcfile = sys.argv[2]
hdr = None; fname = ''; per = {}; smp = {}
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': fname = r[1].split('/')[-1]
    if len(r) > 8 and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] not in ('', 'Line No'):
        d = dict(zip(hdr, r))
        try:
            k = (fname, int(r[0])); per[k] = per.get(k, 0) + int(d['Instructions Executed']); smp[k] = smp.get(k, 0) + int(d['# Samples'])
        except ValueError:
            pass
tot = sum(per.values()); tots = sum(smp.values())
src = open(srcfile).read().split('\n')
base = srcfile.split('/')[-1]
marks = [(i + 1, l) for i, l in enumerate(src) if re.match(r'^(template|__device__|__global__|struct|static|constexpr|size_t|void|int) ', l) or '// ---- ' in l]
def region(line):
    name = '?'
    for ln, l in marks:
        if ln <= line: name = f'{ln}:{l.strip()[:80]}'
    return name
agg = {}; aggs = {}
for (f, ln), v in per.items():
    key = region(ln) if f == base else f
    agg[key] = agg.get(key, 0) + v; aggs[key] = aggs.get(key, 0) + smp[(f, ln)]
print('total warp-instr', tot, 'samples', tots)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    print('%6.2f%% inst %6.2f%% smp  %s' % (100 * v / tot, 100 * aggs[k] / max(1, tots), k))
