#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals from `ncu --page source --csv --print-source cuda,sass` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
out = []
fname = ''
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': fname = r[1].split('/')[-1]
    if len(r) > 8 and r[0] == 'Line No':
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] not in ('', 'Line No'):
        d = dict(zip(hdr, r))
        try:
            out.append((int(d['Instructions Executed']), int(d['# Samples']), fname, r[0], r[1].strip()[:110], d))
        except ValueError:
            pass
tot_i = sum(o[0] for o in out); tot_s = sum(o[1] for o in out)
print('total warp-instr', tot_i, 'samples', tot_s)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(o[5][h]) for o in out if o[5][h].isdigit()) for h in stalls}
print('stalls:', {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for o in sorted(out, key=lambda o: -o[0])[:top]:
    print('%6.2f%% inst %5.2f%% smp  %s:%s  %s' % (100 * o[0] / tot_i, 100 * o[1] / max(1, tot_s), o[2], o[3], o[4]))
