#!/bin/bash
# Multi-sample path on one GPU: parity tests, the configs[3] bench, an ncu launch list of it.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_ms.log 2>&1 || { tail -20 gpurun_out/build_ms.log; exit 1; }
if [ "$1" != "--no-tests" ]; then
timeout 900 python -m pytest tests/test_gpu_multisample_parity.py -x -q -m gpu > gpurun_out/pytest_ms.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_ms.log
fi
timeout 900 python scripts/bench_multisample.py > gpurun_out/bench_ms.json 2> gpurun_out/bench_ms.err; echo "bench rc=$?"; cat gpurun_out/bench_ms.json; tail -3 gpurun_out/bench_ms.err
timeout 900 python scripts/bench_multisample.py --mode wgs --samples 24 --depth 8 --contigs 2 --contig-len 600000 --oracle-bp 30000 > gpurun_out/bench_ms_wgs.json 2> gpurun_out/bench_ms_wgs.err; echo "bench wgs rc=$?"; cat gpurun_out/bench_ms_wgs.json; tail -3 gpurun_out/bench_ms_wgs.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/ms_launches.csv python scripts/bench_multisample.py --steps 1 --oracle-bp 0 > gpurun_out/ms_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/ms_launches.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    if r[ui] == 'ns': v /= 1e3
    elif r[ui] == 'ms': v *= 1e3
    tot[r[ki][:60]][0] += 1; tot[r[ki][:60]][1] += v
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]): print(f'{us:12.1f} us {n:5d} x {k}')
PY
