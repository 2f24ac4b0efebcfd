#!/bin/bash
# quick iteration on the multi-sample kernels: parity tests + the two configs[3] bench lines
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_ms.log 2>&1 || { tail -20 gpurun_out/build_ms.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_multisample_parity.py -x -q -m gpu > gpurun_out/pytest_ms.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ms.log
for tag in "$@"; do
  set -- $tag
  t=$1; shift
  env "$@" timeout 600 python scripts/bench_multisample.py --oracle-bp 0 > gpurun_out/msq_gbs_$t.json 2> gpurun_out/msq_gbs_$t.err
  env "$@" timeout 600 python scripts/bench_multisample.py --oracle-bp 0 --mode wgs --samples 24 --depth 8 --contigs 2 --contig-len 600000 > gpurun_out/msq_wgs_$t.json 2> gpurun_out/msq_wgs_$t.err
  python - $t <<'PY'
import json,sys
t=sys.argv[1]
for m in ('gbs','wgs'):
    try:
        l=json.loads(open(f'gpurun_out/msq_{m}_{t}.json').readline())
        print(t, m, 'kernels_ms %.3f'%l['kernels_ms'], 'records', l['vcf_records'], 'positions', l['genotyped_positions'])
    except Exception as e:
        print(t, m, 'FAILED', e, open(f'gpurun_out/msq_{m}_{t}.err').read()[-500:])
PY
done
