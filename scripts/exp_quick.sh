#!/bin/bash
# quick experiment: build, selected tests, bench lines with env variants:  exp_quick.sh "tag ENV=.." ...
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_exp.log 2>&1 || { tail gpurun_out/build_exp.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_snv_parity.py tests/test_gpu_detector_parity.py -x -q -m gpu > gpurun_out/pytest_exp.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_exp.log
for cfg in "$@"; do
  set -- $cfg
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-pipeline > gpurun_out/exp_$tag.json 2> gpurun_out/exp_$tag.err
  python - $tag <<'PY'
import json,sys
tag=sys.argv[1]
try:
    l=json.loads(open(f'gpurun_out/exp_{tag}.json').readline()); r=l['roofline']
    print(tag, 'step_ms %.4f'%l['ms_per_step'], 'pileup_ms %.4f'%r['kernel_ms_per_step'], 'all_kernels_ms %.4f'%r['both_kernels_ms_per_step'], 'frac %.4f'%r['frac'], 'open', l['config'].get('open_columns'), 'e2e %.3e'%l['e2e']['value'], 'e2e_ms %.2f'%l['e2e']['ms_per_step'], 'same', l['e2e'].get('same_calls_as_resident_run'))
except Exception as e:
    print(tag,'FAILED',e); print(open(f'gpurun_out/exp_{tag}.err').read()[-800:])
PY
done
