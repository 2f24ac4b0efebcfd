#!/bin/bash
# GPU check + launch-shape sweep on one GPU:
#   gpurun -- 'bash scripts/tune_gpu.sh [--no-tests] "T cap stages minblocks threads" ...'
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_tune.log 2>&1 || { tail -20 gpurun_out/build_tune.log; exit 1; }
if [ "$1" == "--no-tests" ]; then shift; else
  timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_tune.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_tune.log
fi
[ $# -eq 0 ] && set -- "640 3840 2 4 256"
for cfg in "$@"; do
  set -- $cfg
  tag="$1_$2_$3_$4_${5:-256}"
  NGSEP_WINDOW=$1 NGSEP_PAY_CAP_WORDS=$2 NGSEP_STAGES=$3 NGSEP_MIN_BLOCKS=$4 NGSEP_THREADS=${5:-256} timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/tune_$tag.json 2> gpurun_out/tune_$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    l=json.loads(open(f'gpurun_out/tune_{tag}.json').readline())
    print(f'cfg={tag}', 'ms/step %.3f'%l['ms_per_step'], 'value %.3e'%l['value'], 'frac %.4f'%l['roofline']['frac'], 'kernel_ms %.3f'%l['roofline']['kernel_ms_per_step'], 'e2e %.3e'%l['e2e']['value'], 'cand', l['config']['candidate_positions'], 'calls', l['config']['calls'], 'covered', l['config']['genotyped_positions_per_gpu'])
except Exception as e:
    print(f'cfg={tag} FAILED', e); print(open(f'gpurun_out/tune_{tag}.err').read()[-1500:])
PY
done
