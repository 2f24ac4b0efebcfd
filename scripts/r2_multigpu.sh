#!/bin/bash
# Multi-GPU evidence of round 2 (run under `gpurun --gpus 8`): strong scaling of BASELINE configs[2] / [4] with device-generated
# inputs at N = 2, 4, 8, the one-process multi-GPU detector test, and the default bench (configs[1], weak scaling) at N = 8.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_mg.log 2>&1
NG=$(nvidia-smi -L | wc -l)
echo "gpus: $NG"
(python -m pytest tests/test_gpu_detector_parity.py -m gpu -q -x -k "several_gpus" > gpurun_out/r2_mg_test.log 2>&1; tail -3 gpurun_out/r2_mg_test.log)
port=29500
for n in 2 4 8; do
  [ $n -le $NG ] || continue
  for c in c5 c3 panel; do
    port=$((port+1))
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --config $c --steps 5 --warmup 3 \
      > gpurun_out/r2_bench_${c}_${n}.json 2> gpurun_out/r2_bench_${c}_${n}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_${c}_${n}.json"))
    print("$c N=$n value %.3e ms %.3f frac %.3f gen_s %.2f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["host"]["generate_s"]))
except Exception as e:
    print("$c N=$n FAILED", e)
PY
  done
done
if [ 8 -le $NG ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-pipeline \
    > gpurun_out/r2_bench_c1_8.json 2> gpurun_out/r2_bench_c1_8.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c1_8.json')); print('c1 N=8 value %.3e e2e %.3e e2e_ms %.2f' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"
fi
