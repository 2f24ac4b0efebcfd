#!/bin/bash
# Round profiling recipe (run under gpurun, one GPU): bench line, launch list, one full ncu capture of the top kernel.
#   gpurun --timeout 1500 -- 'bash scripts/profile_gpu.sh r01'
# Follows /opt/skills/guides/B200_PROFILING.md: every ncu run comes directly after the same command exited 0 without ncu.
set -u
TAG=${1:-r01}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_${TAG}.log 2>&1 || { tail -20 gpurun_out/build_${TAG}.log; exit 1; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > gpurun_out/gpu_${TAG}.csv
# 1. the bench line (never taken under a profiler)
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || { tail -20 gpurun_out/bench_${TAG}.err; exit 1; }
cat gpurun_out/bench_${TAG}.json
# 2. launch list: device time of every launch (cold-cache, serialised: compare shares)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
# 3. the top kernel, full set, 3 launches on the largest tiles
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_pileup_windows -s 20 -c 3 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -20
