#!/bin/bash
# Round profiling recipe (run under gpurun, one GPU): launch list + one full ncu capture of the batched pileup launch.
#   gpurun --timeout 1500 -- 'bash scripts/profile_gpu.sh r01b'
# Follows /opt/skills/guides/B200_PROFILING.md: every ncu run comes directly after the same command exited 0 without ncu.
set -u
TAG=${1:-r01}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_${TAG}.log 2>&1 || { tail -20 gpurun_out/build_${TAG}.log; exit 1; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > gpurun_out/gpu_${TAG}.csv
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pipeline"
# 1. launch list: device time of every launch (cold-cache, serialised: compare shares)
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
# 2. the batched pileup launch (launches 0-2 of that kernel are the warm-up runs over the whole genome), full set
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_pileup_windows -s 3 -c 1 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
# 3. the genotype kernel of the same batched run
ncu --set full --clock-control none --import-source on -k regex:k_genotype_columns -s 3 -c 1 -f -o gpurun_out/prof_geno_${TAG} $CMD > gpurun_out/ncu_geno_${TAG}.log 2>&1
ls -la gpurun_out | grep prof_
