#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_check.log 2>&1 || { tail -20 gpurun_out/build_check.log; exit 1; }
( time timeout 300 python bench.py ) > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_final.err
timeout 200 python scripts/bench_multisample.py --oracle-bp 0 --steps 5 > gpurun_out/bench_ms_gbs2.json 2> gpurun_out/bench_ms_gbs2.err; echo "bench gbs rc=$?"; cat gpurun_out/bench_ms_gbs2.json
