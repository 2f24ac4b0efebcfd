#!/bin/bash
# GPU tests + bench line (with the host stage times of the product detector on stderr)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_check.log 2>&1 || { tail -20 gpurun_out/build_check.log; exit 1; }
lscpu | grep -E "Model name|^CPU\(s\)|avx2" | cut -c1-120
timeout 900 python -m pytest tests -x -q -m gpu --durations=25 > gpurun_out/pytest_check.log 2>&1; echo "pytest rc=$?"; tail -32 gpurun_out/pytest_check.log
( time NGSEP_TIMING=1 timeout 900 python bench.py ) > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; echo "bench rc=$?"; cat gpurun_out/bench_check.json; grep -E "^run:|real" gpurun_out/bench_check.err
