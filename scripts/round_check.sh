#!/bin/bash
# What the driver runs at round end, in one go (one GPU): build, GPU tests, smoke, the bench line, the reference arm.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_check.log 2>&1 || { tail -20 gpurun_out/build_check.log; exit 1; }
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_check.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_check.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_check.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_check.log
( time timeout 900 python bench.py ) > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; echo "bench rc=$?"; cat gpurun_out/bench_check.json; tail -4 gpurun_out/bench_check.err
if [ "$1" == "--ref" ]; then
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref_check.json 2> gpurun_out/bench_ref_check.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_check.json; tail -4 gpurun_out/bench_ref_check.err
fi
