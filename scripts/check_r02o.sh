#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_check.log 2>&1 || { tail -20 gpurun_out/build_check.log; exit 1; }
timeout 150 python -m pytest tests/test_gpu_relative_allele_counts.py -x -q -m gpu > gpurun_out/pytest_subset.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_subset.log
