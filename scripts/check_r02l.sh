#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_check.log 2>&1 || { tail -20 gpurun_out/build_check.log; exit 1; }
timeout 300 python scripts/pipeline_ab.py NGSEP_PACK_HINT 2>&1 | grep -E "^NGSEP"
timeout 400 python -m pytest tests/test_gpu_detector_parity.py tests/test_gpu_multisample_parity.py -x -q -m gpu > gpurun_out/pytest_subset.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_subset.log
timeout 300 python scripts/bench_multisample.py --mode wgs --samples 24 --depth 8 --contigs 2 --contig-len 600000 --oracle-bp 0 --steps 5 > gpurun_out/bench_ms_wgs2.json 2> gpurun_out/bench_ms_wgs2.err; echo "bench wgs rc=$?"; cat gpurun_out/bench_ms_wgs2.json
