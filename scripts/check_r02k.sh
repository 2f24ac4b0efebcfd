#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_check.log 2>&1 || { tail -20 gpurun_out/build_check.log; exit 1; }
timeout 600 python scripts/pipeline_ab.py NGSEP_PACK_HINT 2>&1 | grep -E "^NGSEP"
