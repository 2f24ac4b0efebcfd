#!/bin/bash
# subset of the GPU tests that drive the product detector + bench line + the CLI from a BAM file (scratch_in/: made by tests/bam_util from a simulated 4 Mb x 30x contig)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_check.log 2>&1 || { tail -20 gpurun_out/build_check.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_detector_parity.py tests/test_gpu_sam_torture.py tests/test_gpu_multisample_parity.py "tests/test_gpu_baseline_configs.py::test_configs1_yeast_30x_with_indels_vs_oracle" -x -q -m gpu > gpurun_out/pytest_subset.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_subset.log
( time NGSEP_TIMING=1 timeout 900 python bench.py --no-cpu-baseline ) > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo "bench rc=$?"; python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_h.json").read().strip().splitlines()[-1])
print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "e2e", "pipeline")}))
P
grep -E "^run:|^device loop|real" gpurun_out/bench_h.err
if [ -f scratch_in/big.bam ]; then
  for i in 1 2 3; do ( time NGSEP_TIMING=1 ngsepcore_b200/ngsep_b200 SingleSampleVariantsDetector -i scratch_in/big.bam -r scratch_in/big.fa -o gpurun_out/cli_big ) 2>&1 | grep -E "load_al|^run:|device loop|real"; done
  grep -vc "^#" gpurun_out/cli_big.vcf; md5sum gpurun_out/cli_big.vcf; rm -f gpurun_out/cli_big.vcf
fi
