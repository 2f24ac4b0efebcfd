#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_check.log 2>&1 || { tail -20 gpurun_out/build_check.log; exit 1; }
NGSEP_TIMING=1 timeout 600 python scripts/pipeline_probe.py 2>&1 | grep -E "^run|^device loop|^context|wall"
if [ -f scratch_in/big.bam ]; then
  for i in 1 2 3; do ( time NGSEP_TIMING=1 ngsepcore_b200/ngsep_b200 SingleSampleVariantsDetector -i scratch_in/big.bam -r scratch_in/big.fa -o gpurun_out/cli_big ) 2>&1 | grep -E "load_al|^run:|device loop|real|^context"; done
  md5sum gpurun_out/cli_big.vcf; rm -f gpurun_out/cli_big.vcf
fi
