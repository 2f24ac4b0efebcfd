#!/bin/bash
# What the driver runs at round end (build, GPU tests, smoke, bench line, reference arm) + the CLI from a BAM file when scratch_in/big.bam (+ big.fa) exists + the sample-axis benches
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_check.log 2>&1 || { tail -20 gpurun_out/build_check.log; exit 1; }
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_check.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_check.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_check.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_check.log
( time NGSEP_TIMING=1 timeout 900 python bench.py ) > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; echo "bench rc=$?"; cat gpurun_out/bench_check.json; grep -E "^run:|^device loop|^context|real" gpurun_out/bench_check.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref_check.json 2> gpurun_out/bench_ref_check.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_check.json; tail -4 gpurun_out/bench_ref_check.err
if [ -f scratch_in/big.bam ]; then
  for i in 1 2 3; do ( time NGSEP_TIMING=1 ngsepcore_b200/ngsep_b200 SingleSampleVariantsDetector -i scratch_in/big.bam -r scratch_in/big.fa -o gpurun_out/cli_big ) 2>&1 | grep -E "load_al|^run:|device loop|real|^context"; done
  md5sum gpurun_out/cli_big.vcf; rm -f gpurun_out/cli_big.vcf
fi
timeout 600 python scripts/bench_multisample.py > gpurun_out/bench_ms.json 2> gpurun_out/bench_ms.err; echo "bench ms rc=$?"; cat gpurun_out/bench_ms.json; tail -2 gpurun_out/bench_ms.err
timeout 600 python scripts/bench_multisample.py --mode wgs --samples 24 --depth 8 --contigs 2 --contig-len 600000 --oracle-bp 30000 > gpurun_out/bench_ms_wgs.json 2> gpurun_out/bench_ms_wgs.err; echo "bench wgs rc=$?"; cat gpurun_out/bench_ms_wgs.json; tail -2 gpurun_out/bench_ms_wgs.err
