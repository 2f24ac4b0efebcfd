#!/bin/bash
# Window-kernel launch shapes, one bench line each (gpurun_out/r2_sweep_<tag>.json)
run() { tag=$1; shift; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pipeline > gpurun_out/r2_sweep_$tag.json 2> gpurun_out/r2_sweep_$tag.err; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_sweep_$tag.json"))
    r = d["roofline"]
    print("$tag", "step_ms=%.4f win_ms=%.4f both_ms=%.4f open=%d e2e_ms=%.2f" % (d["ms_per_step"], r["kernel_ms_per_step"], r["both_kernels_ms_per_step"], d["config"]["open_columns"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print("$tag", "FAILED", e)
PY
}
run base NGSEP_X=0
run s2_640 NGSEP_STAGES=2 NGSEP_WINDOW=640 NGSEP_PAY_CAP_WORDS=3200
run s2_768 NGSEP_STAGES=2 NGSEP_WINDOW=768 NGSEP_PAY_CAP_WORDS=3840
run s2_512 NGSEP_STAGES=2 NGSEP_WINDOW=512 NGSEP_PAY_CAP_WORDS=2560 NGSEP_MIN_BLOCKS=5
run s1_1024 NGSEP_WINDOW=1024 NGSEP_PAY_CAP_WORDS=5120 NGSEP_MIN_BLOCKS=5
run geno6 NGSEP_GENO_MIN_BLOCKS=6
run geno_g8 NGSEP_GENO_BLOCKS_PER_SM=8
run geno_g32 NGSEP_GENO_BLOCKS_PER_SM=32
