#!/bin/bash
# Window-kernel launch shapes, one bench line each (gpurun_out/r2_sweep_<tag>.json):  bash scripts/r2_sweep.sh tag1 tag2 ...
run() { tag=$1; shift; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pipeline > gpurun_out/r2_sweep_$tag.json 2> gpurun_out/r2_sweep_$tag.err; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_sweep_$tag.json"))
    r = d["roofline"]
    print("$tag", "step_ms=%.4f win_ms=%.4f both_ms=%.4f open=%d e2e_ms=%.2f frac=%.3f" % (d["ms_per_step"], r["k_pileup_windows"]["ms_per_step"], r["kernels_ms_per_step"], d["config"]["open_columns"], d["e2e"]["ms_per_step"], r["frac"]))
except Exception as e:
    print("$tag", "FAILED", e)
PY
}
for t in "$@"; do
case $t in
base) run base NGSEP_X=0 ;;
s2_640) run s2_640 NGSEP_STAGES=2 NGSEP_WINDOW=640 NGSEP_PAY_CAP_WORDS=3200 ;;
s2_768) run s2_768 NGSEP_STAGES=2 NGSEP_WINDOW=768 NGSEP_PAY_CAP_WORDS=3840 ;;
s2_1024) run s2_1024 NGSEP_STAGES=2 NGSEP_WINDOW=1024 NGSEP_PAY_CAP_WORDS=5120 NGSEP_MIN_BLOCKS=3 ;;
s1_1024) run s1_1024 NGSEP_WINDOW=1024 NGSEP_PAY_CAP_WORDS=5120 NGSEP_MIN_BLOCKS=5 ;;
s1_896) run s1_896 NGSEP_WINDOW=896 NGSEP_PAY_CAP_WORDS=4480 NGSEP_MIN_BLOCKS=5 ;;
s1_768_6) run s1_768_6 NGSEP_WINDOW=768 NGSEP_PAY_CAP_WORDS=3840 NGSEP_MIN_BLOCKS=6 ;;
s1_896_6) run s1_896_6 NGSEP_WINDOW=896 NGSEP_PAY_CAP_WORDS=4480 NGSEP_MIN_BLOCKS=6 ;;
s2_768_3) run s2_768_3 NGSEP_STAGES=2 NGSEP_WINDOW=768 NGSEP_PAY_CAP_WORDS=3840 NGSEP_MIN_BLOCKS=4 ;;
s1_1536) run s1_1536 NGSEP_WINDOW=1536 NGSEP_PAY_CAP_WORDS=7680 NGSEP_MIN_BLOCKS=3 ;;
t128) run t128 NGSEP_THREADS=128 NGSEP_WINDOW=640 NGSEP_PAY_CAP_WORDS=3200 ;;
geno6) run geno6 NGSEP_GENO_MIN_BLOCKS=6 ;;
entries) run entries NGSEP_ENTRIES=1 ;;
nofence) run nofence NGSEP_PILEUP_LIB=$PWD/ngsepcore_b200/_lib_nofence.so ;;
warp0) run warp0 NGSEP_PILEUP_LIB=$PWD/ngsepcore_b200/_lib_warp0.so ;;
g16) run g16 NGSEP_GENO_BLOCKS_PER_SM=16 ;;
s1_640) run s1_640 NGSEP_WINDOW=640 NGSEP_PAY_CAP_WORDS=3200 NGSEP_MIN_BLOCKS=5 ;;
t128_640) run t128_640 NGSEP_THREADS=128 NGSEP_WINDOW=640 NGSEP_PAY_CAP_WORDS=3200 NGSEP_MIN_BLOCKS=10 ;;
t128_1280) run t128_1280 NGSEP_THREADS=128 NGSEP_MIN_BLOCKS=8 ;;
esac
done
