"""Runs the bench (device step only) once per value of an environment switch and prints step / window-kernel / genotype-kernel times.
    python scripts/sweep_env.py NGSEP_GENO_BLOCKS_PER_SM 4 8 16"""
import json, os, subprocess, sys
name, values = sys.argv[1], sys.argv[2:]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for v in values:
    env = dict(os.environ, **{name: v})
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "20", "--warmup", "3", "--no-cpu-baseline", "--no-pipeline"], env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    d = json.loads(out.strip().splitlines()[-1]); r = d["roofline"]
    print(name, v, "step %.4f ms  windows %.4f  genotype %.4f  e2e %.3f ms" % (d["ms_per_step"], r["k_pileup_windows"]["ms_per_step"], r["k_genotype_columns"]["ms_per_step"],
                                                                            d["e2e"]["ms_per_step"]), flush=True)
