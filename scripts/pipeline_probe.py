"""Where the whole-detector wall time goes (bench.py's informational `pipeline` leg): two runs on one detector."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngsepcore_b200 import hostapi, simulate
genome, batch, _ = simulate.simulate(depth=30.0)
det = hostapi.Detector()
det.set_genome(genome)
prev = None
for it in range(2):
    t = time.perf_counter(); vcf = det.run(batch); wall = time.perf_counter() - t
    s = det.stats()
    cur = dict(prepass=s.prepass_s, pack=s.pack_s, device=s.device_s, total=s.total_s)
    d = {k: cur[k] - (prev[k] if prev else 0) for k in cur}
    print("run", it, "wall %.3f" % wall, {k: round(v, 3) for k, v in d.items()}, "unaccounted in C++ %.3f" % (d["total"] - d["prepass"] - d["pack"] - d["device"]), "python overhead %.3f" % (wall - d["total"]))
    prev = cur
