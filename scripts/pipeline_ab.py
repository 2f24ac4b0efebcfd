import sys, time, os
sys.path.insert(0, os.getcwd())
from ngsepcore_b200 import hostapi, simulate
genome, batch, _ = simulate.simulate(depth=30.0)
for mode in ("1", "0", "1", "0"):
    os.environ["NGSEP_WIRE_COMPACT"] = mode
    det = hostapi.Detector(); det.set_genome(genome)
    prev = None
    for it in range(3):
        t = time.perf_counter(); vcf = det.run(batch); wall = time.perf_counter() - t
        s = det.stats()
        cur = dict(prepass=s.prepass_s, pack=s.pack_s, device=s.device_s, total=s.total_s)
        d = {k: cur[k] - (prev[k] if prev else 0) for k in cur}
        prev = cur
    print("compact", mode, "wall %.3f" % wall, {k: round(v, 3) for k, v in d.items()}, flush=True)
    det.close()
