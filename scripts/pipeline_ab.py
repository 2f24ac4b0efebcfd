"""A/B of a switch read from the environment (NAME on the command line, values 1 / 0 alternating) on the product detector: wall time and
stage times of runs 2 .. 6 of a detector object (the first run creates the context and sizes the buffers)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngsepcore_b200 import hostapi, simulate
genome, batch, _ = simulate.simulate(depth=30.0)
name = sys.argv[1] if len(sys.argv) > 1 else "NGSEP_WIRE_COMPACT"
for mode in ("1", "0", "1", "0"):
    os.environ[name] = mode
    det = hostapi.Detector(); det.set_genome(genome)
    prev, walls, stages = None, [], []
    for it in range(6):
        t = time.perf_counter(); vcf = det.run(batch); wall = time.perf_counter() - t
        s = det.stats()
        cur = dict(prepass=s.prepass_s, pack=s.pack_s, device=s.device_s, total=s.total_s)
        d = {k: cur[k] - (prev[k] if prev else 0) for k in cur}
        prev = cur
        if it:
            walls.append(wall); stages.append(d)
    best = min(range(len(walls)), key=lambda i: walls[i])
    print(name, mode, "walls", " ".join("%.3f" % w for w in walls), "| best:", {k: round(v, 3) for k, v in stages[best].items()}, flush=True)
    det.close()
