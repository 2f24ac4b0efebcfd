"""A/B of a host-side switch (NAME=VALUE pairs on the command line) on the product detector: wall time and stage times of the third run."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngsepcore_b200 import hostapi, simulate
genome, batch, _ = simulate.simulate(depth=30.0)
name = sys.argv[1] if len(sys.argv) > 1 else "NGSEP_WIRE_COMPACT"
for mode in ("1", "0", "1", "0"):
    os.environ[name] = mode
    det = hostapi.Detector(); det.set_genome(genome)
    prev = None
    for it in range(3):
        t = time.perf_counter(); vcf = det.run(batch); wall = time.perf_counter() - t
        s = det.stats()
        cur = dict(prepass=s.prepass_s, pack=s.pack_s, device=s.device_s, total=s.total_s)
        d = {k: cur[k] - (prev[k] if prev else 0) for k in cur}
        prev = cur
    print(name, mode, "wall %.3f" % wall, {k: round(v, 3) for k, v in d.items()}, flush=True)
    det.close()
