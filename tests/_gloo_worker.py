"""Worker of tests/test_region_shards_gloo.py: one rank of a world_size-2 gloo job on CPU.

Each rank runs the host pipeline (admission + event pre-pass + tiling; host_only: no device work) for ITS region shards,
then the ranks exchange their records and counters exactly as bench.py / a multi-GPU detector run does: no data-path
collective, only the final gather of per-shard records and the SUM / MAX reduction of the counters."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ngsepcore_b200 import hostapi, simulate  # noqa: E402
import bench  # noqa: E402


def main():
    out_path = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    if len(sys.argv) > 2 and sys.argv[2] == "multisample":   # MultisampleVariantsDetector: the same region sharding, all samples of a tile on one rank
        genome, batch, _ = simulate.simulate_population([("chrA", 90000), ("chrB", 40000)], n_samples=6, depth=9, mode="wgs", indel_rate=2e-3, seeds=(17, 18, 19))
        det = hostapi.MultisampleDetector(list(zip(batch.rg_names, batch.rg_samples)), host_only=True, tile_len=15000, shard=rank, n_shards=world)
    else:
        genome, batch, _ = simulate.simulate(contigs=[("chrA", 150000), ("chrB", 60000)], depth=20, seeds=(7, 8, 9), indel_rate=1e-3)
        det = hostapi.Detector(host_only=True, tile_len=25000, shard=rank, n_shards=world)
    det.set_genome(genome)
    lines = det.run(batch, with_header=False).splitlines()
    st = det.stats()
    gathered = [None] * world
    dist.all_gather_object(gathered, lines)
    totals, maxima = bench.reduce_over_ranks(dist, torch.device("cpu"), sums=[float(len(lines)), float(st.admitted)], maxima=[float(rank + 1)])
    if rank == 0:
        json.dump({"lines": gathered, "totals": totals, "maxima": maxima, "world": world}, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
