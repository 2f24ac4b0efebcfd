"""Multi-rank path on CPU (gloo, world_size 2): region shards of the host pipeline partition the records (no data-path
collective), and bench.py's cross-rank reduction (SUM of units, MAX of time) works as the contract asks."""
import json
import os
import socket
import subprocess
import sys

from ngsepcore_b200 import hostapi, simulate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_ranks_partition_the_records(tmp_path):
    out = str(tmp_path / "gathered.json")
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gloo_worker.py"), out], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    got = json.load(open(out))
    assert got["world"] == 2
    # the single-process, single-shard run is the truth
    genome, batch, _ = simulate.simulate(contigs=[("chrA", 150000), ("chrB", 60000)], depth=20, seeds=(7, 8, 9), indel_rate=1e-3)
    det = hostapi.Detector(host_only=True)
    det.set_genome(genome)
    want = det.run(batch, with_header=False).splitlines()
    assert len(want) > 20
    per_rank = got["lines"]
    assert all(len(l) > 0 for l in per_rank)                      # both ranks own some tiles
    merged = [l for lines in per_rank for l in lines]
    assert len(merged) == len(want) and set(merged) == set(want)  # disjoint shards, nothing lost, nothing duplicated
    order = {n: i for i, n in enumerate(genome)}
    merged.sort(key=lambda l: (order[l.split("\t")[0]], int(l.split("\t")[1])))
    assert merged == want                                          # host merge in (sequence, position) order
    # bench.py reduction: SUM over ranks for units, MAX over ranks for time-like values
    assert got["totals"][0] == len(want)
    assert got["totals"][1] == 2 * det.stats().admitted            # admission is global: every rank admits every record
    assert got["maxima"] == [2.0]


def test_two_ranks_partition_the_population_records(oracle, tmp_path):
    """The sample axis shards by region as well (all samples of a tile on one rank): two gloo ranks own disjoint
    region tiles of the multi-sample host pipeline and together reproduce the ORACLE's population indel / STR records."""
    from tests import oracle_util as ou
    out = str(tmp_path / "gathered_ms.json")
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gloo_worker.py"), out, "multisample"], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    got = json.load(open(out))
    genome, batch, _ = simulate.simulate_population([("chrA", 90000), ("chrB", 40000)], n_samples=6, depth=9, mode="wgs", indel_rate=2e-3, seeds=(17, 18, 19))
    fa, sam = str(tmp_path / "r.fa"), str(tmp_path / "r.sam")
    simulate.write_fasta(fa, genome)
    simulate.write_sam(sam, batch)
    text, _, _, _ = ou.run_msvd(oracle, fa, [sam])
    want = [l for l in text.splitlines() if "TYPE=INDEL" in l or "TYPE=STR" in l]
    assert len(want) > 15
    per_rank = got["lines"]
    assert all(len(l) > 0 for l in per_rank)
    merged = [l for lines in per_rank for l in lines]
    order = {n: i for i, n in enumerate(genome)}
    merged.sort(key=lambda l: (order[l.split("\t")[0]], int(l.split("\t")[1])))
    assert merged == want
