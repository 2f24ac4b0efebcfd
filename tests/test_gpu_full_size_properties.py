"""Full-size checks (BASELINE configs[1], the bench workload: yeast-like 12.16 Mb x 30x, 3.65 M reads) through
size-independent properties — the oracle would need minutes here, so nothing is compared with it:
  * genotyped positions = the positions covered by an admitted read, recomputed with numpy from the read intervals;
  * region tiling is invisible: 300 kb tiles give the same VCF text as one tile per sequence;
  * a second run of the same detector gives the same bytes (no state leaks between runs);
  * every record is well formed (ascending positions per sequence, DP >= sum(BSDP), GQ >= minQuality) and the planted
    homozygous SNVs that are well covered are found with their planted allele."""
import numpy as np
import pytest

from ngsepcore_b200 import hostapi, pypack, simulate

pytestmark = pytest.mark.gpu


def test_yeast_30x_properties():
    genome, batch, truth = simulate.simulate(depth=30.0, indel_rate=0.0)      # the bench workload without indels: reads never change
    det = hostapi.Detector()
    det.set_genome(genome)
    text = det.run(batch)
    st = det.stats()
    # (1) covered positions from the admitted intervals
    idx = pypack.admit(batch)
    covered = 0
    cov_by_contig = {}
    for ci, (name, length) in enumerate(zip(batch.contig_names, batch.contig_lengths)):
        sel = idx[batch.seq_idx[idx] == ci]
        diff = np.zeros(length + 2, np.int32)
        first = batch.first[sel].astype(np.int64)
        last = np.minimum(first + 99, length)
        np.add.at(diff, first, 1)
        np.add.at(diff, last + 1, -1)
        depth = np.cumsum(diff)[1:length + 1]
        cov_by_contig[name] = depth
        covered += int((depth > 0).sum())
    assert st.covered_positions == covered
    assert st.admitted == len(idx)
    # (2) tiling invariance, (3) idempotence
    det_t = hostapi.Detector(tile_len=300000)
    det_t.set_genome(genome)
    assert det_t.run(batch) == text
    assert det.run(batch) == text
    # (4) record sanity + recall of the planted homozygous SNVs
    called = {}
    last = {}
    for line in text.splitlines():
        if line.startswith("#"):
            continue
        f = line.split("\t")
        pos = int(f[1])
        assert pos > last.get(f[0], 0)
        last[f[0]] = pos
        s = dict(zip(f[8].split(":"), f[9].split(":")))
        assert int(s["DP"]) >= sum(int(x) for x in s["BSDP"].split(",")) and int(s["GQ"]) >= 40
        called[(f[0], pos)] = (f[4], s["GT"])
    assert len(called) > 10000
    hit = tot = 0
    for name, variants in truth.items():
        depth = cov_by_contig[name]
        for v in variants:
            if v.kind == "snv" and v.hom and depth[v.pos - 1] >= 15:
                tot += 1
                c = called.get((name, v.pos))
                hit += c is not None and c[0] == v.alt.decode() and c[1] == "1/1"
    assert tot > 4000 and hit / tot > 0.995
