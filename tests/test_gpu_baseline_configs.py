"""GPU parity on BASELINE.json's own configurations, byte for byte against the oracle:

  * configs[1] — the exact bench workload: yeast-like 12.16 Mb x 30x, 2x100 bp, planted SNVs AND indels (the realigner
    fires, reads are versioned, indel records come from the host caller) through hostapi.Detector; the oracle runs as one
    forked process per sequence (-querySeq, the reference's only parallel mode, APG:310-322 — whole sequences, so that the
    concatenation of the per-process records IS the whole-genome run: no realigner state is cut);
  * the same workload at the depth of configs[2] (60x);
  * configs[3] — MultisampleVariantsDetector, 96 GBS-style samples at 10x on the 17-sequence yeast-like genome with
    -maxAlnsPerStartPos 100: the detector runs the whole population (about 48 M reads), the oracle > 1 Mb of it — four
    whole sequences cut only inside gaps no read comes near, so every window is exact —, records of those windows identical.
    Sized for the GPU box (>= 16 cores, >= 100 GB of free RAM); smaller hosts run the same test on fewer tags."""
import os
import numpy as np
import pytest

from ngsepcore_b200 import hostapi, simulate
from tests import oracle_util as ou

pytestmark = pytest.mark.gpu


def _body(text):
    return [l for l in text.splitlines() if not l.startswith("#")]


def _check_yeast(tmp_path, depth, tile_len):
    genome, batch, _ = simulate.simulate(depth=depth)          # bench.py's _workload(rank 0): same seeds, indels included
    fa = str(tmp_path / "ref.fa")
    simulate.write_fasta(fa, genome)
    want, gp, _ = ou.run_regions(genome, batch, fa, piece=1 << 30)      # whole sequences
    det = hostapi.Detector(tile_len=tile_len)
    det.set_genome(genome)
    got = _body(det.run(batch))
    st = det.stats()
    assert st.covered_positions == gp
    assert len(want) > 12000 and sum(1 for l in want if "TYPE=INDEL" in l or "TYPE=STR" in l) > 500
    if got != want:
        bad = next(i for i, (a, b) in enumerate(zip(got + [""], want + [""])) if a != b)
        raise AssertionError(f"{len(got)} vs {len(want)} records; first difference at record {bad}:\n got  {got[bad] if bad < len(got) else None}\n want {want[bad] if bad < len(want) else None}")
    assert st.indel_calls > 500 and st.versions > 1000


def test_configs1_yeast_30x_with_indels_vs_oracle(tmp_path):
    _check_yeast(tmp_path, 30.0, 0)


def test_configs2_depth_60x_with_indels_vs_oracle(tmp_path):
    _check_yeast(tmp_path, 60.0, 400000)     # region tiles, as the 64 Mb configuration is run


def _mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def _gap_cuts(batch, ci, length, piece, margin=400):
    """Windows of about `piece` positions of sequence ci whose borders sit inside gaps: no record starts within `margin`
    positions of a border, so no read, realigner event or indel record can see across it (reads are 100 bp)."""
    lo, hi = np.searchsorted(batch.seq_idx, ci, "left"), np.searchsorted(batch.seq_idx, ci, "right")
    starts = np.unique(batch.first[lo:hi].astype(np.int64))
    cuts, want = [1], piece
    while want < length - piece // 2:
        k = np.searchsorted(starts, want)
        while k + 1 < len(starts) and starts[k + 1] - starts[k] < 2 * margin + 1:   # next gap wide enough
            k += 1
        if k + 1 >= len(starts):
            break
        c = int(starts[k] + starts[k + 1]) // 2
        cuts.append(c)
        want = c + piece
    cuts.append(length + 1)
    return [(ci, cuts[i], cuts[i + 1] - 1) for i in range(len(cuts) - 1)]


def test_configs3_gbs_96_samples_full_genome(tmp_path):
    contigs = simulate.YEAST_CONTIGS
    cores = os.cpu_count() or 1
    full = cores >= 16 and _mem_available_gb() >= 100.0
    n_tags = 50000 if full else 4000
    genome, batch, _ = simulate.simulate_population(contigs, n_samples=96, depth=10.0, mode="gbs", n_tags_total=n_tags, seeds=(1021, 2022, 3023),
                                                    workers=min(cores, 32))
    assert batch.n > (40_000_000 if full else 3_000_000)
    fa = str(tmp_path / "ref.fa")
    simulate.write_fasta(fa, genome)
    det = hostapi.MultisampleDetector(list(zip(batch.rg_names, batch.rg_samples)), maxAlnsPerStartPos=100)
    det.set_genome(genome)
    got = _body(det.run(batch))
    # the oracle on > 1 Mb of it: four whole sequences (chrI, chrIII, chrVI, chrIX), each cut only inside read-free gaps
    regions = []
    for ci in (0, 2, 5, 8):
        regions += _gap_cuts(batch, ci, contigs[ci][1], 60000)
    assert sum(l - f + 1 for _, f, l in regions) >= 1_000_000
    want, _, _ = ou.run_regions(genome, batch, fa, regions=regions, args=("-maxAlnsPerStartPos", "100"), multi=True)
    names = [n for n, _ in contigs]
    pick = {names[ci] for ci, _, _ in regions}
    inside = [l for l in got if l.split("\t", 1)[0] in pick]
    assert len(want) > (300 if full else 30)
    if inside != want:
        bad = next(i for i, (a, b) in enumerate(zip(inside + [""], want + [""])) if a != b)
        raise AssertionError(f"{len(inside)} vs {len(want)} records; first difference at record {bad}:\n got  {inside[bad][:300] if bad < len(inside) else None}\n want {want[bad][:300] if bad < len(want) else None}")
