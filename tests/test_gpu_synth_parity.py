"""Device-generated inputs (BASELINE configs[2] and [4]: the 64 Mb / 3.1 Gb genomes and the 1000x panel are generated in HBM,
ngsep_synth_reference / ngsep_synth_tile) against the oracle: the generated reference and tiles are downloaded, written as
FASTA + SAM text, and run through the CPU restatement of the reference's detector; calls and genotyped positions must be the
oracle's.  Also: a region cut into several tiles gives the same calls as one tile (every tile of a genome sees the same reads)."""
import numpy as np
import pytest

from ngsepcore_b200 import calls as callfmt
from ngsepcore_b200 import capi, hostapi, simulate
from tests import oracle_util as ou

pytestmark = pytest.mark.gpu


def _tile_to_sam(path, name, length, tile, keep_from=None):
    """SAM text of a downloaded synthetic tile (simple reads: start, strand, 2-bit bases, 2-bit quality codes)."""
    hdr, pay, qual = tile["hdr"], tile["payload"], tile["qual"]
    bits, qdict = tile["quality"]
    assert bits == 2
    lut_q = np.frombuffer(qdict, np.uint8)[:4]
    lut_b = np.frombuffer(b"ACGT", np.uint8)
    n = hdr.shape[0]
    L = int(hdr["meta"][0] & 0xFFFF)
    nb, nw = (L + 15) // 16, (L + 31) // 32
    W, Q = 1 + nb + nw + 1, ((L * 2 + 7) // 8 + 3) // 4
    assert pay.size == n * W and qual.size == n * Q * 4
    words = pay.reshape(n, W)[:, 1:1 + nb]
    bases = lut_b[((words[:, :, None] >> (2 * np.arange(16, dtype=np.uint32))[None, None, :]) & 3).reshape(n, -1)[:, :L]]
    qwords = qual.view(np.uint32).reshape(n, Q)[:, :nb]
    quals = lut_q[((qwords[:, :, None] >> (2 * np.arange(16, dtype=np.uint32))[None, None, :]) & 3).reshape(n, -1)[:, :L]]
    neg = (hdr["meta"] & capi.M_NEG) != 0
    with open(path, "w") as f:
        f.write(f"@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:{name}\tLN:{length}\n")
        for i in range(n):
            if keep_from is not None and i < keep_from:
                continue
            f.write(f"r{i}\t{16 if neg[i] else 0}\t{name}\t{int(hdr['first'][i])}\t60\t{L}M\t*\t0\t0\t{bases[i].tobytes().decode()}\t{quals[i].tobytes().decode()}\n")
    return n


@pytest.mark.parametrize("depth,length", [(30.0, 250000), (60.0, 120000), (1000.0, 4000)])
def test_generated_tiles_match_the_oracle(oracle, tmp_path, depth, length):
    with capi.Context() as ctx:
        ctx.synth_reference(0, length, seed=1031, gc=0.41)
        slot = ctx.synth_tile(0, 1, length, depth=depth, seed=3033)
        ctx.resident_run(-1)
        calls, _, st = ctx.fetch(slot=slot)
        tile = ctx.resident_download(slot)
        ref = ctx.ref_download(0, length)
        assert set(np.unique(ref).tolist()) <= set(b"ACGT") and 0.38 < np.isin(ref, list(b"GC")).mean() < 0.44
        assert abs(tile["n_reads"] * 100 / (length - 100) - depth) < 0.02 * depth + 1      # reads end inside the sequence
        fa, sam = str(tmp_path / "r.fa"), str(tmp_path / "r.sam")
        simulate.write_fasta(fa, {"chrS": ref})
        _tile_to_sam(sam, "chrS", length, tile)
        want, _, gp, _ = ou.run_ssvd(oracle, fa, sam, args=("-maxAlnsPerStartPos", "0"))
        want_lines = [r["line"] for r in ou.parse_vcf(want)]
        got_lines = [callfmt.snv_record_line(c, "chrS") for c in calls]
        assert st.covered_positions == gp
        assert len(want_lines) > (100 if depth < 100 else 2)
        assert got_lines == want_lines
        # the same region as three tiles: every tile sees the same reads
        ctx.resident_clear()
        cuts = [1, length // 3, 2 * length // 3, length + 1]
        slots = [ctx.synth_tile(0, cuts[k], cuts[k + 1] - 1, depth=depth, seed=3033) for k in range(3)]
        ctx.resident_run(-1)
        parts = [ctx.fetch(slot=s)[0] for s in slots]
        again = [callfmt.snv_record_line(c, "chrS") for p in parts for c in p]
        assert again == want_lines
        assert ctx.resident_stats().covered_positions == gp
