"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/ngsep_pileup.h declares, and fails
loudly (no fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ngsepcore_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "ngsep_pileup.h")).read()
    declared = set(re.findall(r"\b(ngsep_[a-z_]+)\s*\(", hdr))
    declared -= {"ngsep_tile_submit_"}  # nothing to exclude; kept for clarity
    assert set(capi.EXPORTS) == declared
    lib = capi.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ngsep_abi_version() == 2004 and lib.ngsep_built_sm() == 100


def test_struct_layouts_match_header():
    assert capi.HDR_DTYPE.itemsize == 16
    assert capi.CALL_DTYPE.itemsize == 224
    assert C.sizeof(capi.Params) == 32
    assert C.sizeof(capi.Tile) == 104 and C.sizeof(capi.TileStats) == 64
    assert capi.SITE_DTYPE.itemsize == 44 and capi.SAMPLE_CALL_DTYPE.itemsize == 104   # ngsep_pop_site, ngsep_sample_call


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="GPU present")
def test_no_gpu_fails_loudly():
    with pytest.raises(capi.NgsepError):
        capi.Context()


def test_bad_arguments_rejected():
    lib = capi.load_library()
    h = C.c_void_p()
    p = capi.Params(30, 40, 0.001, 2, 0, 1, 0)
    t = capi.host_tables()
    assert lib.ngsep_ctx_create(0, C.byref(p), t.ctypes.data_as(C.POINTER(C.c_double)), 3, C.byref(h)) == -1
    p3 = capi.Params(30, 40, 0.001, 3, 0, 2, 0)      # the pool algorithm (ploidy >= 3) exists for single-sample contexts only
    assert lib.ngsep_ctx_create(0, C.byref(p3), t.ctypes.data_as(C.POINTER(C.c_double)), 95, C.byref(h)) == -1
    assert b"ploidy" in lib.ngsep_last_error(None)


def test_pypack_layout():
    from ngsepcore_b200 import pypack, simulate
    g, b, _ = simulate.simulate(contigs=[("c", 5000)], depth=5, indel_rate=0.0, n_frac=0.0)
    idx = pypack.admit(b)
    hdr, pay, qual, span = pypack.pack_tile(b, idx)
    assert span == 100 and (np.diff(hdr["first"]) >= 0).all()
    i = 7
    L = int(hdr["meta"][i] & 0xFFFF)
    o = int(hdr["off"][i])
    assert pay[o] == 100 * 8 + 3
    codes = (pay[o + 1 + np.arange(L) // 16] >> ((np.arange(L) % 16) * 2).astype(np.uint32)) & 3
    src = b.bases[b.base_off[idx[i]]:b.base_off[idx[i] + 1]]
    assert bytes(np.frombuffer(b"ACGT", np.uint8)[codes]) == bytes(src)
    src_q = b.quals[b.base_off[idx[i]]:b.base_off[idx[i] + 1]]
    lowq = (pay[o + 1 + 7 + np.arange(L) // 32] >> (np.arange(L) % 32).astype(np.uint32)) & 1     # low-quality mask after the bases
    assert np.array_equal(lowq.astype(bool), src_q.astype(np.int64) - 33 <= 3)
    qoff = int(pay[o + 1 + 7 + 4])                                                                  # then the quality offset
    assert bytes(qual[4 * qoff:4 * qoff + L]) == bytes(src_q) and qual.size % 4 == 0
    assert hdr["meta"][i] & capi.M_SIMPLE
