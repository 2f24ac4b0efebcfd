"""ctypes binding of libngsep_pileup.so (include/ngsep_pileup.h) — the same stub INTEGRATION.md shows.

No fallback: importing works without a GPU (so CPU-only CI can check the exported symbols), but every compute
entry point raises ``NgsepError`` when the library or a B200 is missing.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NGSEP_PILEUP_LIB") or os.path.join(_HERE, "libngsep_pileup.so")   # (the override is for kernel experiments)

N_TABLES = 95
F_STRAND_BIAS, F_IGNORE_LOWER_REF, F_EMIT_ALL, F_QUALITY_IN_PLACE = 0x1, 0x2, 0x8, 0x10
M_NEG, M_NMASK, M_WINDOW, M_SIMPLE = 0x01000000, 0x02000000, 0x04000000, 0x10000000
CALL_HET, CALL_HOMALT, CALL_TRIALLELIC = 1, 2, 3


class NgsepError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("max_base_qs", C.c_int32), ("min_quality", C.c_int32), ("het_rate", C.c_double), ("ploidy", C.c_int32),
                ("flags", C.c_uint32), ("n_samples", C.c_int32), ("reserved", C.c_int32)]


class Tile(C.Structure):
    _fields_ = [("seq_idx", C.c_int32), ("tile_first", C.c_int32), ("tile_last", C.c_int32), ("n_reads", C.c_int32),
                ("max_span", C.c_int32), ("max_back", C.c_int32), ("n_words", C.c_uint64), ("hdr", C.c_void_p),
                ("payload", C.c_void_p), ("group", C.c_void_p), ("qual", C.c_void_p), ("n_qual", C.c_uint64),
                ("qual_bits", C.c_uint32), ("qual_dict", C.c_uint8 * 16), ("layout", C.c_uint32), ("n_wire_words", C.c_uint64)]


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("depth", C.c_double), ("read_len", C.c_int32), ("reserved", C.c_int32), ("snv_rate", C.c_double),
                ("error_rate", C.c_double)]


class TileStats(C.Structure):
    _fields_ = [("covered_positions", C.c_uint64), ("candidate_positions", C.c_uint64), ("calls", C.c_uint64),
                ("aligned_bases", C.c_uint64), ("kernel_ms", C.c_float), ("pileup_ms", C.c_float), ("kernel_launches", C.c_uint32), ("open_columns", C.c_uint32),
                ("h2d_bytes", C.c_uint64), ("quality_in_place", C.c_uint32), ("d2h_bytes", C.c_uint32)]


HDR_DTYPE = np.dtype([("first", "<i4"), ("meta", "<u4"), ("ign_start", "<i2"), ("ign_end", "<i2"), ("off", "<u4")])
CALL_DTYPE = np.dtype([("seq_idx", "<i4"), ("pos", "<i4"), ("count", "<i4", (4,)), ("depth", "<i4"), ("strand", "<i4", (8,)),
                       ("logcond", "<f8", (16,)), ("post_ref", "<f8"), ("post_max", "<f8"), ("gq", "<i2"), ("qual", "<i2"),
                       ("a0", "i1"), ("a1", "i1"), ("ref", "i1"), ("kind", "i1"), ("sample", "<i4")], align=True)

class IndelEvents(C.Structure):      # ngsep_indel_events
    _fields_ = [("n_events", C.c_int32), ("n_alleles", C.c_int32), ("n_calls", C.c_int32), ("reserved", C.c_int32), ("allele_first", C.c_void_p),
                ("call_first", C.c_void_p), ("allele_off", C.c_void_p), ("call_off", C.c_void_p), ("allele_chars", C.c_void_p), ("call_chars", C.c_void_p),
                ("call_quals", C.c_void_p), ("call_neg", C.c_void_p), ("log_err", C.c_double)]


EXPORTS = ["ngsep_ctx_create", "ngsep_ctx_destroy", "ngsep_last_error", "ngsep_ref_upload", "ngsep_tile_submit",
           "ngsep_tile_fetch", "ngsep_resident_add", "ngsep_resident_clear", "ngsep_resident_run", "ngsep_resident_fetch", "ngsep_resident_stats",
           "ngsep_timer_start", "ngsep_timer_stop", "ngsep_host_register", "ngsep_host_unregister", "ngsep_abi_version", "ngsep_built_sm",
           "ngsep_samples_set", "ngsep_tile_fetch_pop", "ngsep_resident_fetch_pop",
           "ngsep_synth_reference", "ngsep_synth_tile", "ngsep_resident_download", "ngsep_ref_download", "ngsep_coverage_histogram", "ngsep_tile_sites", "ngsep_tile_allele_counts", "ngsep_indel_tallies"]

# multi-sample results (include/ngsep_pileup.h: ngsep_pop_site, ngsep_sample_call)
SITE_DTYPE = np.dtype([("seq_idx", "<i4"), ("pos", "<i4"), ("count", "<i4", (4,)), ("depth", "<i4"), ("qs", "<i2"), ("ref", "i1"), ("n_alleles", "i1"),
                       ("allele", "i1", (4,)), ("n_alleles0", "i1"), ("pad", "i1", (7,))])
SAMPLE_CALL_DTYPE = np.dtype([("count", "<i4", (4,)), ("depth", "<i4"), ("gq", "<i2"), ("g0", "i1"), ("g1", "i1"), ("logcond", "<f8", (10,))])

_lib = None


def load_library() -> C.CDLL:
    """Loads the CUDA library; raises NgsepError (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NgsepError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(LIB_PATH)
    lib.ngsep_ctx_create.argtypes = [C.c_int, C.POINTER(Params), C.POINTER(C.c_double), C.c_size_t, C.POINTER(C.c_void_p)]
    lib.ngsep_ctx_destroy.argtypes = [C.c_void_p]
    lib.ngsep_last_error.argtypes = [C.c_void_p]
    lib.ngsep_last_error.restype = C.c_char_p
    lib.ngsep_ref_upload.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]
    lib.ngsep_tile_submit.argtypes = [C.c_void_p, C.POINTER(Tile)]
    lib.ngsep_tile_sites.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.ngsep_indel_tallies.argtypes = [C.c_void_p, C.POINTER(IndelEvents), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ngsep_tile_allele_counts.argtypes = [C.c_void_p, C.POINTER(Tile), C.c_int32, C.c_int32, C.c_void_p]
    lib.ngsep_resident_add.argtypes = [C.c_void_p, C.POINTER(Tile), C.POINTER(C.c_int)]
    lib.ngsep_resident_clear.argtypes = [C.c_void_p]
    lib.ngsep_resident_run.argtypes = [C.c_void_p, C.c_int]
    lib.ngsep_resident_fetch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p, C.POINTER(TileStats)]
    lib.ngsep_resident_stats.argtypes = [C.c_void_p, C.POINTER(TileStats)]
    lib.ngsep_timer_start.argtypes = [C.c_void_p]
    lib.ngsep_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    lib.ngsep_host_register.argtypes = [C.c_void_p, C.c_size_t]
    lib.ngsep_host_unregister.argtypes = [C.c_void_p]
    lib.ngsep_tile_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p, C.POINTER(TileStats)]
    lib.ngsep_samples_set.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_double]
    lib.ngsep_tile_fetch_pop.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p, C.POINTER(TileStats)]
    lib.ngsep_resident_fetch_pop.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p,
                                             C.POINTER(TileStats)]
    lib.ngsep_synth_reference.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double]
    lib.ngsep_synth_tile.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(SynthParams), C.POINTER(C.c_int)]
    lib.ngsep_resident_download.argtypes = [C.c_void_p, C.c_int, C.POINTER(Tile), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ngsep_ref_download.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]
    lib.ngsep_coverage_histogram.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


def host_register(arr: np.ndarray) -> None:
    """Page-locks a numpy buffer (ngsep_host_register)."""
    rc = load_library().ngsep_host_register(arr.ctypes.data, arr.nbytes)
    if rc != 0:
        raise NgsepError(f"ngsep_host_register failed ({rc}): {load_library().ngsep_last_error(None).decode()}")


def host_unregister(arr: np.ndarray) -> None:
    load_library().ngsep_host_unregister(arr.ctypes.data)


def host_tables(het_rate: float = 0.001) -> np.ndarray:
    """The 95 doubles of include/ngsep_pileup.h (NGSEP_N_TABLES), restating CountsHelper.updateProbabilitiesCache
    (reference: src/ngsep/discovery/CountsHelper.java:135-187, 414-415) with the host libm."""
    import math
    t = np.zeros(N_TABLES, np.float64)
    for q in range(3, 31):
        e = math.pow(10.0, -0.1 * q)          # PhredScoreHelper.calculateProbability
        s = 1 - e
        af = 250.0 / 500.0
        t[3 * q] = math.log10(s)
        t[3 * q + 1] = -0.1 * q - math.log10(3.0)
        t[3 * q + 2] = math.log10(af * s + (1 - af) * e / 3)
    t[93] = math.log10((1 - het_rate) / 4)
    t[94] = math.log10(het_rate / 12)
    return t


class Context:
    """One GPU context (ngsep_ctx_create .. ngsep_ctx_destroy)."""

    def __init__(self, device: int = 0, max_base_qs: int = 30, min_quality: int = 40, het_rate: float = 0.001, ploidy: int = 2,
                 flags: int = 0, tables: np.ndarray | None = None, n_samples: int = 1):
        self.lib = load_library()
        self.n_samples = n_samples
        self.params = Params(max_base_qs, min_quality, het_rate, ploidy, flags, n_samples, 0)
        tab = np.ascontiguousarray(host_tables(het_rate) if tables is None else tables, np.float64)
        self.h = C.c_void_p()
        rc = self.lib.ngsep_ctx_create(device, C.byref(self.params), tab.ctypes.data_as(C.POINTER(C.c_double)), tab.size, C.byref(self.h))
        if rc != 0:
            raise NgsepError(f"ngsep_ctx_create failed ({rc}): {self.lib.ngsep_last_error(None).decode()}")
        self._keep = []
        self._slot_npos = {}
        self._inflight_npos = []      # positions of the pipelined tiles, oldest first (tiles are fetched in submission order)

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise NgsepError(f"{what} failed ({rc}): {self.lib.ngsep_last_error(self.h).decode()}")

    def ref_upload(self, seq_idx: int, bases: np.ndarray) -> None:
        b = np.ascontiguousarray(bases, np.uint8)
        self._check(self.lib.ngsep_ref_upload(self.h, seq_idx, b.ctypes.data, b.size), "ngsep_ref_upload")

    def _tile(self, seq_idx, tile_first, tile_last, hdr, payload, qual, max_span, max_back, group=None, quality=None, layout=None) -> Tile:
        """quality: None = `qual` holds phred+33 bytes; (bits, dict16) = 2- / 4-bit codes (ngsep_tile.qual_bits / qual_dict).
        layout: None = `payload` is the full layout; (1, n_words_full) = the compact wire stream (ngsep_tile.layout / n_words)."""
        hdr = np.ascontiguousarray(hdr, HDR_DTYPE)
        payload = np.ascontiguousarray(payload, np.uint32)
        qual = np.ascontiguousarray(qual, np.uint8)
        group = None if group is None else np.ascontiguousarray(group, np.uint16)
        self._keep = (getattr(self, "_keep", []) + [hdr, payload, qual, group])[-16:]   # buffers of the tiles still in flight
        bits, qdict = (0, bytes(16)) if quality is None else (int(quality[0]), bytes(quality[1]))
        lay, n_full = (0, payload.size) if layout is None or not layout[0] else (int(layout[0]), int(layout[1]))
        return Tile(seq_idx, tile_first, tile_last, hdr.shape[0], max_span, max_back, n_full,
                    hdr.ctypes.data if hdr.size else None, payload.ctypes.data if payload.size else None,
                    group.ctypes.data if group is not None and group.size else None,
                    qual.ctypes.data if qual.size else None, qual.size, bits, (C.c_uint8 * 16)(*qdict), lay, payload.size)

    def samples_set(self, sample_groups, n_groups: int, min_allele_depth_freq: float = 0.0) -> None:
        """sample_groups: per sample, the read-group ids in the order PR.getAlleleCalls(span, Set) visits them."""
        off = np.zeros(len(sample_groups) + 1, np.int32)
        off[1:] = np.cumsum([len(g) for g in sample_groups])
        flat = np.array([g for gs in sample_groups for g in gs] or [0], np.uint16)
        self._check(self.lib.ngsep_samples_set(self.h, n_groups, off.ctypes.data, flat.ctypes.data, min_allele_depth_freq), "ngsep_samples_set")

    def fetch_pop(self, cap: int = 4096, want_counts: bool = False, slot: int = 0):
        """Multi-sample contexts: (sites sorted by pos, sample calls [n_sites, n_samples], pooled counts or None, TileStats)."""
        stats = TileStats()
        n_pos = (self._inflight_npos[0] if self._inflight_npos else 0) if slot == 0 else self._slot_npos[slot]
        counts = np.zeros((n_pos, 5), np.int32) if want_counts else None
        while True:
            sites = np.zeros(cap, SITE_DTYPE)
            calls = np.zeros((cap, self.n_samples), SAMPLE_CALL_DTYPE)
            n = C.c_size_t(0)
            cp = counts.ctypes.data if want_counts else None
            if slot == 0:
                rc = self.lib.ngsep_tile_fetch_pop(self.h, sites.ctypes.data, cap, C.byref(n), calls.ctypes.data, cp, C.byref(stats))
            else:
                rc = self.lib.ngsep_resident_fetch_pop(self.h, slot, sites.ctypes.data, cap, C.byref(n), calls.ctypes.data, cp, C.byref(stats))
            if rc == -6 and n.value > cap:
                cap = int(n.value) + 16
                continue
            if slot == 0:
                if self._inflight_npos:
                    self._inflight_npos.pop(0)      # the tile left the queue, whether the fetch succeeded or not
            self._check(rc, "ngsep_tile_fetch_pop")
            break
        o = np.argsort(sites["pos"][: n.value], kind="stable")
        return sites[: n.value][o], calls[: n.value][o], counts, stats

    def tile_sites(self, pos, alleles) -> None:
        """-knownVariants: the SNV sites of the NEXT submitted tile.  alleles: per site the base indexes (A0 C1 G2 T3), reference first."""
        pos = np.ascontiguousarray(pos, np.int32)
        codes = np.ascontiguousarray([sum(b << (2 * k) for k, b in enumerate(a)) | (len(a) << 8) for a in alleles], np.uint16)
        self._check(self.lib.ngsep_tile_sites(self.h, len(pos), pos.ctypes.data, codes.ctypes.data), "ngsep_tile_sites")

    def submit(self, seq_idx, tile_first, tile_last, hdr, payload, qual, max_span, max_back=0, group=None, quality=None, layout=None) -> None:
        t = self._tile(seq_idx, tile_first, tile_last, hdr, payload, qual, max_span, max_back, group, quality, layout)
        self._check(self.lib.ngsep_tile_submit(self.h, C.byref(t)), "ngsep_tile_submit")
        self._inflight_npos.append(tile_last - tile_first + 1)

    def resident_add(self, seq_idx, tile_first, tile_last, hdr, payload, qual, max_span, max_back=0, group=None, quality=None, layout=None) -> int:
        """Stages a tile in HBM for repeated runs; returns its slot."""
        t = self._tile(seq_idx, tile_first, tile_last, hdr, payload, qual, max_span, max_back, group, quality, layout)
        slot = C.c_int(0)
        self._check(self.lib.ngsep_resident_add(self.h, C.byref(t), C.byref(slot)), "ngsep_resident_add")
        self._slot_npos[slot.value] = tile_last - tile_first + 1
        return slot.value

    def indel_tallies(self, events, log_err: float):
        """events: list of (alleles [str, reference first], calls [(allele str, quality str phred+33, negative strand bool)]) -> per event
        dict(total, counts, strand [n,2], diag, call_index, call_fit): the per-call part of CountsHelper.updateCountsIndel (ngsep_indel_tallies)."""
        af, cf, ao, co = [0], [0], [0], [0]
        ac, cc, cq, cn = bytearray(), bytearray(), bytearray(), bytearray()
        for alleles, calls in events:
            for a in alleles:
                ac += a.encode(); ao.append(len(ac))
            for a, q, neg in calls:
                assert len(a) == len(q)
                cc += a.encode(); cq += q.encode(); co.append(len(cc)); cn.append(1 if neg else 0)
            af.append(len(ao) - 1); cf.append(len(co) - 1)
        arr = lambda v, t: np.ascontiguousarray(v, t)
        af, cf, ao, co = arr(af, np.int32), arr(cf, np.int32), arr(ao, np.int32), arr(co, np.int32)
        ac, cc, cq, cn = (arr(np.frombuffer(bytes(b) + b"\0", np.uint8), np.uint8) for b in (ac, cc, cq, cn))
        n_ev, n_al, n_ca = len(events), len(ao) - 1, len(co) - 1
        e = IndelEvents(n_ev, n_al, n_ca, 0, af.ctypes.data, cf.ctypes.data, ao.ctypes.data, co.ctypes.data, ac.ctypes.data, cc.ctypes.data, cq.ctypes.data,
                        cn.ctypes.data, log_err)
        total, counts, strand = np.zeros(n_ev, np.int32), np.zeros(n_al, np.int32), np.zeros((n_al, 2), np.int32)
        diag, idx, fit = np.zeros(n_al, np.float64), np.zeros(n_ca + 1, np.int32), np.zeros(n_ca + 1, np.float64)
        self._check(self.lib.ngsep_indel_tallies(self.h, C.byref(e), total.ctypes.data, counts.ctypes.data, strand.ctypes.data, diag.ctypes.data, idx.ctypes.data,
                                                 fit.ctypes.data), "ngsep_indel_tallies")
        return [dict(total=int(total[k]), counts=counts[af[k]:af[k + 1]].copy(), strand=strand[af[k]:af[k + 1]].copy(), diag=diag[af[k]:af[k + 1]].copy(),
                     call_index=idx[cf[k]:cf[k + 1]].copy(), call_fit=fit[cf[k]:cf[k + 1]].copy()) for k in range(n_ev)]

    def tile_allele_counts(self, seq_idx, tile_first, tile_last, hdr, payload, qual, max_span, max_back=0, min_rd=10, min_base_qs=20, quality=None):
        """RelativeAlleleCountsCalculator.onPileup over a tile: uint32 [n_pos, 2] = (largest allele count | alleles seen << 28, second largest count)."""
        t = self._tile(seq_idx, tile_first, tile_last, hdr, payload, qual, max_span, max_back, None, quality)
        out = np.zeros((tile_last - tile_first + 1, 2), np.uint32)
        self._check(self.lib.ngsep_tile_allele_counts(self.h, C.byref(t), min_rd, min_base_qs, out.ctypes.data), "ngsep_tile_allele_counts")
        return out

    def coverage_histogram(self, tile_first: int, tile_last: int, first, last, unique, max_coverage: int = 300):
        """Read-depth histograms (all / unique alignments) of a region: arrays of max_coverage + 1 counts, [max_coverage] = more."""
        first = np.ascontiguousarray(first, np.int32); last = np.ascontiguousarray(last, np.int32); unique = np.ascontiguousarray(unique, np.uint8)
        ha, hu = np.zeros(max_coverage + 1, np.uint64), np.zeros(max_coverage + 1, np.uint64)
        self._check(self.lib.ngsep_coverage_histogram(self.h, tile_first, tile_last, first.size, first.ctypes.data, last.ctypes.data, unique.ctypes.data,
                                                      max_coverage, ha.ctypes.data, hu.ctypes.data), "ngsep_coverage_histogram")
        return ha, hu

    # ---- synthetic inputs generated on the device (BASELINE configs[2], [4]) ---------------------------------------------------
    def synth_reference(self, seq_idx: int, length: int, seed: int = 1031, gc: float = 0.41) -> None:
        self._check(self.lib.ngsep_synth_reference(self.h, seq_idx, length, seed, gc), "ngsep_synth_reference")

    def synth_tile(self, seq_idx: int, tile_first: int, tile_last: int, depth: float = 30.0, read_len: int = 100, seed: int = 3033,
                   snv_rate: float = 1e-3, error_rate: float = 2e-3) -> int:
        """Generates the packed tile of a region in HBM and stages it as a resident slot; returns the slot."""
        p = SynthParams(seed, depth, read_len, 0, snv_rate, error_rate)
        slot = C.c_int(0)
        self._check(self.lib.ngsep_synth_tile(self.h, seq_idx, tile_first, tile_last, C.byref(p), C.byref(slot)), "ngsep_synth_tile")
        self._slot_npos[slot.value] = tile_last - tile_first + 1
        return slot.value

    def resident_download(self, slot: int, data: bool = True):
        """dict(first, last, hdr, payload, qual, quality=(bits, dict16), max_span, max_back) of a resident tile (arrays only with data)."""
        t = Tile()
        self._check(self.lib.ngsep_resident_download(self.h, slot, C.byref(t), None, None, None), "ngsep_resident_download")
        out = dict(seq_idx=t.seq_idx, first=t.tile_first, last=t.tile_last, n_reads=t.n_reads, n_words=t.n_words, n_qual=t.n_qual, max_span=t.max_span,
                   max_back=t.max_back, quality=(t.qual_bits, bytes(t.qual_dict)))
        if data:
            hdr, pay, qual = np.zeros(t.n_reads, HDR_DTYPE), np.zeros(t.n_words, np.uint32), np.zeros(t.n_qual, np.uint8)
            self._check(self.lib.ngsep_resident_download(self.h, slot, C.byref(t), hdr.ctypes.data, pay.ctypes.data, qual.ctypes.data), "ngsep_resident_download")
            out.update(hdr=hdr, payload=pay, qual=qual)
        return out

    def ref_download(self, seq_idx: int, length: int) -> np.ndarray:
        out = np.zeros(length, np.uint8)
        self._check(self.lib.ngsep_ref_download(self.h, seq_idx, out.ctypes.data, length), "ngsep_ref_download")
        return out

    def resident_run(self, slot: int = -1) -> None:
        self._check(self.lib.ngsep_resident_run(self.h, slot), "ngsep_resident_run")

    def resident_stats(self) -> TileStats:
        """Counters of the last batched run summed over every resident tile; kernel_ms = device time of its one launch."""
        st = TileStats()
        self._check(self.lib.ngsep_resident_stats(self.h, C.byref(st)), "ngsep_resident_stats")
        return st

    def resident_clear(self) -> None:
        self._check(self.lib.ngsep_resident_clear(self.h), "ngsep_resident_clear")
        self._slot_npos = {}

    def timer_start(self) -> None:
        self._check(self.lib.ngsep_timer_start(self.h), "ngsep_timer_start")

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        self._check(self.lib.ngsep_timer_stop(self.h, C.byref(ms)), "ngsep_timer_stop")
        return ms.value

    def fetch(self, cap: int = 1 << 16, want_counts: bool = False, slot: int = 0, sort: bool = True):
        """Returns (calls structured array sorted by pos, counts [n_pos,5] or None, TileStats)."""
        stats = TileStats()
        n_pos = (self._inflight_npos[0] if self._inflight_npos else 0) if slot == 0 else self._slot_npos[slot]
        counts = np.zeros((n_pos, 5), np.int32) if want_counts else None
        while True:
            if getattr(self, "_call_buf", None) is None or self._call_buf.shape[0] < cap:
                self._call_buf = np.zeros(cap, CALL_DTYPE)      # reused from fetch to fetch: the caller gets a copy of the n records
            out = self._call_buf
            n = C.c_size_t(0)
            rc = self.lib.ngsep_resident_fetch(self.h, slot, out.ctypes.data, cap, C.byref(n), counts.ctypes.data if want_counts else None, C.byref(stats))
            if rc == -6 and n.value > cap:
                cap = int(n.value) + 16
                continue
            if slot == 0:
                if self._inflight_npos:
                    self._inflight_npos.pop(0)      # the tile left the queue, whether the fetch succeeded or not
            self._check(rc, "ngsep_tile_fetch")
            break
        calls = out[: n.value].copy()
        if sort:
            calls = calls[np.argsort(calls["pos"], kind="stable")]
        return calls, counts, stats

    def close(self) -> None:
        if self.h:
            self.lib.ngsep_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


assert CALL_DTYPE.itemsize == 224 and HDR_DTYPE.itemsize == 16, (CALL_DTYPE.itemsize, HDR_DTYPE.itemsize)
