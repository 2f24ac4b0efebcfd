"""numpy packer: AlignmentBatch -> the tile layout of include/ngsep_pileup.h.

This is the Python mirror of the host-side packing the Java/C++ host performs before ``ngsep_tile_submit``:
  * the record filter of ReadAlignmentFileReader (reference: src/ngsep/alignments/io/ReadAlignmentFileReader.java:
    229-236, 284-291, 350) with the flag set of AlignmentsPileupGenerator.createReader (APG:363-375);
  * the per-read-group same-start cap and the 5'/3' ignore mapping of APG.processSameStartAlns (APG:407-433,
    RA:613-644);
  * CIGAR words, 2-bit bases, N mask and phred+33 qualities (RA:60-69, 581-595, 1180-1198).
It does NOT run the indel realigner (records carry no valid windows); the C++ host pipeline does.
"""
from __future__ import annotations

import numpy as np

from .capi import HDR_DTYPE, M_NEG, M_NMASK, M_SIMPLE
from .simulate import AlignmentBatch

FLAG_UNMAPPED, FLAG_REVERSE, FLAG_SECONDARY, FLAG_MULTIPLE = 0x4, 0x10, 0x100, 0x1000

_CODE = np.full(256, 4, np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i
    _CODE[_c | 0x20] = _i


def admit(batch: AlignmentBatch, min_mq: int = 20, max_alns_per_start: int = 5, secondary: bool = False,
          non_unique_primary: bool = False) -> np.ndarray:
    """Indices of the records AlignmentsPileupGenerator would admit, in admission order."""
    flags = batch.flags.astype(np.int64)
    multiple = (flags & FLAG_SECONDARY) != 0
    nh = batch.nh.astype(np.int64)
    multiple |= nh > 1
    multiple |= (nh < 0) & (batch.mapq.astype(np.int64) < min_mq) & ~multiple
    multiple &= ~((nh == 1) & ((flags & FLAG_SECONDARY) == 0))
    drop = (flags & FLAG_UNMAPPED) != 0
    if not secondary:
        drop |= (flags & FLAG_SECONDARY) != 0
        if not non_unique_primary:
            drop |= multiple
    keep = np.nonzero(~drop)[0]
    if keep.size == 0:
        return keep
    # admission order: (seq, start) groups in file order; primaries before secondaries inside a group (APG:416-418)
    sec = ((flags[keep] & FLAG_SECONDARY) != 0).astype(np.int64)
    order = np.lexsort((np.arange(keep.size), sec, batch.first[keep], batch.seq_idx[keep]))
    keep = keep[order]
    if max_alns_per_start > 0:
        key = np.stack([batch.seq_idx[keep].astype(np.int64), batch.first[keep].astype(np.int64), batch.rg[keep].astype(np.int64)], 1)
        # rank inside (seq, start, rg) in admission order
        o2 = np.lexsort((np.arange(keep.size), key[:, 2], key[:, 1], key[:, 0]))
        ks = key[o2]
        new = np.ones(keep.size, bool)
        new[1:] = (ks[1:] != ks[:-1]).any(1)
        grp_start = np.maximum.accumulate(np.where(new, np.arange(keep.size), 0))
        rank = np.arange(keep.size) - grp_start
        ok = np.zeros(keep.size, bool)
        ok[o2] = rank < max_alns_per_start
        keep = keep[ok]
    return keep


def pack_tile(batch: AlignmentBatch, idx: np.ndarray, ignore5: int = 0, ignore3: int = 0):
    """Packs the records idx (already in admission order, one sequence) -> (hdr, payload words, quality bytes, max_span)."""
    n = int(idx.size)
    hdr = np.zeros(n, HDR_DTYPE)
    if n == 0:
        return hdr, np.zeros(0, np.uint32), np.zeros(0, np.uint8), 1
    L = (batch.base_off[idx + 1] - batch.base_off[idx]).astype(np.int64)
    nc = (batch.cigar_off[idx + 1] - batch.cigar_off[idx]).astype(np.int64)
    if (L > 0xffff).any() or (nc > 0xff).any():
        raise ValueError("read length / CIGAR size exceeds the packed header fields")
    neg = (batch.flags[idx].astype(np.int64) & FLAG_REVERSE) != 0
    # per-base codes
    tot = int(L.sum())
    b_off = np.zeros(n + 1, np.int64); np.cumsum(L, out=b_off[1:])
    rep = np.repeat(np.arange(n), L)
    within = np.arange(tot, dtype=np.int64) - b_off[rep]
    src = batch.base_off[idx][rep] + within
    codes = _CODE[batch.bases[src]]
    quals = batch.quals[src]
    is_n = codes > 3
    has_n = np.zeros(n, bool)
    np.logical_or.at(has_n, rep, is_n)
    w_b2 = (L + 15) // 16
    w_nm = np.where(has_n, (L + 31) // 32, 0)
    w_lq = (L + 31) // 32          # low-quality mask
    words = nc + w_b2 + w_nm + w_lq + 1   # + qoff
    off = np.zeros(n + 1, np.int64); np.cumsum(words, out=off[1:])
    payload = np.zeros(int(off[-1]), np.uint32)
    # cigar
    c_tot = int(nc.sum())
    c_rep = np.repeat(np.arange(n), nc)
    c_off = np.zeros(n + 1, np.int64); np.cumsum(nc, out=c_off[1:])
    c_within = np.arange(c_tot, dtype=np.int64) - c_off[c_rep]
    cig = batch.cigar[batch.cigar_off[idx][c_rep] + c_within]
    payload[off[:-1][c_rep] + c_within] = cig
    # reference span from the cigar (ops consuming the reference: bit 0)
    span = np.zeros(n, np.int64)
    np.add.at(span, c_rep, np.where((cig & 1) != 0, cig >> 3, 0).astype(np.int64))
    # 2-bit bases
    b2_word = off[:-1][rep] + nc[rep] + within // 16
    np.add.at(payload, b2_word, ((codes & 3).astype(np.uint32) << ((within % 16) * 2).astype(np.uint32)))
    # N mask
    if is_n.any():
        sel = is_n
        nm_word = off[:-1][rep[sel]] + nc[rep[sel]] + w_b2[rep[sel]] + within[sel] // 32
        np.add.at(payload, nm_word, (np.uint32(1) << (within[sel] % 32).astype(np.uint32)))
    # low-quality mask: quality char - 33 <= 3 is never counted (CountsHelper.java:214)
    low = quals.astype(np.int64) - 33 <= 3
    if low.any():
        sel = low
        lq_word = off[:-1][rep[sel]] + nc[rep[sel]] + w_b2[rep[sel]] + w_nm[rep[sel]] + within[sel] // 32
        np.add.at(payload, lq_word, (np.uint32(1) << (within[sel] % 32).astype(np.uint32)))
    # quality bytes: their own array, each read padded to a multiple of 4; qoff = start / 4
    q_len = (L + 3) // 4 * 4
    q_off = np.zeros(n + 1, np.int64); np.cumsum(q_len, out=q_off[1:])
    qual = np.full(int(q_off[-1]), ord("!"), np.uint8)
    qual[q_off[:-1][rep] + within] = quals
    payload[off[:-1] + nc + w_b2 + w_nm + w_lq] = (q_off[:-1] // 4).astype(np.uint32)
    hdr["first"] = batch.first[idx]
    first_cig = np.zeros(n, np.int64)
    first_cig[nc > 0] = batch.cigar[batch.cigar_off[idx][nc > 0]]
    simple = (nc == 1) & ~has_n & (first_cig == ((L << 3) | 3))   # one match block as long as the read (NGSEP_M_SIMPLE)
    hdr["meta"] = (L | (nc << 16) | np.where(neg, M_NEG, 0) | np.where(has_n, M_NMASK, 0) | np.where(simple, M_SIMPLE, 0)).astype(np.uint32)
    # 5' is the alignment start for + strand reads and the alignment end for - strand reads (RA:613-644)
    hdr["ign_start"] = np.where(neg, ignore3, ignore5)
    hdr["ign_end"] = np.where(neg, ignore5, ignore3)
    hdr["off"] = off[:-1].astype(np.uint32)
    return hdr, payload, qual, int(max(1, span.max()))
