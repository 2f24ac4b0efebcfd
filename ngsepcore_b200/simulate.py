"""Seeded synthetic workload: planted-variant diploid genome + coordinate-sorted read ALIGNMENTS.

There is no aligner in the loop: reads are emitted directly as alignments, in the
decoded-record form htsjdk hands to ``ReadAlignmentFileReader.loadAlignment``
(reference: src/ngsep/alignments/io/ReadAlignmentFileReader.java:219-272), as struct-of-arrays numpy
buffers plus an optional SAM text rendering of the very same records.

Variant rates follow the reference's own simulator (src/ngsep/simulation/SingleIndividualSimulator.java:57-58,
277-377): SNV 1e-3/bp, indel 1e-4/bp, indel length U{1..10}, half deletions, het / hom-alt 1:1.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

YEAST_CONTIGS = [  # S288C-like extents (training/Saccharomyces_cerevisiae_STRs.txt max coordinates)
    ("chrI", 230218), ("chrII", 813184), ("chrIII", 316620), ("chrIV", 1531933), ("chrV", 576874),
    ("chrVI", 270161), ("chrVII", 1090940), ("chrVIII", 562643), ("chrIX", 439888), ("chrX", 745751),
    ("chrXI", 666816), ("chrXII", 1078177), ("chrXIII", 924431), ("chrXIV", 784333), ("chrXV", 1091291),
    ("chrXVI", 948066), ("chrmt", 85779),
]

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
OP_D, OP_I, OP_M, OP_S = 1, 2, 3, 6


@dataclass
class AlignmentBatch:
    """Decoded alignment records, coordinate-sorted, struct-of-arrays (one entry per record)."""
    contig_names: list
    contig_lengths: list
    seq_idx: np.ndarray      # int32
    first: np.ndarray        # int32, 1-based alignment start
    flags: np.ndarray        # uint16 SAM flags
    mapq: np.ndarray         # uint8
    rg: np.ndarray           # int16 read-group index (-1 = no RG tag)
    nh: np.ndarray           # int16 NH tag (-1 = absent)
    cigar_off: np.ndarray    # int64 [n+1] offsets into cigar
    cigar: np.ndarray        # uint32 len*8+op (H0 D1 I2 M3 P4 N5 S6 X7)
    base_off: np.ndarray     # int64 [n+1] offsets into bases/quals
    bases: np.ndarray        # uint8 ASCII
    quals: np.ndarray        # uint8 phred+33
    rg_names: list = field(default_factory=list)
    rg_samples: list = field(default_factory=list)

    @property
    def n(self) -> int:
        return int(self.first.shape[0])

    def select(self, idx: np.ndarray) -> "AlignmentBatch":
        idx = np.asarray(idx, dtype=np.int64)
        cl = (self.cigar_off[idx + 1] - self.cigar_off[idx]).astype(np.int64)
        bl = (self.base_off[idx + 1] - self.base_off[idx]).astype(np.int64)
        co = np.zeros(len(idx) + 1, np.int64); np.cumsum(cl, out=co[1:])
        bo = np.zeros(len(idx) + 1, np.int64); np.cumsum(bl, out=bo[1:])
        cg = _gather_ranges(self.cigar, self.cigar_off[idx], cl)
        bs = _gather_ranges(self.bases, self.base_off[idx], bl)
        qs = _gather_ranges(self.quals, self.base_off[idx], bl)
        return AlignmentBatch(self.contig_names, self.contig_lengths, self.seq_idx[idx], self.first[idx], self.flags[idx],
                              self.mapq[idx], self.rg[idx], self.nh[idx], co, cg, bo, bs, qs, self.rg_names, self.rg_samples)


def _gather_ranges(arr: np.ndarray, starts: np.ndarray, lens: np.ndarray) -> np.ndarray:
    total = int(lens.sum())
    if total == 0:
        return arr[:0].copy()
    out_off = np.zeros(len(lens) + 1, np.int64); np.cumsum(lens, out=out_off[1:])
    rep = np.repeat(np.arange(len(lens)), lens)
    within = np.arange(total, dtype=np.int64) - out_off[rep]
    return arr[starts[rep] + within]


def make_genome(contigs, gc: float = 0.38, seed: int = 1001, n_frac: float = 0.0, lower_frac: float = 0.0):
    """iid bases with the given GC content; optional N runs / lower-case (soft-masked) runs."""
    rng = np.random.default_rng(seed)
    p = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
    genome = {}
    for name, length in contigs:
        seq = _ACGT[rng.choice(4, size=length, p=p)].copy()
        for frac, kind in ((n_frac, "N"), (lower_frac, "l")):
            if frac > 0 and length > 2000:
                nruns = max(1, int(length * frac / 500))
                starts = rng.integers(0, length - 500, size=nruns)
                for s in starts:
                    if kind == "N":
                        seq[s:s + 500] = ord("N")
                    else:
                        seq[s:s + 500] |= 0x20
        genome[name] = seq
    return genome


@dataclass
class Variant:
    pos: int        # 1-based; SNV position, or anchor (last reference base before the event) for indels
    kind: str       # "snv" | "ins" | "del"
    length: int     # indel length (0 for SNV)
    alt: bytes      # SNV: alt base; ins: inserted bases; del: b""
    hom: bool       # True = both haplotypes


def plant_variants(ref: np.ndarray, rng, snv_rate=1e-3, indel_rate=1e-4, max_indel=10):
    n = len(ref)
    out = []
    n_snv = rng.binomial(n, snv_rate)
    n_ind = rng.binomial(n, indel_rate)
    pos_snv = np.unique(rng.integers(30, max(31, n - 30), size=n_snv))
    pos_ind = np.unique(rng.integers(30, max(31, n - 30), size=n_ind))
    blocked = np.zeros(n + 2, bool)
    for p in pos_ind:
        length = int(rng.integers(1, max_indel + 1))
        is_del = rng.random() < 0.5
        lo, hi = p - 2, p + (length if is_del else 0) + 3
        if blocked[lo:hi + 12].any() or hi + 12 >= n:
            continue
        if (ref[p - 1:p + length + 1] == ord("N")).any():
            continue
        blocked[lo:hi + 12] = True
        alt = b"" if is_del else bytes(_ACGT[rng.integers(0, 4, size=length)])
        out.append(Variant(int(p), "del" if is_del else "ins", length, alt, bool(rng.random() < 0.5)))
    for p in pos_snv:
        if blocked[p - 1:p + 2].any():
            continue
        r = ref[p - 1] & 0xDF
        if r not in _ACGT:
            continue
        choices = _ACGT[_ACGT != r]
        out.append(Variant(int(p), "snv", 0, bytes([int(choices[rng.integers(0, 3)])]), bool(rng.random() < 0.5)))
    out.sort(key=lambda v: v.pos)
    return out


def _build_haplotype(ref: np.ndarray, variants, hap: int, rng):
    """Return (hap_seq uint8, hap2ref int32) where hap2ref is the 1-based reference position of every haplotype
    base, or 0 for inserted bases.  Heterozygous variants go to a random haplotype."""
    pieces, maps = [], []
    cur = 0  # 0-based next reference index to copy
    up = ref & 0xDF
    for v in variants:
        on = v.hom or v._hap == hap
        if not on:
            continue
        if v.kind == "snv":
            pieces.append(up[cur:v.pos - 1]); maps.append(np.arange(cur + 1, v.pos, dtype=np.int32))
            pieces.append(np.frombuffer(v.alt, np.uint8)); maps.append(np.array([v.pos], np.int32))
            cur = v.pos
        elif v.kind == "ins":
            pieces.append(up[cur:v.pos]); maps.append(np.arange(cur + 1, v.pos + 1, dtype=np.int32))
            pieces.append(np.frombuffer(v.alt, np.uint8)); maps.append(np.zeros(v.length, np.int32))
            cur = v.pos
        else:
            pieces.append(up[cur:v.pos]); maps.append(np.arange(cur + 1, v.pos + 1, dtype=np.int32))
            cur = v.pos + v.length
    pieces.append(up[cur:]); maps.append(np.arange(cur + 1, len(ref) + 1, dtype=np.int32))
    return np.concatenate(pieces), np.concatenate(maps)


def _cigar_from_map(m: np.ndarray, ungapped_end_bp: int):
    """CIGAR (list of (len,op)) + first ref position for a read whose bases map to reference positions m
    (0 = inserted).  Indels closer than ungapped_end_bp to a read end are absorbed ungapped (what short-read
    aligners do), which leaves mismatching ends for the realigner to fix or trim."""
    L = len(m)
    ops = []
    i = 0
    # leading / trailing inserted bases become soft clips
    lead = 0
    while lead < L and m[lead] == 0:
        lead += 1
    trail = 0
    while trail < L - lead and m[L - 1 - trail] == 0:
        trail += 1
    if lead:
        ops.append([lead, OP_S])
    i = lead
    end = L - trail
    first = int(m[i])
    prev = first - 1
    run = 0
    while i < end:
        if m[i] == 0:
            if run:
                ops.append([run, OP_M]); run = 0
            j = i
            while j < end and m[j] == 0:
                j += 1
            ops.append([j - i, OP_I])
            i = j
            continue
        if m[i] != prev + 1:
            if run:
                ops.append([run, OP_M]); run = 0
            ops.append([int(m[i] - prev - 1), OP_D])
        run += 1
        prev = int(m[i])
        i += 1
    if run:
        ops.append([run, OP_M])
    if trail:
        ops.append([trail, OP_S])
    # absorb indels that sit too close to a read end
    if ungapped_end_bp > 0:
        body = [o for o in ops if o[1] != OP_S]
        if len(body) >= 3 and body[0][1] == OP_M and body[0][0] < ungapped_end_bp and body[1][1] in (OP_I, OP_D):
            k, (il, iop) = body[0][0], body[1]
            # ungapped: the first k bases slide next to the main block
            if iop == OP_I:
                # the k prefix bases + il inserted bases all become aligned bases before the main block
                first = first + k - (k + il)
                body = [[k + il + body[2][0], OP_M]] + body[3:]
            else:
                first = first + k + il - k
                body = [[k + body[2][0], OP_M]] + body[3:]
        if len(body) >= 3 and body[-1][1] == OP_M and body[-1][0] < ungapped_end_bp and body[-2][1] in (OP_I, OP_D):
            k, (il, iop) = body[-1][0], body[-2]
            if iop == OP_I:
                body = body[:-3] + [[body[-3][0] + il + k, OP_M]]
            else:
                body = body[:-3] + [[body[-3][0] + k, OP_M]]
        ops = ([[lead, OP_S]] if lead else []) + body + ([[trail, OP_S]] if trail else [])
    # merge adjacent equal ops
    merged = []
    for l, o in ops:
        if merged and merged[-1][1] == o:
            merged[-1][0] += l
        else:
            merged.append([l, o])
    return first, merged


def simulate(contigs=None, depth: float = 30.0, read_len: int = 100, seeds=(1001, 2002, 3003), gc: float = 0.38,
             snv_rate: float = 1e-3, indel_rate: float = 1e-4, error_rate: float = 0.002, paired: bool = True,
             frag_mean: float = 350.0, frag_sd: float = 35.0, low_mapq_frac: float = 0.02, n_frac: float = 0.0,
             lower_frac: float = 0.0, indel_jitter: float = 0.3, ungapped_end_bp: int = 6, n_read_groups: int = 1,
             genome=None, starts_per_pos_cap: int | None = None, variants_override: dict | None = None,
             tag_sites: dict | None = None, qual_vals=None):
    """Returns (genome dict name->uint8 ASCII, AlignmentBatch, truth dict name->[Variant]).

    variants_override: name -> [Variant] (each with ``_hap`` set) used instead of planting (population runs).
    tag_sites: name -> (ref_starts int64[], neg bool[]): GBS-style single-end reads, Poisson(depth) reads per tag, all
    starting at the tag (README.txt:588-596 of the reference describes this data type)."""
    if contigs is None:
        contigs = YEAST_CONTIGS
    if genome is None:
        genome = make_genome(contigs, gc=gc, seed=seeds[0], n_frac=n_frac, lower_frac=lower_frac)
    rng_v = np.random.default_rng(seeds[1])
    rng_r = np.random.default_rng(seeds[2])
    if qual_vals is None:
        qual_vals = np.array([40, 30, 20, 10, 2], np.uint8) + 33
        qual_p = np.array([0.70, 0.20, 0.07, 0.02, 0.01])
    else:       # another quality alphabet (phred values): the higher the value the more frequent, 1 % at the lowest one
        qv = np.sort(np.array(qual_vals, np.int64))
        w = np.arange(1, len(qv) + 1, dtype=np.float64) ** 2
        w[0] = 0.01 * w.sum()
        qual_vals, qual_p = (qv + 33).astype(np.uint8), w / w.sum()
    truth = {}
    per_contig = []
    for ci, (name, length) in enumerate(contigs):
        ref = genome[name]
        if variants_override is not None:
            variants = variants_override[name]
        else:
            variants = plant_variants(ref, rng_v, snv_rate, indel_rate)
            for v in variants:
                v._hap = int(rng_v.integers(0, 2))
        truth[name] = variants
        haps = [_build_haplotype(ref, variants, h, rng_v) for h in (0, 1)]
        n_reads = int(depth * length / read_len)
        if paired:
            n_frag = n_reads // 2
            hap_id = rng_r.integers(0, 2, size=n_frag)
            frag_len = np.clip(np.rint(rng_r.normal(frag_mean, frag_sd, size=n_frag)).astype(np.int64), read_len, None)
            hlen = np.array([len(haps[0][0]), len(haps[1][0])])[hap_id]
            frag_len = np.minimum(frag_len, hlen)
            fstart = (rng_r.random(n_frag) * (hlen - frag_len + 1)).astype(np.int64)
            starts = np.concatenate([fstart, fstart + frag_len - read_len])
            hap_of = np.concatenate([hap_id, hap_id])
            swap = rng_r.random(n_frag) < 0.5  # which mate is read 1
            fl_fwd = np.where(swap, 0x1 | 0x2 | 0x20 | 0x40, 0x1 | 0x2 | 0x20 | 0x80)
            fl_rev = np.where(swap, 0x1 | 0x2 | 0x10 | 0x80, 0x1 | 0x2 | 0x10 | 0x40)
            flags = np.concatenate([fl_fwd, fl_rev]).astype(np.uint16)
        elif tag_sites is not None:
            t_start, t_neg = tag_sites[name]
            per_tag = rng_r.poisson(depth, size=len(t_start))
            tag_of = np.repeat(np.arange(len(t_start)), per_tag)
            hap_of = rng_r.integers(0, 2, size=len(tag_of))
            starts = np.zeros(len(tag_of), np.int64)
            for h in (0, 1):
                seq_h, mp_h = haps[h]
                ff = np.maximum.accumulate(mp_h)          # inserted bases take the position before them
                hs = np.searchsorted(ff, t_start, side="left")
                hs = np.minimum(hs, len(seq_h) - read_len)
                sel = hap_of == h
                starts[sel] = hs[tag_of[sel]]
            flags = np.where(t_neg[tag_of], 0x10, 0).astype(np.uint16)
        else:
            hap_of = rng_r.integers(0, 2, size=n_reads)
            hlen = np.array([len(haps[0][0]), len(haps[1][0])])[hap_of]
            starts = (rng_r.random(n_reads) * (hlen - read_len + 1)).astype(np.int64)
            flags = np.where(rng_r.random(n_reads) < 0.5, 0, 0x10).astype(np.uint16)
        n = len(starts)
        bases = np.empty((n, read_len), np.uint8)
        first = np.zeros(n, np.int32)
        simple = np.zeros(n, bool)
        cigars = [None] * n
        for h in (0, 1):
            seq, mp = haps[h]
            sel = np.nonzero(hap_of == h)[0]
            st = starts[sel]
            idx = st[:, None] + np.arange(read_len)[None, :]
            bases[sel] = seq[idx]
            m0, m1 = mp[st], mp[st + read_len - 1]
            # simple = contiguous reference interval with no inserted bases inside
            ins_cum = np.concatenate([[0], np.cumsum(mp == 0)])
            has_ins = (ins_cum[st + read_len] - ins_cum[st]) > 0
            ok = (~has_ins) & (m1 - m0 == read_len - 1)
            simple[sel] = ok
            first[sel[ok]] = m0[ok]
            for k in sel[~ok]:
                f, ops = _cigar_from_map(mp[starts[k]:starts[k] + read_len], ungapped_end_bp)
                first[k] = f
                cigars[k] = ops
        # indel placement jitter inside repeats: shift right while the read sequence stays identical
        if indel_jitter > 0:
            up = ref & 0xDF
            for k in np.nonzero(~simple)[0]:
                ops = cigars[k]
                if len(ops) == 3 and ops[0][1] == OP_M and ops[2][1] == OP_M and ops[1][1] in (OP_I, OP_D) and rng_r.random() < indel_jitter:
                    il, iop = ops[1]
                    anchor = first[k] + ops[0][0] - 1  # 1-based last ref base before the event
                    rd = bases[k]
                    rpos = ops[0][0]
                    shift = 0
                    while shift < ops[2][0] - 3:
                        if iop == OP_D:
                            a = anchor + shift  # moving right: ref[a+1] must equal ref[a+il+1]
                            if a + il >= len(up) or up[a] != up[a + il]:
                                break
                        else:
                            if rd[rpos + shift] != rd[rpos + shift + il]:
                                break
                        shift += 1
                    if shift > 0:
                        s = int(rng_r.integers(0, shift + 1))
                        ops[0][0] += s
                        ops[2][0] -= s
        # sequencing errors + qualities
        err = rng_r.random((n, read_len)) < error_rate
        ne = int(err.sum())
        if ne:
            sub = rng_r.integers(1, 4, size=ne)
            cur = bases[err]
            lut = np.full(256, 255, np.uint8); lut[_ACGT] = np.arange(4, dtype=np.uint8)
            ci4 = lut[cur]
            okb = ci4 < 4
            new = cur.copy()
            new[okb] = _ACGT[(ci4[okb] + sub[okb]) % 4]
            bases[err] = new
        quals = qual_vals[rng_r.choice(len(qual_vals), size=(n, read_len), p=qual_p)]
        mapq = np.where(rng_r.random(n) < low_mapq_frac, 10, 60).astype(np.uint8)
        rg = (rng_r.integers(0, n_read_groups, size=n)).astype(np.int16) if n_read_groups > 0 else np.full(n, -1, np.int16)
        order = np.argsort(first, kind="stable")
        per_contig.append((ci, first[order], flags[order], mapq[order], rg[order], bases[order], quals[order],
                           [cigars[i] for i in order], simple[order]))
    # assemble SoA
    tot = sum(len(p[1]) for p in per_contig)
    seq_idx = np.concatenate([np.full(len(p[1]), p[0], np.int32) for p in per_contig]) if tot else np.zeros(0, np.int32)
    first = np.concatenate([p[1] for p in per_contig])
    flags = np.concatenate([p[2] for p in per_contig])
    mapq = np.concatenate([p[3] for p in per_contig])
    rg = np.concatenate([p[4] for p in per_contig])
    bases = np.concatenate([p[5] for p in per_contig]).reshape(-1)
    quals = np.concatenate([p[6] for p in per_contig]).reshape(-1)
    base_off = np.arange(tot + 1, dtype=np.int64) * read_len
    ncig = np.ones(tot, np.int64)
    all_cig = []
    i = 0
    for p in per_contig:
        for ops, s in zip(p[7], p[8]):
            if not s:
                ncig[i] = len(ops)
            i += 1
    cigar_off = np.zeros(tot + 1, np.int64); np.cumsum(ncig, out=cigar_off[1:])
    cigar = np.full(int(cigar_off[-1]), read_len * 8 + OP_M, np.uint32)
    i = 0
    for p in per_contig:
        for ops, s in zip(p[7], p[8]):
            if not s:
                o = cigar_off[i]
                for j, (l, op) in enumerate(ops):
                    cigar[o + j] = l * 8 + op
            i += 1
    names = [c[0] for c in contigs]
    lens = [c[1] for c in contigs]
    rg_names = [f"RG{i}" for i in range(max(n_read_groups, 0))]
    batch = AlignmentBatch(names, lens, seq_idx, first, flags, mapq, rg, np.full(tot, -1, np.int16), cigar_off, cigar,
                           base_off, bases, quals, rg_names, list(rg_names))
    return genome, batch, truth


_POP = {}


def _population_sample(smp: int):
    """Reads of one sample of simulate_population (module state _POP: inherited by forked workers, not pickled)."""
    P = _POP
    _, b, _ = simulate(contigs=P["contigs"], depth=P["depth"], read_len=P["read_len"], seeds=(P["seeds"][0], P["seeds"][1], P["seeds"][2] + smp),
                       paired=False, genome=P["genome"], variants_override=P["ovs"][smp], tag_sites=P["tags"], n_read_groups=1, **P["kw"])
    rng_s = np.random.default_rng(P["seeds"][2] + 7919 * (smp + 1))
    rg = (smp * P["rgs_per_sample"] + rng_s.integers(0, P["rgs_per_sample"], size=b.n)).astype(np.int16)
    if P["no_rg_frac"] > 0:
        rg[rng_s.random(b.n) < P["no_rg_frac"]] = -1
    b.rg = rg
    return b


def simulate_population(contigs, n_samples: int = 8, depth: float = 10.0, mode: str = "wgs", n_tags: int = 200,
                        seeds=(1021, 2022, 3023), read_len: int = 100, rgs_per_sample: int = 1, no_rg_frac: float = 0.0,
                        snv_rate: float = 1e-3, indel_rate: float = 1e-4, multi_allelic_frac: float = 0.1, workers: int = 1, n_tags_total: int | None = None, **kw):
    """Population workload for MultisampleVariantsDetector (BASELINE configs[3]): shared planted sites, per-sample
    genotypes drawn from a population allele frequency ~ Beta(0.2, 0.2); mode "wgs" = uniform single-end reads at `depth`
    per sample, "gbs" = `n_tags` tag loci per contig with Poisson(depth) identical-start reads per sample and tag.
    All samples' records are returned as ONE coordinate-sorted batch (one SAM/BAM with one @RG per sample and read group:
    the single-file input of MSVD, APG:216-219); ties keep sample order.  Returns (genome, batch, truth)."""
    genome = make_genome(contigs, gc=kw.pop("gc", 0.38), seed=seeds[0], n_frac=kw.pop("n_frac", 0.0), lower_frac=kw.pop("lower_frac", 0.0))
    rng_v = np.random.default_rng(seeds[1])
    shared, af, tags = {}, {}, {}
    for name, length in contigs:
        shared[name] = plant_variants(genome[name], rng_v, snv_rate, indel_rate)
        af[name] = rng_v.beta(0.2, 0.2, size=len(shared[name]))
        for v in shared[name]:        # a fraction of the SNV sites segregates a second alternative allele
            v._alt2 = None
            if v.kind == "snv" and rng_v.random() < multi_allelic_frac:
                r = genome[name][v.pos - 1] & 0xDF
                others = [int(x) for x in _ACGT if x != r and x != v.alt[0]]
                v._alt2 = bytes([others[int(rng_v.integers(0, 2))]])
        if mode == "gbs":
            nt = n_tags if n_tags_total is None else max(1, int(round(n_tags_total * length / sum(l for _, l in contigs))))   # tags in proportion to the sequence length
            st = np.sort(rng_v.integers(1, max(2, length - read_len - 20), size=nt)).astype(np.int64)
            tags[name] = (st, rng_v.random(nt) < 0.5)
    # per-sample genotypes first (one sequential stream of draws), then the samples' reads — independent given their own
    # seeds, so `workers` > 1 simulates them in forked processes with the same result
    ovs = []
    for smp in range(n_samples):
        ov = {}
        for name, _ in contigs:
            vs = []
            for v, f in zip(shared[name], af[name]):
                a0, a1 = rng_v.random() < f, rng_v.random() < f
                if not (a0 or a1):
                    continue
                if v._alt2 is not None:
                    x0 = v._alt2 if rng_v.random() < 0.4 else v.alt
                    x1 = v._alt2 if rng_v.random() < 0.4 else v.alt
                    if a0 and a1 and x0 != x1:      # the two haplotypes carry different alternative alleles
                        for hp, x in ((0, x0), (1, x1)):
                            w = Variant(v.pos, v.kind, v.length, x, False)
                            w._hap = hp
                            vs.append(w)
                        continue
                    w = Variant(v.pos, v.kind, v.length, x0 if a0 else x1, bool(a0 and a1))
                else:
                    w = Variant(v.pos, v.kind, v.length, v.alt, bool(a0 and a1))
                w._hap = 0 if a0 else 1
                vs.append(w)
            ov[name] = vs
        ovs.append(ov)
    _POP.update(contigs=contigs, depth=depth, read_len=read_len, seeds=seeds, genome=genome, ovs=ovs, tags=tags if mode == "gbs" else None,
                rgs_per_sample=rgs_per_sample, no_rg_frac=no_rg_frac, kw=kw)
    if workers > 1 and n_samples > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(workers, n_samples)) as pool:
            batches = pool.map(_population_sample, range(n_samples), chunksize=1)
    else:
        batches = [_population_sample(smp) for smp in range(n_samples)]
    _POP.clear()
    names = [c[0] for c in contigs]
    lens = [c[1] for c in contigs]
    seq_idx = np.concatenate([b.seq_idx for b in batches])
    first = np.concatenate([b.first for b in batches])
    order = np.lexsort((first, seq_idx))      # stable: ties keep sample, then in-sample order
    cl = np.concatenate([np.diff(b.cigar_off) for b in batches])
    bl = np.concatenate([np.diff(b.base_off) for b in batches])
    c_start = np.concatenate([[0], np.cumsum(cl)])[:-1]
    b_start = np.concatenate([[0], np.cumsum(bl)])[:-1]
    cigar = np.concatenate([b.cigar for b in batches])
    bases = np.concatenate([b.bases for b in batches])
    quals = np.concatenate([b.quals for b in batches])
    cl_o, bl_o = cl[order], bl[order]
    co = np.zeros(len(order) + 1, np.int64); np.cumsum(cl_o, out=co[1:])
    bo = np.zeros(len(order) + 1, np.int64); np.cumsum(bl_o, out=bo[1:])
    rg_names = [f"RG{i}" for i in range(n_samples * rgs_per_sample)]
    rg_samples = [f"S{i // rgs_per_sample:03d}" for i in range(n_samples * rgs_per_sample)]
    batch = AlignmentBatch(names, lens, seq_idx[order], first[order], np.concatenate([b.flags for b in batches])[order],
                           np.concatenate([b.mapq for b in batches])[order], np.concatenate([b.rg for b in batches])[order],
                           np.full(len(order), -1, np.int16), co, _gather_ranges(cigar, c_start[order], cl_o), bo,
                           _gather_ranges(bases, b_start[order], bl_o), _gather_ranges(quals, b_start[order], bl_o), rg_names, rg_samples)
    return genome, batch, {"variants": shared, "af": af}


_OPC = "HDIMPNSX"


def write_fasta(path: str, genome: dict, width: int = 80) -> None:
    with open(path, "wb") as f:
        for name, seq in genome.items():
            f.write(b">" + name.encode() + b"\n")
            n = len(seq)
            full = (n // width) * width
            if full:
                body = np.empty((full // width, width + 1), np.uint8)
                body[:, :width] = seq[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if n > full:
                f.write(seq[full:].tobytes() + b"\n")


def write_sam(path: str, batch: AlignmentBatch, with_rg: bool = True) -> None:
    """SAM text of exactly the batch's records (htsjdk reads SAM directly: RAFR:177-179)."""
    with open(path, "w") as f:
        f.write("@HD\tVN:1.6\tSO:coordinate\n")
        for n, l in zip(batch.contig_names, batch.contig_lengths):
            f.write(f"@SQ\tSN:{n}\tLN:{l}\n")
        if with_rg:
            for r, s in zip(batch.rg_names, batch.rg_samples):
                f.write(f"@RG\tID:{r}\tSM:{s}\n")
        bases = batch.bases.tobytes()
        quals = batch.quals.tobytes()
        co, bo = batch.cigar_off, batch.base_off
        lines = []
        for i in range(batch.n):
            cg = "".join(f"{int(v) >> 3}{_OPC[int(v) & 7]}" for v in batch.cigar[co[i]:co[i + 1]])
            tags = ""
            if with_rg and batch.rg[i] >= 0:
                tags += f"\tRG:Z:{batch.rg_names[batch.rg[i]]}"
            if batch.nh[i] >= 0:
                tags += f"\tNH:i:{int(batch.nh[i])}"
            lines.append(f"r{i}\t{int(batch.flags[i])}\t{batch.contig_names[batch.seq_idx[i]]}\t{int(batch.first[i])}\t"
                         f"{int(batch.mapq[i])}\t{cg}\t=\t0\t0\t{bases[bo[i]:bo[i + 1]].decode()}\t{quals[bo[i]:bo[i + 1]].decode()}{tags}\n")
            if len(lines) >= 65536:
                f.write("".join(lines)); lines = []
        f.write("".join(lines))
