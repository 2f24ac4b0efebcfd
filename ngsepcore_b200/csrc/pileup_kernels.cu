// sm_100a kernels of libngsep_pileup.so — see pileup_kernels.cuh for the design summary.
// Reference citations are relative to /root/reference/src/ngsep/ (RA = alignments/ReadAlignment.java,
// PR = discovery/PileupRecord.java, CH = discovery/CountsHelper.java, VD = discovery/VariantDiscoverySNVQAlgorithm.java,
// SSVPL = discovery/SingleSampleVariantPileupListener.java, APG = discovery/AlignmentsPileupGenerator.java).
#include "pileup_kernels.cuh"

#include <climits>
#include <cstdlib>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>

namespace ngsep_b200 {

// ------------------------------------------------------------------------------------------------------------
// shared-memory layout
// ------------------------------------------------------------------------------------------------------------
struct SmemLayout {
    unsigned bar, t0, kmin, misc, cnt, list, ulist, openq, chunk, wmax, rare;   // CTA-wide arrays
    unsigned stage0, stage_stride;                          // staged windows
    unsigned o_win, o_hdr, o_pay, o_ref2, o_refn, o_refl;   // offsets inside a stage
    unsigned total, max_reads;
};

// s_misc holds two sets of kMiscSet ints (one per window parity)
constexpr int kMiscSet = 16;
constexpr int M_NCAND = 0, M_UNC = 1, M_NRARE = 2, M_HASVER = 3, M_CARRY = 4, M_R1 = 5, M_PBASE = 6, M_FBCOV = 7, M_NOPEN = 8, M_NUNRES = 9;
constexpr int kSkipTable = 16;   // k_min[m] for m = 0..15 counted non-reference bases
constexpr int kMaxOpen = 32;     // open columns listed in shared memory per window (more go straight to the global list)
constexpr unsigned int kEntChunk = 1024u;   // entries a warp reserves in the global entry list at a time
// Resolving the records of an open column into (record, offset) entries inside the window kernel was built and measured in
// round 2 (profiles/README.md): the genotype kernel gains 0.013 ms per yeast pass, the window kernel loses 0.045 ms (a
// warp-serial piece of work in front of a block barrier).  The code stays behind this switch; the shipped kernels hand the
// open columns over as record ranges.
constexpr bool kWithEntries = false;
constexpr uint32_t kOpenSearch = 0xffffu;   // open-column entry without a record range: the genotype kernel searches it
// open-column entry, .w: either the number of records of the range that starts at record .z (0xffff = search), or — bit 31
// set — the number of (record, offset) entries the window kernel resolved for the column (bits 0-23; they start at entry .z
// of the entry list) with the reference allele in bits 24-25
constexpr uint32_t kOpenEntries = 0x80000000u, kOpenCountMask = 0x00ffffffu;
constexpr int kOpenRefShift = 24;
constexpr uint32_t kEntNeg = 1u << 18, kEntNotAcgt = 1u << 19;   // ColumnEntry.y: bits 0-15 read offset, 16-17 base, 18 strand, 19 base outside ACGT

// State only thread 0 touches (kept in shared memory so that it costs no registers in the other 255 threads).
struct Thread0State {
    WinRange dq[2];              // descriptors of window w and of window w + stride, read ahead of their use
    const unsigned char *ref2, *ref_n, *ref_l; const ngsep_read_hdr* hdr; const uint32_t* payload;   // fields of tile `tile`
    unsigned long long acc_cov, acc_cand;   // statistics of tile acc_tile, flushed when the tile changes
    int tile, T, acc_tile, prefetched;
};

static SmemLayout make_layout(int T, unsigned pay_cap_words, int n_stages) {
    SmemLayout L;
    L.max_reads = pay_cap_words / 12;   // a staged record has at least 12 payload words at the read lengths this is sized for
    unsigned o = 0;
    L.bar = o; o += 16;
    L.t0 = o; o += sizeof(Thread0State);
    L.kmin = o; o += kSkipTable * 4;
    L.misc = o; o += 2 * kMiscSet * 4;
    L.cnt = o; o += (T + 1) * 4;
    L.list = o; o += T * 4;
    L.ulist = o; o += ((T * 2 + 3) & ~3u);      // candidates the quarter-warp proofs left open (the sample axis keeps its candidate bitmap here)
    o = (o + 15u) & ~15u;
    L.openq = o; o += 2 * kMaxOpen * 16;        // open columns of the window (finished list entries), one list per window parity
    L.chunk = o; o += (kThreads / 32) * 8;      // per warp: the piece of the entry list it fills (next, end)
    o = (o + 7u) & ~7u;
    L.wmax = o; o += ((L.max_reads + kThreads) / 32 + 2) * 4;
    o = (o + 7u) & ~7u;
    L.rare = o; o += (L.max_reads + 1) * 8;
    o = (o + 127u) & ~127u;
    L.stage0 = o;
    unsigned s = 0;
    L.o_win = s; s += sizeof(WinRange);
    L.o_hdr = s; s += (L.max_reads + 1) * 16;
    L.o_pay = s; s += (pay_cap_words + 8) * 4;
    s = (s + 15u) & ~15u;
    L.o_ref2 = s; s += (T + 2 * kRefLeft) / 4;            // T/4 + 32 bytes
    L.o_refn = s; s += T / 8;
    L.o_refl = s; s += T / 8;
    s = (s + 127u) & ~127u;
    L.stage_stride = s;
    L.total = L.stage0 + (unsigned)n_stages * s;
    return L;
}

size_t pileup_smem_bytes(const LaunchCfg& cfg) { return make_layout(cfg.T_lay, cfg.pay_cap_words, cfg.n_stages).total; }

// ------------------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA 1-D bulk copy (cp.async.bulk -> SASS UBLKCP)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t addr = smem_u32(bar);
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(20000u)   // suspend-time hint: sleep in hardware instead of spinning
            : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// record views
// ------------------------------------------------------------------------------------------------------------
// L2 prefetch hint for a span a later bulk copy will read (16-byte aligned, a multiple of 16 bytes)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ uint32_t range_mask32(int lo, int hi) {  // bits [lo,hi), 0 <= lo < hi <= 32
    uint32_t m = (hi >= 32) ? 0xffffffffu : ((1u << hi) - 1u);
    return m & ~((1u << lo) - 1u);
}

struct RecView {
    int first, vfrom, vto, L, nc;
    uint32_t meta;
    const uint32_t* cig;
    const uint32_t* b2;
    const uint32_t* nm;
    const uint32_t* lowq;   // 1 bit/base: quality <= 3
    uint32_t qoff;          // qualities at byte 4*qoff of the tile's quality array
};
__device__ __forceinline__ RecView view_record(const ngsep_read_hdr& h, const uint32_t* pay) {
    RecView r;
    r.meta = h.meta;
    r.L = (int)(h.meta & 0xffffu);
    r.nc = (int)((h.meta >> 16) & 0xffu);
    r.first = h.first;
    r.vfrom = INT_MIN;
    r.vto = INT_MAX;
    const uint32_t* p = pay;
    if (h.meta & NGSEP_M_WINDOW) { r.vfrom = (int)p[0]; r.vto = (int)p[1]; r.first = (int)p[2]; p += 3; }
    r.cig = p;
    r.b2 = p + r.nc;
    const uint32_t* t = r.b2 + ((r.L + 15) >> 4);
    r.nm = t;
    if (h.meta & NGSEP_M_NMASK) t += (r.L + 31) >> 5;
    r.lowq = t;
    r.qoff = t[(r.L + 31) >> 5];
    return r;
}

__device__ __forceinline__ bool record_is_simple(const ngsep_read_hdr& h) {
    // the packer's promise (verified at staging): one match block as long as the read, no N mask, no valid window
    return (h.meta & NGSEP_M_SIMPLE) != 0u;
}

// ------------------------------------------------------------------------------------------------------------
// screen: positions of one window where a read base that may be counted differs from the reference
// ------------------------------------------------------------------------------------------------------------
struct WinCtx {
    int wbeg, g_lo, g_hi;        // 0-based index of the window start; [g_lo,g_hi) = the window inside tile and sequence
    const uint32_t* s_ref2;      // staged reference, base index wbeg - kRefLeft at bit 0 of word 0
    uint32_t* s_cnt;             // per position: flagged bases (an upper bound of the counted non-reference bases)
    uint32_t* s_list;            // candidate list: position in the window | record that flagged it first << kListShift
    int* mset;                   // counters of this window
    bool collect;                // false in EMIT_ALL mode (every position is a candidate already)
};
// candidate list entry as the screen writes it: bits 0-12 position in the window (windows hold at most 8192 positions),
// 13-31 the flagging record.  The resolve pass (C1) rewrites it in place for the proofs: 0-12 position, 13-14 reference
// allele, 15-20 `need` = k_min[m_up] (63 = no number of reference bases closes the column), 21-31 flagging record.
constexpr int kListShift = 13;
constexpr uint32_t kListPosMask = (1u << kListShift) - 1u;
constexpr int kListMaxRec = (1 << (32 - kListShift)) - 1;
constexpr int kProofRefShift = 13, kProofNeedShift = 15, kProofRecShift = 21;
constexpr int kProofMaxRec = (1 << (32 - kProofRecShift)) - 1, kProofNoNeed = 63;
constexpr int kProofRings = 8;   // rings of 2 x 32 records the last proof stage may add around the flagging record (deep columns)

// rec = index of the flagging record inside the window's record range
__device__ __forceinline__ void flag_candidate(const WinCtx& w, int idx, int rec) {
    const int li = idx - w.wbeg;
    if (atomicAdd(&w.s_cnt[li], 1u) == 0u)
        w.s_list[atomicAdd(&w.mset[M_NCAND], 1)] = (uint32_t)li | ((uint32_t)min(rec, kListMaxRec) << kListShift);
}

// 16 read bases (one payload word) against the reference at 0-based index `idx0` of the word's first base
__device__ __forceinline__ uint32_t ref16(const WinCtx& w, int idx0) {
    const int r = idx0 - w.wbeg + kRefLeft;   // >= 0: callers only ask for words that intersect [g_lo, g_hi)
    return __funnelshift_r(w.s_ref2[r >> 4], w.s_ref2[(r >> 4) + 1], (r & 15) * 2);
}

// Simple record.  Returns the last reference position (1-based) the record covers (PR.addAlignment, PR:154-156).
__device__ __forceinline__ int screen_simple(const ngsep_read_hdr& h, const uint32_t* pay, const WinCtx& w, int rec) {
    const int L = (int)(h.meta & 0xffffu);         // one match block of L bases
    const int ref0 = h.first - 1;                  // 0-based index of read offset 0
    const int a = max(ref0, w.g_lo), b = min(ref0 + L, w.g_hi);
    if (a < b && w.collect) {
        const uint32_t* b2 = pay + 1;
        const int j0 = (a - ref0) >> 4, j1 = (b - 1 - ref0) >> 4;       // payload words that intersect the window
        const int r0 = ref0 + 16 * j0 - w.wbeg + kRefLeft;
        const int sh = (r0 & 15) * 2;
        const uint32_t* rp = w.s_ref2 + (r0 >> 4);
        const int t_lo = a - ref0 - 16 * j0, t_hi = b - ref0 - 16 * j1;  // valid bases: [t_lo,16) of word j0, [0,t_hi) of word j1
        const uint32_t head = 0xffffffffu << (2 * t_lo);
        const uint32_t tail = (t_hi >= 16) ? 0xffffffffu : ((1u << (2 * t_hi)) - 1u);
        // mismatch bits of payload word j (2 bits per base, low bit set when the base differs from the reference)
        auto diff = [&](int j) {
            const uint32_t x = b2[j] ^ __funnelshift_r(rp[j - j0], rp[j - j0 + 1], sh);
            uint32_t m = (x | (x >> 1)) & 0x55555555u;
            if (j == j0) m &= head;
            if (j == j1) m &= tail;
            return m;
        };
        // rare: some base of word j differs — flag the ones CountsHelper would count (callable, RA:782-784; q > 3, CH:214)
        auto flag_word = [&](int j, uint32_t m) {
            const uint32_t* lowq = b2 + ((L + 15) >> 4);   // simple record: the low-quality mask follows the bases
            const int c_lo = max((int)h.ign_start, 0), c_hi = L - max((int)h.ign_end, 0);
            while (m) {
                const int bit = __ffs(m) - 1;
                m &= m - 1;
                const int off = 16 * j + (bit >> 1);
                if (off >= c_lo && off < c_hi && !((lowq[off >> 5] >> (off & 31)) & 1u)) flag_candidate(w, ref0 + off, rec);
            }
        };
        if (j1 - j0 < 32) {
            // Branch-free pass over the words (a warp would otherwise take the mismatch path on almost every iteration,
            // since SOME lane has a mismatch), then the few words that differ.  First and last word carry the masks.
            // wm collects one bit per word, most recent word in bit 0 (word j0 + t ends up in bit n - 1 - t)
            uint32_t wlo = rp[0], whi = rp[1];
            uint32_t x = b2[j0] ^ __funnelshift_r(wlo, whi, sh);
            uint32_t m = (x | (x >> 1)) & 0x55555555u & head;
            if (j1 == j0) m &= tail;
            uint32_t wm = min(m, 1u);
#pragma unroll 4
            for (int j = j0 + 1; j < j1; j++) {
                wlo = whi;
                whi = rp[j - j0 + 1];
                x = b2[j] ^ __funnelshift_r(wlo, whi, sh);
                m = (x | (x >> 1)) & 0x55555555u;
                wm = (wm << 1) | min(m, 1u);
            }
            if (j1 > j0) {
                x = b2[j1] ^ __funnelshift_r(whi, rp[j1 - j0 + 1], sh);
                m = (x | (x >> 1)) & 0x55555555u & tail;
                wm = (wm << 1) | min(m, 1u);
            }
            while (wm) {
                const int bit = 31 - __clz(wm);
                wm &= ~(1u << bit);
                const int j = j1 - bit;
                flag_word(j, diff(j));
            }
        } else {
            for (int j = j0; j <= j1; j++) { const uint32_t m = diff(j); if (m) flag_word(j, m); }
        }
    }
    return ref0 + L;
}

// General record (indels, clips, N mask, realigner versions): every aligned base that differs from the reference
// is flagged (a superset of the counted ones).  Returns the last covered position, or INT_MIN for a version record
// (its coverage goes through the bitmap fallback).
__device__ __forceinline__ int screen_general(const ngsep_read_hdr& h, const uint32_t* pay, const WinCtx& w, int rec) {
    const RecView r = view_record(h, pay);
    const int lo = max(w.g_lo, (r.vfrom == INT_MIN) ? INT_MIN : r.vfrom - 1);
    const int hi = min(w.g_hi, (r.vto == INT_MAX) ? INT_MAX : r.vto - 1);
    int ref_idx = r.first - 1, rpos = 0;
    for (int c = 0; c < r.nc; c++) {
        const uint32_t v = r.cig[c];
        const int len = (int)(v >> 3);
        if ((v & 3u) == 3u) {   // consumes reference and read: M(3) or X(7)
            const int a = max(ref_idx, lo), b = min(ref_idx + len, hi);
            if (a < b && w.collect) {
                const int delta = rpos - ref_idx;   // read offset = index + delta
                // walk 16 bases at a time on the READ grid
                for (int o = (a + delta) & ~15; o < b + delta; o += 16) {
                    const int wi = o >> 4;
                    uint32_t x = r.b2[wi] ^ ref16(w, o - delta);
                    uint32_t m = (x | (x >> 1)) & 0x55555555u;
                    const int o_lo = max(a + delta, o) - o, o_hi = min(b + delta, o + 16) - o;
                    m &= range_mask32(2 * o_lo, 2 * o_hi);
                    while (m) {
                        const int bit = __ffs(m) - 1;
                        m &= m - 1;
                        flag_candidate(w, o + (bit >> 1) - delta, rec);
                    }
                }
            }
            ref_idx += len;
            rpos += len;
        } else {
            if (v & 1u) ref_idx += len;
            if (v & 2u) rpos += len;
        }
    }
    if (h.meta & NGSEP_M_WINDOW) { w.mset[M_HASVER] = 1; return INT_MIN; }
    return ref_idx;   // 1-based last covered position (0-based exclusive end)
}

// Coverage interval [s,e] (1-based, inclusive) of any record: own start .. last reference position, cut by the valid window.
__device__ __forceinline__ void record_interval(const ngsep_read_hdr& h, const uint32_t* pay, int& s, int& e) {
    const RecView r = view_record(h, pay);
    int ref_len = 0;
    for (int c = 0; c < r.nc; c++) { const uint32_t v = r.cig[c]; if (v & 1u) ref_len += (int)(v >> 3); }
    s = r.first; e = r.first + ref_len - 1;
    if (r.vfrom != INT_MIN) s = max(s, r.vfrom);
    if (r.vto != INT_MAX) e = min(e, r.vto - 1);
}

// ------------------------------------------------------------------------------------------------------------
// exact: restatement of getAlleleCalls(1) + updateCounts for one (record, position)
// ------------------------------------------------------------------------------------------------------------
// Read offset aligned to reference position pos when that base is SNV evidence, else -1.
// Restates RA.getAlignedReadPosition (RA:842-871) and the skip rules of RA.updateAlleleCallsInfo (RA:782-795);
// an indel anchor has allele-call length > 1 and is dropped by PR.getAlleleCalls (PR:145).
__device__ __forceinline__ int resolve_offset(const RecView& r, int ign_s, int ign_e, int pos) {
    if (pos < r.first || pos < r.vfrom || pos >= r.vto) return -1;
    if (r.nc == 1 && (r.cig[0] & 7u) == 3u) {  // single match block
        int off = pos - r.first;
        if (off >= r.L || off >= (int)(r.cig[0] >> 3)) return -1;
        bool skip = off < ign_s || (r.L - off) <= ign_e;
        return skip ? -1 : off;
    }
    int ref_pos = r.first, read_pos = 0;
    bool prev_indel = false;
    for (int c = 0; c < r.nc; c++) {
        uint32_t v = r.cig[c];
        int len = (int)(v >> 3);
        uint32_t op = v & 7u;
        if ((v & 3u) == 3u) {
            if (pos < ref_pos) return -1;
            if (pos < ref_pos + len) {
                int j = pos - ref_pos;
                int rp = read_pos + j;
                if (rp < 0 || rp >= r.L) return -1;
                bool next_indel = false;
                if (c + 1 < r.nc) {
                    uint32_t nv = r.cig[c + 1];
                    uint32_t nop = nv & 7u;
                    next_indel = (nop == 1u || nop == 2u);
                }
                bool skip = rp < ign_s;
                skip = skip || (r.L - rp) <= ign_e;
                skip = skip || (prev_indel && j < 2);
                skip = skip || (next_indel && j < len - 1 && j >= len - 2);
                if (next_indel && j == len - 1) return -1;  // anchor: either skipped or an indel call (length > 1)
                return skip ? -1 : rp;
            }
        }
        if (v & 1u) ref_pos += len;
        if (v & 2u) read_pos += len;
        prev_indel = (op == 1u || op == 2u);
    }
    return -1;
}

// The tile's quality array: phred+33 bytes, or 2- / 4-bit codes into a small dictionary (ngsep_tile.qual_bits).
struct QualView {
    const uint8_t* __restrict__ p;
    uint32_t bits, d0, d1, d2, d3;
    __device__ __forceinline__ QualView(const TileDev& d) : p(d.qual), bits(d.qual_bits), d0(d.qual_dict[0]), d1(d.qual_dict[1]), d2(d.qual_dict[2]), d3(d.qual_dict[3]) {}
    __device__ __forceinline__ QualView(const uint8_t* q, uint32_t b, uint4 dict) : p(q), bits(b), d0(dict.x), d1(dict.y), d2(dict.z), d3(dict.w) {}
    // phred+33 byte of base `off` of the read whose run starts at 4-byte unit qoff
    __device__ __forceinline__ uint32_t at(uint32_t qoff, uint32_t off) const {
        if (bits == 8u) return p[(size_t)qoff * 4u + (size_t)off];
        const uint32_t bit = off * bits;
        const uint32_t code = ((uint32_t)p[(size_t)qoff * 4u + (size_t)(bit >> 3)] >> (bit & 7u)) & ((1u << bits) - 1u);
        const uint32_t w = (code < 4u) ? d0 : ((code < 8u) ? d1 : ((code < 12u) ? d2 : d3));
        return (w >> ((code & 3u) * 8u)) & 0xffu;
    }
};

// One lane = one record.  code: bits 0-1 allele, 2-6 quality, 7 negative strand, 8 counted (CH:223), 9 call returned
// by getAlleleCalls(1) (counts for totalCount, CH:210).
constexpr uint32_t kCodeNeg = 1u << 7, kCodeCounted = 1u << 8, kCodeCall = 1u << 9;

__device__ __forceinline__ uint32_t lane_eval(const ngsep_read_hdr& h, const uint32_t* pay, const QualView& qual, int pos, int max_qs) {
    uint32_t qc, bw;
    int off;
    if (h.meta & NGSEP_M_SIMPLE) {
        // one match block as long as the read: the offset needs no payload (RA:842-871), quality and bases load together
        const int L = (int)(h.meta & 0xffffu);
        off = pos - h.first;
        if (off < 0 || off >= L || off < (int)h.ign_start || (L - off) <= (int)h.ign_end) return 0u;   // RA:782-784
        const uint32_t* b2 = pay + 1;
        const uint32_t qoff = b2[((L + 15) >> 4) + ((L + 31) >> 5)];   // after the bases and the low-quality mask
        bw = b2[off >> 4];
        qc = qual.at(qoff, (uint32_t)off);
    } else {
        const RecView r = view_record(h, pay);
        off = resolve_offset(r, h.ign_start, h.ign_end, pos);
        if (off < 0) return 0u;
        qc = qual.at(r.qoff, (uint32_t)off);
        bw = r.b2[off >> 4];
        if ((r.meta & NGSEP_M_NMASK) && ((r.nm[off >> 5] >> (off & 31)) & 1u)) qc = 0u;   // allele not in ACGT: a call that is not counted (CH:220-221)
    }
    int q = (int)qc - 33;
    q = min(30, q);                                   // CH:91
    if (q <= 3) return kCodeCall;                     // CH:214-216 (and the N case above)
    if (q > max_qs) q = max_qs;                       // CH:217-219
    const uint32_t b = (bw >> ((off & 15) * 2)) & 3u;
    return kCodeCall | kCodeCounted | b | ((uint32_t)q << 2) | ((h.meta & NGSEP_M_NEG) ? kCodeNeg : 0u);
}

__device__ __forceinline__ int dev_phred(double p) {  // math/PhredScoreHelper.java:31-40
    if (p == 0.0) return 255;
    double score = -10.0 * log10(p);
    if (score > 255.0) return 255;
    double f = floor(score);
    return (int)((score - f >= 0.5) ? f + 1.0 : f);  // Math.round
}

constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------------------------
// SNVQ genotyping of one column, one warp per column (used by k_genotype_columns)
// ------------------------------------------------------------------------------------------------------------
// index of the ordered event (i,j) in CH.getPosteriorProbabilities' enumeration (CH:416-430): per first allele i, the
// homozygous event, then the heterozygous ones with j ascending
__device__ __forceinline__ int ev_index(int i, int j) { return i * 4 + ((i == j) ? 0 : ((j < i) ? j + 1 : j)); }

struct ColCounts { int c[4], sn[4], depth; };

// CH.getPosteriorProbabilities (CH:410-495) + VD.discoverSNV (VD:100-222) + the filter of SSVPL:223.
// `acc` is lane-owned: lanes 0-3 hold L[i][i], lanes 4-9 hold L[0][1], L[0][2], L[0][3], L[1][2], L[1][3], L[2][3];
// the 16 ordered events live on lanes 0-15 (lanes 16-31 mirror them).
__device__ __forceinline__ void finalize_warp(double acc, const ColCounts& cc, int lane, int pos, int ref, const TileDev& d, int min_quality,
                                              double lp_hom, double lp_het) {
    const int k = lane & 15;
    const int i = k >> 2, r = k & 3;
    const int j = (r == 0) ? i : ((r - 1 < i) ? r - 1 : r);
    const int lo = min(i, j), hi = max(i, j);
    const int owner = (i == j) ? i : ((lo == 0) ? 3 + hi : ((lo == 1) ? 5 + hi : 9));
    const double Lij = __shfl_sync(kFull, acc, owner);
    const double ev = Lij + ((i == j) ? lp_hom : lp_het);
    double mx = ev;   // CH:474-479: the quirk "starts at 1 and is replaced by the first event" reduces to the maximum
    for (int o = 8; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(kFull, mx, o));
    double e = ev - mx;
    e = (e < -20.0) ? 0.0 : exp10(e);
    double total = 0.0;
    for (int t = 0; t < 16; t++) total += __shfl_sync(kFull, e, t);   // CH:480-486, same order
    const double post = e / total;
    // VD.getIndexesMaxGenotype (VD:223-243), every lane redundantly
    const double p_ref = __shfl_sync(kFull, post, ev_index(ref, ref));
    int ii = ref, jj = ref;
    double pmax = p_ref;
    for (int a = 0; a < 4; a++)
        for (int b = a; b < 4; b++) {
            double gp = __shfl_sync(kFull, post, ev_index(a, b));
            if (a != b) gp += __shfl_sync(kFull, post, ev_index(b, a));
            if (gp > pmax + 0.01) { pmax = gp; ii = a; jj = b; }
        }
    double max_p = __shfl_sync(kFull, post, ev_index(ii, jj));
    if (ii != jj) max_p += __shfl_sync(kFull, post, ev_index(jj, ii));
    const double p_ii = __shfl_sync(kFull, post, ev_index(ii, ii));
    const double p_jj = __shfl_sync(kFull, post, ev_index(jj, jj));
    int kind, a0, a1;
    if (ii != jj && ii != ref && jj != ref) {  // VD:128-177
        kind = NGSEP_CALL_TRIALLELIC;
        if (p_ii > p_jj + 0.01) { a0 = ii; a1 = jj; } else { a0 = jj; a1 = ii; }
    } else if (ii != jj) {
        kind = NGSEP_CALL_HET; a0 = ref; a1 = (ref != ii) ? ii : jj;
    } else if (ii != ref) {
        kind = NGSEP_CALL_HOMALT; a0 = ii; a1 = ii;
    } else {
        return;  // homozygous reference: dropped by SSVPL:223
    }
    const int gq = dev_phred(1.0 - max_p);
    if (min_quality > gq) return;  // SSVPL:223
    unsigned int slot = 0;
    if (lane == 0) slot = atomicAdd(d.n_calls, 1u);
    slot = __shfl_sync(kFull, slot, 0);
    if (slot >= d.cap_calls) return;   // host sees n_calls > cap, grows the buffer and reruns
    ngsep_snv_call* o = &d.calls[slot];
    if (lane < 16) o->logcond[i * 4 + j] = Lij;
    if (lane == 16) {
        o->seq_idx = d.seq_idx; o->pos = pos;
        o->count[0] = cc.c[0]; o->count[1] = cc.c[1]; o->count[2] = cc.c[2]; o->count[3] = cc.c[3];
        o->depth = cc.depth;
        for (int b = 0; b < 4; b++) { o->strand[2 * b] = cc.sn[b]; o->strand[2 * b + 1] = cc.c[b] - cc.sn[b]; }
    }
    if (lane == 17) {
        o->post_ref = p_ref;
        o->post_max = max_p;
        o->gq = (int16_t)gq;
        o->qual = (int16_t)dev_phred(p_ref);
        o->a0 = (int8_t)a0; o->a1 = (int8_t)a1; o->ref = (int8_t)ref; o->kind = (int8_t)kind;
        o->sample = 0;
    }
}

// Records of the window as the exact phase sees them: the staged copy in shared memory, or (windows too deep to stage
// in one piece) the tile's arrays in global memory through L2.
template <bool kGlobal>
struct Recs {
    const ngsep_read_hdr* hdr;   // record 0 = first record of the window
    const uint32_t* pay;         // payload word 0 = word `pay_base` of the tile
    uint32_t pay_base;
    int n;
    __device__ __forceinline__ ngsep_read_hdr header(int i) const {
        int4 v;
        if (kGlobal) v = __ldg((const int4*)(hdr + i)); else v = *(const int4*)(hdr + i);
        ngsep_read_hdr h;
        h.first = v.x; h.meta = (uint32_t)v.y; h.ign_start = (int16_t)(v.z & 0xffff); h.ign_end = (int16_t)((uint32_t)v.z >> 16); h.off = (uint32_t)v.w;
        return h;
    }
    __device__ __forceinline__ int first(int i) const { return kGlobal ? __ldg(&hdr[i].first) : hdr[i].first; }
    __device__ __forceinline__ const uint32_t* payload(const ngsep_read_hdr& h) const { return pay + (h.off - pay_base); }
};

// first record index whose ordering start is >= want (warp-cooperative 32-ary search; every lane returns the result)
template <bool kGlobal>
__device__ __forceinline__ int warp_lower_bound(const Recs<kGlobal>& R, int want, int lane) {
    int lo = 0, hi = R.n;
    while (hi - lo > 32) {
        const int step = (hi - lo + 31) >> 5;
        const int idx = lo + lane * step;
        const bool lt = idx < hi && R.first(idx) < want;
        const int kk = __popc(__ballot_sync(kFull, lt));   // lanes 0..kk-1 probe records below `want` (monotone)
        if (kk == 0) { hi = lo; break; }
        const int nlo = lo + (kk - 1) * step + 1;
        hi = min(hi, lo + kk * step);
        lo = nlo;
    }
    if (hi > lo) {
        const int idx = lo + lane;
        const bool lt = idx < hi && R.first(idx) < want;
        lo += __popc(__ballot_sync(kFull, lt));
    }
    return lo;
}

struct ExactCtx {
    int wbeg, n_cand, n_unres, max_span, max_back, tile, rec0;   // rec0: index in the tile of the window's first record
    const uint32_t* s_list;
    const uint16_t* s_ulist;     // candidates the quarter-warp proofs left open (indexes into s_list)
    uint32_t* s_cnt;
    const uint32_t *s_ref2, *s_refn, *s_refl;
    const int* s_kmin;
    bool emit_all, ignore_lower, span_capped;
};

// Is record idx a simple record whose base at pos is a COUNTED reference base?  (offset = pos - first with one match
// block, RA:842-871; callable, RA:782-784; q > 3, CH:214.)
template <bool kGlobal>
__device__ __forceinline__ int counted_ref_base(const Recs<kGlobal>& R, int idx, int pos, int ref) {
    if (idx < 0 || idx >= R.n) return 0;
    if (kGlobal) {
        const ngsep_read_hdr h = R.header(idx);
        const int L = (int)(h.meta & 0xffffu), off = pos - h.first;
        if (!(h.meta & NGSEP_M_SIMPLE) || off < 0 || off >= L || off < (int)h.ign_start || (L - off) <= (int)h.ign_end) return 0;
        const uint32_t* b2 = R.payload(h) + 1;
        const uint32_t* lowq = b2 + ((L + 15) >> 4);
        return (!((lowq[off >> 5] >> (off & 31)) & 1u) && (int)((b2[off >> 4] >> ((off & 15) * 2)) & 3u) == ref) ? 1 : 0;
    } else {
        // staged window: the screen left the header in its proof form {first, first callable position, end of the callable
        // positions (0, 0 for a record that is not simple), b2 word index | lowq word offset << 16}
        const int4 dh = *(const int4*)(R.hdr + idx);
        if (pos < dh.y || pos >= dh.z) return 0;
        const int off = pos - dh.x;
        const uint32_t* b2 = R.pay + (dh.w & 0xffff);
        const uint32_t* lowq = b2 + ((uint32_t)dh.w >> 16);
        return (!((lowq[off >> 5] >> (off & 31)) & 1u) && (int)((b2[off >> 4] >> ((off & 15) * 2)) & 3u) == ref) ? 1 : 0;
    }
}

// The exact column (warp-cooperative, records through L2): every record of [i_lo, i_hi) in pileup order, one record per
// lane and round.  lane_eval restates PR.getAlleleCalls(1) + the quality rules of CH.updateCounts; ballots give the exact
// counts[4], countsStrand[4][2] and totalCount; columns whose exact counts prove the homozygous-reference call are
// dropped; the others get the 10 distinct log-likelihoods, one per lane, SUMMED IN PILEUP ORDER in fp64
// (CH:231-249) from the host-supplied tables — bit-identical to the Java doubles — and finalize_warp.
constexpr int kRegRounds = 4;
struct GenoCtx { const double* s_tab; const double* s_tab2; const int* s_kmin; int max_qs, min_quality; bool may_skip, emit_all; };
__device__ __forceinline__ void count_round(uint32_t code, ColCounts& cc) {
    cc.depth += __popc(__ballot_sync(kFull, (code & kCodeCall) != 0u));   // CH:210 totalCount: every call, whatever its quality or base
    const unsigned ng = __ballot_sync(kFull, (code & kCodeNeg) != 0u);
    const bool counted = (code & kCodeCounted) != 0u;
    for (int b = 0; b < 4; b++) {
        const unsigned mb = __ballot_sync(kFull, counted && (int)(code & 3u) == b);
        cc.c[b] += __popc(mb);
        cc.sn[b] += __popc(mb & ng);
    }
}
// The records of one round (one per lane), in lane order = pileup order: every log-likelihood lane adds its table value.
// The sums stay sequential per likelihood (bit-identical to the reference's loop, CH:231-249); what is parallel is the
// table index, computed once by the lane that owns the record.
__device__ __forceinline__ void accumulate_round(uint32_t code, int n_in_round, double& acc, int lane, const double* s_tab2) {
    const int row = (code & kCodeCounted) ? (int)(((code >> 2) & 31u) * 4u + (code & 3u)) : kTabZeroRow;
    const int col = min(lane, kTabLanes - 1);
    for (int k = 0; k < n_in_round; k++) acc += s_tab2[__shfl_sync(kFull, row, k) * kTabLanes + col];
}

// code_at(k) = the evidence code (lane_eval / entry_code) of the k-th record of the column, k in [0, n); 0 for k >= n
template <class CodeAt>
__device__ __forceinline__ void exact_column(CodeAt code_at, int n, const TileDev& d, const GenoCtx& g, int pos, int ref, bool ref_bad, int lane) {
    // lanes 0-3 own L[i][i]; lanes 4-9 own L[0][1], L[0][2], L[0][3], L[1][2], L[1][3], L[2][3] (see the table in k_genotype_columns)
    double acc = 0.0;
    ColCounts cc;
    for (int b = 0; b < 4; b++) { cc.c[b] = 0; cc.sn[b] = 0; }
    cc.depth = 0;
    const bool in_regs = n <= 32 * kRegRounds;
    uint32_t codes[kRegRounds];
    if (in_regs) {
        // every load of the column is issued before the first ballot
#pragma unroll
        for (int r = 0; r < kRegRounds; r++) codes[r] = code_at(32 * r + lane);
#pragma unroll
        for (int r = 0; r < kRegRounds; r++) count_round(codes[r], cc);
    } else {
        for (int base = 0; base < n; base += 32) {   // deep column: count and sum on the fly
            const uint32_t code = code_at(base + lane);
            count_round(code, cc);
            accumulate_round(code, min(32, n - base), acc, lane, g.s_tab2);
        }
    }
    if (g.emit_all && d.counts_dbg && lane == 0) {
        int32_t* o = d.counts_dbg + (int64_t)(pos - d.tile_first) * 5;
        o[0] = cc.c[0]; o[1] = cc.c[1]; o[2] = cc.c[2]; o[3] = cc.c[3]; o[4] = cc.depth;
    }
    if (ref_bad || cc.depth == 0) return;   // VD:101-107
    const int tot = cc.c[0] + cc.c[1] + cc.c[2] + cc.c[3];
    const int k_ref = (ref == 0) ? cc.c[0] : ((ref == 1) ? cc.c[1] : ((ref == 2) ? cc.c[2] : cc.c[3]));
    const int m_alt = tot - k_ref;
    // Proven hom-ref with the exact counts: VD.discoverSNV returns the homozygous-reference call that
    // SSVPL:223 drops.
    if (g.may_skip && (m_alt == 0 || (m_alt < kSkipTable && k_ref >= g.s_kmin[m_alt]))) return;
    if (in_regs) {
#pragma unroll
        for (int r = 0; r < kRegRounds; r++) accumulate_round(codes[r], min(32, max(0, n - 32 * r)), acc, lane, g.s_tab2);
    }
    finalize_warp(acc, cc, lane, pos, ref, d, g.min_quality, g.s_tab[93], g.s_tab[94]);
}

// ------------------------------------------------------------------------------------------------------------
// open columns: what the window kernel hands to k_genotype_columns
// ------------------------------------------------------------------------------------------------------------
// One record of an open column, resolved by the window kernel while the record sits in shared memory: everything
// PR.getAlleleCalls(1) + CH.updateCounts need except the quality byte itself.  Returns false when the record gives no
// call at pos (same rules as lane_eval).
__device__ __forceinline__ bool resolve_entry(const ngsep_read_hdr& h, const uint32_t* pay, int pos, uint2& ent) {
    uint32_t bw, qoff, extra = 0u;
    int off;
    if (h.meta & NGSEP_M_SIMPLE) {
        const int L = (int)(h.meta & 0xffffu);
        off = pos - h.first;
        if (off < 0 || off >= L || off < (int)h.ign_start || (L - off) <= (int)h.ign_end) return false;   // RA:782-784
        const uint32_t* b2 = pay + 1;
        qoff = b2[((L + 15) >> 4) + ((L + 31) >> 5)];
        bw = b2[off >> 4];
    } else {
        const RecView r = view_record(h, pay);
        off = resolve_offset(r, h.ign_start, h.ign_end, pos);
        if (off < 0) return false;
        qoff = r.qoff;
        bw = r.b2[off >> 4];
        if ((r.meta & NGSEP_M_NMASK) && ((r.nm[off >> 5] >> (off & 31)) & 1u)) extra = kEntNotAcgt;   // a call that is not counted (CH:220-221)
    }
    ent.x = qoff;
    ent.y = (uint32_t)off | (((bw >> ((off & 15) * 2)) & 3u) << 16) | ((h.meta & NGSEP_M_NEG) ? kEntNeg : 0u) | extra;
    return true;
}

// The quality rules of CH.updateCounts on a resolved entry (the tail of lane_eval): the only global load is the quality byte.
__device__ __forceinline__ uint32_t entry_code(const uint2 ent, const QualView& qual, int max_qs) {
    uint32_t qc = qual.at(ent.x, ent.y & 0xffffu);
    if (ent.y & kEntNotAcgt) qc = 0u;
    int q = (int)qc - 33;
    q = min(30, q);                                   // CH:91
    if (q <= 3) return kCodeCall;                     // CH:214-216
    if (q > max_qs) q = max_qs;                       // CH:217-219
    return kCodeCall | kCodeCounted | ((ent.y >> 16) & 3u) | ((uint32_t)q << 2) | ((ent.y & kEntNeg) ? kCodeNeg : 0u);
}

// Where the open columns of a window go: a per-window list in shared memory (published to the global list after the
// window, with ONE global atomic, off the critical path) and, for the resolved entries, the piece of the global entry list
// the warp has reserved (kEntChunk entries at a time, so that a global atomic with its round trip is rare).
struct OpenSink {
    uint4* s_openq;          // [kMaxOpen] list entries of this window
    int* n_open;             // shared counter
    uint2* s_chunk;          // this warp's piece of the entry list: x = next free entry, y = end
};

__device__ __forceinline__ void push_open(const OpenSink& sink, const ColumnQueue& cq, uint4 e) {   // one lane
    const int slot = atomicAdd(sink.n_open, 1);
    if (slot < kMaxOpen) sink.s_openq[slot] = e;
    else {   // rare: the window's list is full
        const unsigned int g = atomicAdd(&cq.ctr[2], 1u);
        if (g < cq.cap_open) cq.open[g] = e;
    }
}

// An open column of the window (warp-cooperative): its record range, then — windows staged in one piece — one entry per
// record that gives a call, in pileup order, so that k_genotype_columns only has to fetch quality bytes.  Columns of
// windows too deep for a stage, and columns that do not fit the entry list, are handed over as a record range.
// (The staged headers are in their proof form by now — only word 0, the ordering start, is intact — so the records' own
// headers come from the tile's array, `ghdr` = the header of the window's first record: an L2 hit, a few dozen per open column.)
template <bool kGlobal>
__device__ __forceinline__ void emit_open_column(const Recs<kGlobal>& R, const ColumnQueue& cq, const OpenSink& sink, const ngsep_read_hdr* __restrict__ ghdr, int tile, int rec0,
                                                 int pos, int ref, int max_span, int max_back, bool span_capped, int lane) {
    const int i_lo = warp_lower_bound(R, pos - max_span + 1, lane);
    const int i_hi = warp_lower_bound(R, pos + max_back + 1, lane);   // first record that starts too far to the right
    if (i_hi <= i_lo) return;
    const uint32_t n_range = (uint32_t)(i_hi - i_lo);
    bool as_range = !kWithEntries || kGlobal || span_capped || cq.cap_ent == 0u || n_range > kEntChunk;
    unsigned int ebase = 0u;
    if (!as_range) {
        if (lane == 0) {
            uint2 ch = *sink.s_chunk;
            if (ch.y - ch.x < n_range) {   // the warp's piece is used up: reserve the next one
                ch.x = atomicAdd(&cq.ctr[3], kEntChunk);
                ch.y = ch.x + kEntChunk;
                if (ch.x > cq.cap_ent || kEntChunk > cq.cap_ent - ch.x) ch.x = ch.y = 0xffffffffu;   // the list is full
            }
            ebase = ch.x;
            if (ch.x != 0xffffffffu) ch.x += n_range;
            *sink.s_chunk = ch;
        }
        ebase = __shfl_sync(kFull, ebase, 0);
        if (ebase == 0xffffffffu) as_range = true;   // the genotype kernel resolves the records itself
    }
    if (as_range) {
        if (lane == 0) push_open(sink, cq, make_uint4((unsigned int)tile, (unsigned int)pos, (unsigned int)(rec0 + i_lo), span_capped ? kOpenSearch : min(n_range, kOpenSearch)));
        return;
    }
    if (!kGlobal) {
        unsigned int n_ent = 0u;
        for (int base = i_lo; base < i_hi; base += 32) {
            const int idx = base + lane;
            uint2 ent = make_uint2(0u, 0u);
            bool ok = false;
            if (idx < i_hi) {
                const int4 v = __ldg((const int4*)(ghdr + idx));
                ngsep_read_hdr h;
                h.first = v.x; h.meta = (uint32_t)v.y; h.ign_start = (int16_t)(v.z & 0xffff); h.ign_end = (int16_t)((uint32_t)v.z >> 16); h.off = (uint32_t)v.w;
                ok = resolve_entry(h, R.payload(h), pos, ent);
            }
            const unsigned int m = __ballot_sync(kFull, ok);
            if (ok) cq.ent[ebase + n_ent + (unsigned int)__popc(m & ((1u << lane) - 1u))] = ent;
            n_ent += (unsigned int)__popc(m);
        }
        if (lane == 0 && n_ent > 0u) push_open(sink, cq, make_uint4((unsigned int)tile, (unsigned int)pos, ebase, kOpenEntries | ((uint32_t)ref << kOpenRefShift) | n_ent));
    }
}

// Phase C2.  First proof attempt by a QUARTER-WARP per candidate: the 8 records next to the flagging one (4 on each side)
// usually hold the k_min[1] = 7 counted reference bases that close a column with one flagged base.  What stays open is
// listed for the half-warp pass.
template <bool kGlobal>
__device__ __forceinline__ void quarter_proof_phase(const Recs<kGlobal>& R, const ExactCtx& x, uint16_t* s_ulist, int* n_unres, int warp, int lane, int nwarps) {
    const int grp = lane >> 3, gl = lane & 7;
    const int delta = (gl < 4) ? gl - 4 : gl - 3;   // -4..-1, +1..+4
    for (int base = 0; base < x.n_cand; base += 4 * nwarps) {
        const int ci = base + 4 * warp + grp;
        int ok = 0, need = 0;
        bool open = false;
        if (ci < x.n_cand) {
            const uint32_t ent = x.s_list[ci];
            need = (int)((ent >> kProofNeedShift) & 63u);
            open = need > 0;                               // need == 0: closed by the resolve pass
            if (open && need <= 8)
                ok = counted_ref_base(R, (int)(ent >> kProofRecShift) + delta, x.wbeg + (int)(ent & kListPosMask) + 1, (int)((ent >> kProofRefShift) & 3u));
        }
        const unsigned okm = __ballot_sync(kFull, ok != 0);
        if (open && __popc(okm & (0xffu << (8 * grp))) >= need) open = false;
        if (open && gl == 0) s_ulist[atomicAdd(n_unres, 1)] = (uint16_t)ci;
    }
}

// Phase C3.  Fast proof of hom-ref by a HALF-WARP per candidate the quarter-warp pass left open: at most m_up counted non-reference bases
// were flagged; the 16 records around the flagging one give a lower bound k_lb of the counted reference bases (any subset
// of the records does).  k_lb >= need = k_min[m_up] closes the column.  What stays open is listed, with its record range,
// for k_genotype_columns (emit_open_column).  Entries come from the resolve pass (position | reference allele | need | flagging record).
template <bool kGlobal>
__device__ __forceinline__ void proof_phase(const Recs<kGlobal>& R, const ExactCtx& x, const ColumnQueue& cq, const OpenSink& sink, const ngsep_read_hdr* __restrict__ ghdr,
                                            int warp, int lane, int nwarps) {
    const int grp = lane >> 4, gl = lane & 15;
    for (int base = 0; base < x.n_unres; base += 2 * nwarps) {
        const int ui = base + 2 * warp + grp;
        uint32_t ent = 0u;
        int ok = 0, need = 0;
        bool open = false;
        if (ui < x.n_unres) {
            ent = x.s_list[x.s_ulist[ui]];
            need = (int)((ent >> kProofNeedShift) & 63u);
            open = true;
            if (need <= 15)
                ok = counted_ref_base(R, (int)(ent >> kProofRecShift) - 8 + gl, x.wbeg + (int)(ent & kListPosMask) + 1, (int)((ent >> kProofRefShift) & 3u));
        }
        const unsigned okm = __ballot_sync(kFull, ok != 0);
        if (open && __popc(okm & (0xffffu << (16 * grp))) >= need) open = false;
        // the (few) columns still open: second chance with the whole warp (the 32 records around the flagging one), then
        // the record range of the column, then the list
        unsigned sm = __ballot_sync(kFull, open && gl == 0);
        while (sm) {
            const int src = __ffs(sm) - 1;
            sm &= sm - 1;
            const uint32_t e = __shfl_sync(kFull, ent, src);
            const int e_need = (int)((e >> kProofNeedShift) & 63u);
            const int e_li = (int)(e & kListPosMask), e_pos = x.wbeg + e_li + 1;
            if (e_need < kProofNoNeed) {
                // the 32 records around the flagging one, then — deep columns, whose k_min is beyond 32 — ring after ring of 32
                // records on either side (any subset of the records gives a lower bound of the counted reference bases)
                const int rc = (int)(e >> kProofRecShift), e_ref = (int)((e >> kProofRefShift) & 3u);
                int got = __popc(__ballot_sync(kFull, counted_ref_base(R, rc - 16 + lane, e_pos, e_ref) != 0));
                for (int ring = 1; got < e_need && ring < kProofRings && (rc - 16 - 32 * (ring - 1) > 0 || rc + 16 + 32 * (ring - 1) < R.n); ring++) {
                    const int before = got;
                    got += __popc(__ballot_sync(kFull, counted_ref_base(R, rc - 16 - 32 * ring + lane, e_pos, e_ref) != 0));
                    got += __popc(__ballot_sync(kFull, counted_ref_base(R, rc + 16 + 32 * (ring - 1) + lane, e_pos, e_ref) != 0));
                    if (got == before) break;   // the records this far from the flagging one no longer reach the position
                }
                if (got >= e_need) continue;
            }
            emit_open_column<kGlobal>(R, cq, sink, ghdr, x.tile, x.rec0, e_pos, (int)((e >> kProofRefShift) & 3u), x.max_span, x.max_back, x.span_capped, lane);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// the window kernel
// ------------------------------------------------------------------------------------------------------------
struct KArgs {
    const TileDev* tiles;
    const WinRange* win;
    int n_windows;
    const DevParams* prm;
    ColumnQueue cq;
    LaunchCfg cfg;
    SmemLayout lay;
};

template <int kThr, int kMinBlocks>
__global__ void __launch_bounds__(kThr, kMinBlocks) k_pileup_windows(const KArgs a) {
    constexpr int nthr = kThr, nwarps = kThr / 32;
    extern __shared__ __align__(128) unsigned char smem[];
    const SmemLayout& lay = a.lay;
    uint64_t* s_bar = (uint64_t*)(smem + lay.bar);
    int* s_kmin = (int*)(smem + lay.kmin);
    int* s_misc = (int*)(smem + lay.misc);
    uint32_t* s_cnt = (uint32_t*)(smem + lay.cnt);
    uint32_t* s_list = (uint32_t*)(smem + lay.list);
    uint16_t* s_ulist = (uint16_t*)(smem + lay.ulist);
    uint2* s_chunk = (uint2*)(smem + lay.chunk);
    int* s_wmax = (int*)(smem + lay.wmax);
    int2* s_rare = (int2*)(smem + lay.rare);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevParams& prm = *a.prm;
    const int T_lay = a.cfg.T_lay;
    if (tid < kSkipTable) s_kmin[tid] = prm.k_min[tid];
    if (tid < 2 * kMiscSet) s_misc[tid] = ((tid % kMiscSet) == M_CARRY) ? INT_MIN : 0;
    for (int i = tid; i < T_lay + 1; i += nthr) s_cnt[i] = 0u;
    if (tid < nwarps) s_chunk[tid] = make_uint2(0u, 0u);
    if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); }
    __syncthreads();
    uint32_t phase_bits = 0u;   // bit s = parity the next wait on stage s must see
    const bool emit_all = (prm.flags & NGSEP_F_EMIT_ALL) != 0;
    const bool ignore_lower = (prm.flags & NGSEP_F_IGNORE_LOWER_REF) != 0;
    const bool multi = prm.n_samples > 1;
    const int n_stages = a.cfg.n_stages;
    const int stride = (int)gridDim.x;
    // The thread that issues the bulk copies.  One stage: thread 0 (the copies of the next window start when the stage frees).
    // Two stages: lane 0 of the LAST warp issues the next window's copies right after the top barrier, while the other warps
    // are already screening the current window, so the copies overlap the whole window.
    const int issuer = (n_stages == 2) ? nthr - 32 : 0;

    auto fits = [&](const WinRange& wr) {
        return (wr.hi - wr.lo) <= (int)lay.max_reads && (wr.off_hi - (wr.off_lo & ~3u)) <= a.cfg.pay_cap_words;
    };

    // ---- issuer only: window descriptors are read ahead of their use, tile fields are cached while the tile stays -----------
    Thread0State& t0 = *(Thread0State*)(smem + lay.t0);
    // one mbarrier phase = the window descriptor + the reference window + (windows that fit a stage) the whole record range
    auto issue_window = [&](const WinRange& wr, int w_idx, int stage) {
        if (wr.tile != t0.tile) {
            const TileDev& d = a.tiles[wr.tile];
            t0.tile = wr.tile; t0.T = d.T;
            t0.ref2 = (const unsigned char*)d.ref2; t0.ref_n = (const unsigned char*)d.ref_n; t0.ref_l = (const unsigned char*)d.ref_l;
            t0.hdr = d.hdr; t0.payload = d.payload;
        }
        unsigned char* stg = smem + lay.stage0 + (unsigned)stage * lay.stage_stride;
        uint64_t* bar = &s_bar[stage];
        // the stage was written with generic stores (proof-form headers) by the previous window: order them before the
        // asynchronous-proxy writes of the bulk copies (the block barrier ordered the threads, this orders the proxies)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const int64_t wbeg = wr.wbeg;
        const uint32_t b_ref2 = (uint32_t)(t0.T + 2 * kRefLeft) / 4u, b_bits = (uint32_t)t0.T / 8u;
        const bool recs = fits(wr) && wr.hi > wr.lo;
        const uint32_t o0 = wr.off_lo & ~3u;
        const uint32_t hb = recs ? (uint32_t)(wr.hi - wr.lo + 1) * 16u : 0u;   // + the next record's header (sentinel)
        const uint32_t pb = recs ? (((wr.off_hi - o0) * 4u + 15u) & ~15u) : 0u;
        mbar_expect_tx(bar, (uint32_t)sizeof(WinRange) + b_ref2 + b_bits + (ignore_lower ? b_bits : 0u) + hb + pb);
        bulk_g2s(stg + lay.o_win, a.win + w_idx, (uint32_t)sizeof(WinRange), bar);
        bulk_g2s(stg + lay.o_ref2, t0.ref2 + (wbeg - kRefLeft) / 4, b_ref2, bar);
        bulk_g2s(stg + lay.o_refn, t0.ref_n + wbeg / 8, b_bits, bar);
        if (ignore_lower) bulk_g2s(stg + lay.o_refl, t0.ref_l + wbeg / 8, b_bits, bar);
        if (hb) bulk_g2s(stg + lay.o_hdr, t0.hdr + wr.lo, hb, bar);
        if (pb) bulk_g2s(stg + lay.o_pay, t0.payload + o0, pb, bar);
    };
    auto flush_stats = [&]() {
        if (t0.acc_tile >= 0) {
            const TileDev& dp = a.tiles[t0.acc_tile];
            if (t0.acc_cov) atomicAdd(&dp.counters[0], t0.acc_cov);
            if (t0.acc_cand) atomicAdd(&dp.counters[1], t0.acc_cand);
        }
    };

    int w = blockIdx.x;
    if (w >= a.n_windows) return;
    if (tid == issuer) {
        t0.tile = -1; t0.acc_tile = -1; t0.acc_cov = 0ull; t0.acc_cand = 0ull; t0.prefetched = 0;
        t0.dq[0] = a.win[w];
        t0.dq[1] = (w + stride < a.n_windows) ? a.win[w + stride] : t0.dq[0];
    }
    for (int it = 0; w < a.n_windows; it++) {
        const int cur = (n_stages == 2) ? (it & 1) : 0;
        const int wn = w + stride;
        if (tid == issuer) {
            if (n_stages == 2) {
                if (!t0.prefetched) issue_window(t0.dq[0], w, cur);   // first window of the CTA only
            } else {
                WinRange dq2;   // window w + 2 stride: requested now, consumed at the end of this block (its latency overlaps the issue)
                const bool have2 = wn + stride < a.n_windows;
                if (have2) dq2 = a.win[wn + stride];
                if (!t0.prefetched) issue_window(t0.dq[0], w, cur);
                t0.prefetched = 0;
                t0.dq[0] = t0.dq[1];
                if (have2) t0.dq[1] = dq2;
            }
        }
        unsigned char* stg = smem + lay.stage0 + (unsigned)cur * lay.stage_stride;
        const ngsep_read_hdr* s_hdr = (const ngsep_read_hdr*)(stg + lay.o_hdr);
        const uint32_t* s_pay = (const uint32_t*)(stg + lay.o_pay);
        int* mset = s_misc + (it & 1) * kMiscSet;

        if (warp == 0) mbar_wait(&s_bar[cur], (phase_bits >> cur) & 1u);   // one warp polls; the others park in the barrier
        phase_bits ^= 1u << cur;
        __syncthreads();

        if (n_stages == 2 && tid == issuer) {
            // the other stage has been free since the trailing barrier of the previous window: the next window's copies run
            // under this window's screen and proofs
            const bool have2 = wn + stride < a.n_windows;
            WinRange dq2;
            if (have2) dq2 = a.win[wn + stride];
            t0.prefetched = 0;
            if (wn < a.n_windows) { issue_window(t0.dq[1], wn, cur ^ 1); t0.prefetched = 1; }
            t0.dq[0] = t0.dq[1];
            if (have2) t0.dq[1] = dq2;
        }

        const WinRange wr = *(const WinRange*)(stg + lay.o_win);
        const int wbeg = wr.wbeg, g_lo = wr.wbeg + (int)(wr.g_pack & 0xffffu), g_hi = wr.wbeg + (int)(wr.g_pack >> 16);
        const int A = g_lo + 1, B = g_hi;   // 1-based inclusive positions of the window (empty when A > B)
        const int n_rec = wr.hi - wr.lo;
        const bool single = fits(wr);
        const int max_span = (int)(wr.span_pack & 0xffffu), max_back = (int)(wr.span_pack >> 16);
        const bool span_capped = max_span == 0xffff || max_back == 0xffff;   // the tile's real spans are larger: columns go out without a record range
        WinCtx wc;
        wc.wbeg = wbeg; wc.g_lo = g_lo; wc.g_hi = g_hi;
        wc.s_ref2 = (const uint32_t*)(stg + lay.o_ref2);
        wc.s_cnt = s_cnt; wc.s_list = s_list; wc.mset = mset; wc.collect = !emit_all;

        // ---- phase A: screen + coverage scan, one thread per record -------------------------------------------------
        // Coverage: with the records sorted by start s_i and reach_i = max(e_0..e_i) (e = last covered position), the
        // positions of (reach_i, s_{i+1}) are covered by nobody.  reach_i is at least the warp-local prefix maximum; only
        // when that does not reach s_{i+1} - 1 is the record queued for the exact cross-warp evaluation (process_rare).
        auto screen_chunk = [&](int n_chunk, uint32_t pay_base, int rec0, bool first_chunk) {
            for (int base = 0; base < n_chunk; base += nthr) {
                const int i = base + tid;
                const bool act = i < n_chunk;
                int e = INT_MIN, nxt = INT_MIN, floor_i = INT_MAX;
                if (act) {
                    const ngsep_read_hdr h = s_hdr[i];
                    const uint32_t* pay = s_pay + (h.off - pay_base);
                    e = record_is_simple(h) ? screen_simple(h, pay, wc, rec0 + i) : screen_general(h, pay, wc, rec0 + i);
                    {   // proof form of the staged header (counted_ref_base<false>); word 0 (the ordering start) stays
                        int4 dh = make_int4(h.first, 0, 0, 0);
                        if (record_is_simple(h)) {
                            const int L = (int)(h.meta & 0xffffu);
                            dh.y = h.first + max((int)h.ign_start, 0);
                            dh.z = h.first + L - max((int)h.ign_end, 0);
                            dh.w = (int)((uint32_t)(pay + 1 - s_pay) | ((uint32_t)((L + 15) >> 4) << 16));
                        }
                        *(int4*)(const_cast<ngsep_read_hdr*>(s_hdr) + i) = dh;
                    }
                    nxt = min(s_hdr[i + 1].first, B + 1);
                    floor_i = max(h.first - 1, A - 1);   // record i accounts for the positions of [s_i, s_{i+1}) only
                    if (first_chunk && i == 0) { const int lead = min(h.first, B + 1) - A; if (lead > 0) atomicAdd(&mset[M_UNC], lead); }
                }
                int run = e;
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(kFull, run, o); if (lane >= o) run = max(run, t); }
                if (lane == 31) s_wmax[(base >> 5) + warp] = run;
                if (act && nxt - 1 > max(run, floor_i)) s_rare[atomicAdd(&mset[M_NRARE], 1)] = make_int2(i, run);
            }
        };
        auto process_rare = [&](int carry) {
            const int n_rare = mset[M_NRARE];
            for (int k = tid; k < n_rare; k += nthr) {
                const int2 rr = s_rare[k];
                int reach = max(max(rr.y, A - 1), max(carry, s_hdr[rr.x].first - 1));
                for (int c = 0; c < (rr.x >> 5); c++) reach = max(reach, s_wmax[c]);
                const int gap = min(s_hdr[rr.x + 1].first, B + 1) - 1 - reach;
                if (gap > 0) atomicAdd(&mset[M_UNC], gap);
            }
        };

        if (single) {
            if (n_rec > 0) screen_chunk(n_rec, wr.off_lo & ~3u, 0, true);
            else if (tid == 0 && B >= A) mset[M_UNC] = B - A + 1;
        } else {
            // windows too deep for one stage: stream the record range chunk by chunk through the current stage
            const TileDev& d = a.tiles[wr.tile];
            int r0 = wr.lo;
            bool first_chunk = true;
            while (r0 < wr.hi) {
                if (tid == 0) {
                    const uint32_t base = d.hdr[r0].off & ~3u;
                    int lo_r = r0 + 1, hi_r = min(wr.hi, r0 + (int)lay.max_reads), r1 = r0;
                    while (lo_r <= hi_r) {
                        const int mid = (lo_r + hi_r) >> 1;
                        if (d.hdr[mid].off - base <= a.cfg.pay_cap_words) { r1 = mid; lo_r = mid + 1; } else hi_r = mid - 1;
                    }
                    if (r1 == r0) r1 = r0 + 1;   // cannot happen: staging rejects records larger than a stage
                    const uint32_t hb = (uint32_t)(r1 - r0 + 1) * 16u;
                    const uint32_t pb = ((d.hdr[r1].off - base) * 4u + 15u) & ~15u;
                    mset[M_R1] = r1;
                    mset[M_PBASE] = (int)base;
                    mbar_expect_tx(&s_bar[cur], hb + pb);
                    bulk_g2s(stg + lay.o_hdr, d.hdr + r0, hb, &s_bar[cur]);
                    if (pb) bulk_g2s(stg + lay.o_pay, d.payload + base, pb, &s_bar[cur]);
                }
                __syncthreads();
                const int r1 = mset[M_R1];
                const uint32_t pay_base = (uint32_t)mset[M_PBASE];
                const int carry = mset[M_CARRY];
                mbar_wait(&s_bar[cur], (phase_bits >> cur) & 1u);
                phase_bits ^= 1u << cur;
                screen_chunk(r1 - r0, pay_base, r0 - wr.lo, first_chunk);
                __syncthreads();
                process_rare(carry);
                __syncthreads();
                if (tid == 0) {
                    int c = carry;
                    for (int k = 0; k < ((r1 - r0 + 31) >> 5); k++) c = max(c, s_wmax[k]);
                    mset[M_CARRY] = c;
                    mset[M_NRARE] = 0;
                }
                first_chunk = false;
                r0 = r1;
            }
        }
        __syncthreads();   // ---- sync 1: candidates listed, cover deltas and low-quality counts complete ----------------------

        if (tid < kMiscSet) s_misc[((it & 1) ^ 1) * kMiscSet + tid] = (tid == M_CARRY) ? INT_MIN : 0;   // next window's counters
        if (single && n_rec > 0) process_rare(INT_MIN);
        const int n_cand = emit_all ? 0 : mset[M_NCAND];
        if (emit_all) {   // test hook: every position of the window is an open column
            const int n_all = (n_rec > 0) ? max(0, g_hi - g_lo) : 0;
            if (tid == 0 && n_all > 0) mset[M_R1] = (int)atomicAdd(&a.cq.ctr[2], (unsigned int)n_all);
            __syncthreads();
            const unsigned int gbase = (unsigned int)mset[M_R1];
            for (int i = tid; i < n_all; i += nthr)
                if (gbase + i < a.cq.cap_open) a.cq.open[gbase + i] = make_uint4((unsigned int)wr.tile, (unsigned int)(g_lo + i + 1), 0u, kOpenSearch);
        } else if (multi && prm.multi_ordered && n_cand > 0) {
            // Sample axis: no proof is attempted.  The window's candidates go to k_genotype_columns_multi IN POSITION ORDER
            // and as one contiguous piece of the list: a CTA there walks a run of neighbouring columns, whose records are
            // largely the same (L1 reuse).  Bitmap of the open positions -> ordered list -> record range per position.
            const uint32_t* s_refn = (const uint32_t*)(stg + lay.o_refn);
            const uint32_t* s_refl = (const uint32_t*)(stg + lay.o_refl);
            uint32_t* s_bm = (uint32_t*)s_ulist;
            const int nw = (T_lay + 31) >> 5;
            for (int i = tid; i < nw; i += nthr) s_bm[i] = 0u;
            __syncthreads();
            for (int ci = tid; ci < n_cand; ci += nthr) {
                const int li = (int)(s_list[ci] & kListPosMask);
                s_cnt[li] = 0u;   // (nothing else clears s_cnt)
                bool closed = ((s_refn[li >> 5] >> (li & 31)) & 1u) != 0u;                                       // SSVPL:298-305
                if (ignore_lower && ((s_refl[li >> 5] >> (li & 31)) & 1u)) closed = true;                        // SSVPL:198
                if (!closed) atomicOr(&s_bm[li >> 5], 1u << (li & 31));
            }
            __syncthreads();
            for (int wi = tid; wi <= nw; wi += nthr) {
                int o = 0;
                for (int k = 0; k < wi; k++) o += __popc(s_bm[k]);
                if (wi == nw) {   // the total: one reservation in the global list for the whole window
                    mset[M_NUNRES] = o;
                    mset[M_R1] = o ? (int)atomicAdd(&a.cq.ctr[2], (unsigned int)o) : 0;
                } else {
                    uint32_t bits = s_bm[wi];
                    while (bits) { const int b = __ffs(bits) - 1; bits &= bits - 1u; s_list[o++] = (uint32_t)((wi << 5) | b); }
                }
            }
            __syncthreads();
            const int n_o = mset[M_NUNRES];
            const unsigned int gbase = (unsigned int)mset[M_R1];
            for (int k = warp; k < n_o; k += nwarps) {
                const int e_pos = wbeg + (int)s_list[k] + 1;
                int i_lo, i_hi;
                if (single) {
                    Recs<false> R; R.hdr = s_hdr; R.pay = s_pay; R.pay_base = wr.off_lo & ~3u; R.n = n_rec;
                    i_lo = warp_lower_bound<false>(R, e_pos - max_span + 1, lane);
                    i_hi = warp_lower_bound<false>(R, e_pos + max_back + 1, lane);
                } else {
                    const TileDev& d = a.tiles[wr.tile];
                    Recs<true> R; R.hdr = d.hdr + wr.lo; R.pay = d.payload; R.pay_base = 0u; R.n = n_rec;
                    i_lo = warp_lower_bound<true>(R, e_pos - max_span + 1, lane);
                    i_hi = warp_lower_bound<true>(R, e_pos + max_back + 1, lane);
                }
                if (lane == 0 && gbase + (unsigned int)k < a.cq.cap_open)
                    a.cq.open[gbase + k] = make_uint4((unsigned int)wr.tile, (unsigned int)e_pos, (unsigned int)(wr.lo + i_lo),
                                                      span_capped ? kOpenSearch : (uint32_t)min(max(i_hi - i_lo, 0), (int)kOpenSearch));
            }
        } else if (n_cand > 0) {
            const uint32_t* s_refn = (const uint32_t*)(stg + lay.o_refn);
            const uint32_t* s_refl = (const uint32_t*)(stg + lay.o_refl);
            // C1, resolve pass (thread per candidate): everything the proofs need goes into the list entry
            for (int ci = tid; ci < n_cand; ci += nthr) {
                const uint32_t ent = s_list[ci];
                const int li = (int)(ent & kListPosMask);
                const uint32_t m_up = s_cnt[li];
                s_cnt[li] = 0u;   // the counter of a listed position goes back to zero here: nothing else clears s_cnt
                bool closed = ((s_refn[li >> 5] >> (li & 31)) & 1u) != 0u;                                       // VD:101-107
                if (ignore_lower && ((s_refl[li >> 5] >> (li & 31)) & 1u)) closed = true;                        // SSVPL:198
                const int rr = li + kRefLeft;
                const uint32_t ref = (wc.s_ref2[rr >> 4] >> ((rr & 15) * 2)) & 3u;
                int need = (m_up < (uint32_t)kSkipTable) ? min(s_kmin[m_up], kProofNoNeed) : kProofNoNeed;
                if (need <= 0) need = 1;
                if (multi || prm.pool_ploidy >= 3) need = kProofNoNeed;   // sample axis / pools: every candidate goes to the genotype stage (no hom-ref proof holds there)
                if (closed) need = 0;
                s_list[ci] = (uint32_t)li | (ref << kProofRefShift) | ((uint32_t)need << kProofNeedShift) |
                             ((uint32_t)min((int)(ent >> kListShift), kProofMaxRec) << kProofRecShift);
            }
            __syncthreads();   // ---- sync 3: entries resolved -----------------------------------------------------------------
            ExactCtx x;
            x.wbeg = wbeg; x.n_cand = n_cand; x.max_span = max_span; x.max_back = max_back; x.tile = wr.tile; x.rec0 = wr.lo;
            x.s_list = s_list; x.s_cnt = s_cnt; x.s_ref2 = wc.s_ref2; x.s_refn = s_refn; x.s_refl = s_refl;
            x.s_kmin = s_kmin; x.emit_all = emit_all; x.ignore_lower = ignore_lower;
            x.s_ulist = s_ulist; x.n_unres = 0;
            x.span_capped = span_capped || multi;   // the sample axis takes record ranges, not resolved entries
            OpenSink sink;
            sink.s_openq = (uint4*)(smem + lay.openq) + (it & 1) * kMaxOpen;
            sink.n_open = &mset[M_NOPEN];
            sink.s_chunk = &s_chunk[warp];
            const ngsep_read_hdr* ghdr = a.tiles[wr.tile].hdr + wr.lo;
            if (single) {
                Recs<false> R; R.hdr = s_hdr; R.pay = s_pay; R.pay_base = wr.off_lo & ~3u; R.n = n_rec;
                quarter_proof_phase<false>(R, x, s_ulist, &mset[M_NUNRES], warp, lane, nwarps);
                __syncthreads();   // ---- sync 4: candidates left to the half-warp pass listed --------------------------------------
                x.n_unres = mset[M_NUNRES];
                proof_phase<false>(R, x, a.cq, sink, ghdr, warp, lane, nwarps);
            } else {
                const TileDev& d = a.tiles[wr.tile];
                Recs<true> R; R.hdr = d.hdr + wr.lo; R.pay = d.payload; R.pay_base = 0u; R.n = n_rec;
                quarter_proof_phase<true>(R, x, s_ulist, &mset[M_NUNRES], warp, lane, nwarps);
                __syncthreads();
                x.n_unres = mset[M_NUNRES];
                proof_phase<true>(R, x, a.cq, sink, ghdr, warp, lane, nwarps);
            }
        }
        __syncthreads();   // ---- sync 2: the stage and the window arrays are free -------------------------------------------

        // one stage: the next window's copies start NOW (its descriptor has been in dq[0] since the top of this iteration), so
        // that their latency overlaps the tail of this window instead of opening the next iteration
        if (tid == 0 && n_stages == 1 && wn < a.n_windows) {
            issue_window(t0.dq[0], wn, 0);
            t0.prefetched = 1;
            // ... and the records of the window after that one are asked into L2 (dq[1] has held its descriptor since the top)
            if (wn + stride < a.n_windows && a.cfg.l2_prefetch) {
                const WinRange& q = t0.dq[1];
                if (q.tile == t0.tile && fits(q) && q.hi > q.lo) {
                    const uint32_t o0 = q.off_lo & ~3u;
                    bulk_prefetch_l2(t0.hdr + q.lo, (uint32_t)(q.hi - q.lo + 1) * 16u);
                    bulk_prefetch_l2(t0.payload + o0, ((q.off_hi - o0) * 4u + 15u) & ~15u);
                }
            }
        }
        // warp 1 hands the window's open columns to the global list: ONE global atomic per window, its round trip hidden behind
        // the wait for the next window (the list and the counters are per window parity)
        if (warp == 1 % nwarps) {
            const int n_open = min(mset[M_NOPEN], kMaxOpen);
            if (n_open > 0) {
                const uint4* s_openq = (const uint4*)(smem + lay.openq) + (it & 1) * kMaxOpen;
                unsigned int gbase = 0u;
                if (lane == 0) gbase = atomicAdd(&a.cq.ctr[2], (unsigned int)n_open);
                gbase = __shfl_sync(kFull, gbase, 0);
                for (int i = lane; i < n_open; i += 32)
                    if (gbase + i < a.cq.cap_open) a.cq.open[gbase + i] = s_openq[i];
            }
        }

        int uncovered = mset[M_UNC];
        const int win_len = max(0, B - A + 1);
        if (uncovered > 0 && mset[M_HASVER] && n_rec > 0) {
            // rare: a gap in the cover of the unversioned records while realigner versions are present — count the cover
            // of every record with a bitmap (s_cnt is all zero here — every listed position reset its counter — and is cleared again below)
            __syncthreads();
            const TileDev& d = a.tiles[wr.tile];
            uint32_t* s_bm = s_cnt;
            const int n_words = (g_hi - wbeg + 31) >> 5;
            for (int i = tid; i < n_rec; i += nthr) {
                const ngsep_read_hdr h = d.hdr[wr.lo + i];
                int s, e;
                record_interval(h, d.payload + h.off, s, e);
                s = max(s, A); e = min(e, B);
                for (int p = s; p <= e; p++) { const int li = p - 1 - wbeg; atomicOr(&s_bm[li >> 5], 1u << (li & 31)); }
            }
            __syncthreads();
            int cnt = 0;
            for (int i = tid; i < n_words; i += nthr) cnt += __popc(s_bm[i]);
            for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(kFull, cnt, o);
            if (lane == 0 && cnt) atomicAdd(&mset[M_FBCOV], cnt);
            __syncthreads();
            uncovered = win_len - mset[M_FBCOV];
            for (int i = tid; i < n_words; i += nthr) s_bm[i] = 0u;
            __syncthreads();
        }
        if (tid == issuer) {
            if (wr.tile != t0.acc_tile) { flush_stats(); t0.acc_tile = wr.tile; t0.acc_cov = 0ull; t0.acc_cand = 0ull; }
            t0.acc_cov += (unsigned long long)(win_len - uncovered);
            if (!emit_all) t0.acc_cand += (unsigned long long)n_cand;
        }
        w = wn;
    }
    if (tid == issuer) flush_stats();
}

// ------------------------------------------------------------------------------------------------------------
// SNVQ genotyping of the open columns: one warp per column, records read through L2
// ------------------------------------------------------------------------------------------------------------
template <int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks) k_genotype_columns(const TileDev* __restrict__ tiles, const ColumnQueue cq, const DevParams* __restrict__ gprm) {
    __shared__ double s_tab[96];
    __shared__ double s_tab2[kTab2Size];
    __shared__ int s_kmin[kSkipTable];
    const DevParams& prm = *gprm;
    for (int i = threadIdx.x; i < NGSEP_N_TABLES; i += blockDim.x) s_tab[i] = prm.tables[i];
    for (int i = threadIdx.x; i < kTab2Size; i += blockDim.x) s_tab2[i] = prm.tab2[i];   // (filled once per context: DevParams.tab2)
    if (threadIdx.x < kSkipTable) s_kmin[threadIdx.x] = prm.k_min[threadIdx.x];
    __syncthreads();
    const bool ignore_lower = (prm.flags & NGSEP_F_IGNORE_LOWER_REF) != 0;
    GenoCtx g;
    g.s_tab = s_tab; g.s_tab2 = s_tab2; g.s_kmin = s_kmin; g.max_qs = prm.max_base_qs; g.min_quality = prm.min_quality;
    g.emit_all = (prm.flags & NGSEP_F_EMIT_ALL) != 0;
    g.may_skip = prm.fast_homref != 0;
    const int lane = threadIdx.x & 31;
    const unsigned int n_open = min(cq.ctr[2], cq.cap_open);
    const unsigned int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (unsigned int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_open; t += n_warps) {
        const uint4 oc = cq.open[t];   // tile, position, then a record range or a resolved entry list (kOpenEntries)
        const TileDev& d = tiles[oc.x];
        const int pos = (int)oc.y;
        if (kWithEntries && (oc.w & kOpenEntries)) {
            // the window kernel resolved the column's records while they sat in shared memory (and found the reference base
            // callable): one coalesced entry load and one quality byte per record
            const int n = (int)(oc.w & kOpenCountMask);
            const uint2* ent = cq.ent + oc.z;
            const QualView qual(d);
            const int max_qs = g.max_qs;
            exact_column([&](int k) { return (k < n) ? entry_code(ent[k], qual, max_qs) : 0u; }, n, d, g, pos, (int)((oc.w >> kOpenRefShift) & 3u), false, lane);
            continue;
        }
        const int64_t gi = (int64_t)pos - 1;
        if (ignore_lower && ((d.ref_l[gi >> 5] >> (gi & 31)) & 1u)) continue;           // SSVPL:198: no discovery at all
        const bool ref_bad = (d.ref_n[gi >> 5] >> (gi & 31)) & 1u;
        const int ref = (int)((d.ref2[gi >> 4] >> ((gi & 15) * 2)) & 3u);
        Recs<true> R; R.hdr = d.hdr; R.pay = d.payload; R.pay_base = 0u; R.n = d.n_reads;
        int i_lo = (int)oc.z, i_hi = (int)oc.z + (int)oc.w;
        if (oc.w >= kOpenSearch) {   // no range from the window kernel (test hook, a very deep column, spans beyond 65535)
            i_lo = warp_lower_bound(R, pos - d.max_span + 1, lane);
            i_hi = warp_lower_bound(R, pos + d.max_back + 1, lane);   // first record that starts too far to the right
        }
        const QualView qual(d);
        const int max_qs = g.max_qs;
        exact_column([&](int k) {
            if (i_lo + k >= i_hi) return 0u;
            const ngsep_read_hdr h = R.header(i_lo + k);
            return lane_eval(h, R.payload(h), qual, pos, max_qs);
        }, i_hi - i_lo, d, g, pos, ref, ref_bad, lane);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Pools / polyploids (ploidy >= 3): the columns with a supported alternative allele, with their evidence in pileup order
// ------------------------------------------------------------------------------------------------------------
// SingleSampleVariantPileupListener.discoverSNV for ploidy >= DEF_MIN_PLOIDY_POOL_ALGORITHM (SSVPL:240-254).  No hom-ref proof holds
// for the pool genotyper, so every candidate of the screen comes here (as on the sample axis).  The warp counts the column exactly
// (PR.getAlleleCalls(1) + the quality rules of CH.updateCounts), applies createSNVVariantPool's support rule (SSVPL:306-319: an
// alternative base needs max(1, 0.5 / ploidy * counted bases) counted bases) — which closes almost every candidate, a lone sequencing
// error — and hands the surviving columns to the host WITH THEIR COUNTED BASES IN PILEUP ORDER (one byte each: quality << 2 | base):
// genotypeVariantPool (SSVPL:402-503) needs one CountsHelper per allele-frequency hypothesis and allele list, each a sum over the same
// reads in the same order, and its posteriors hang on libm; the host's table sums over these bytes are the reference's doubles.
// A column's record (kind NGSEP_CALL_POOL) holds counts, depth and the first 128 evidence bytes in its logcond field; a deeper column
// continues in the call slots behind it, 224 bytes each (`sample` = number of evidence bytes).
constexpr int kPoolInline = 128;
// site >= 0: a known site (-knownVariants with ploidy >= 3, SSVPL:366-370): the column is handed over whatever it holds, post_max = site
template <class CodeAt>
__device__ __forceinline__ void pool_column(CodeAt code_at, int n, const TileDev& d, int pos, int ref, bool ref_bad, int lane, double pool_frac, int site = -1) {
    ColCounts cc;
    for (int b = 0; b < 4; b++) { cc.c[b] = 0; cc.sn[b] = 0; }
    cc.depth = 0;
    const bool in_regs = n <= 32 * kRegRounds;
    uint32_t codes[kRegRounds];
    if (in_regs) {
#pragma unroll
        for (int r = 0; r < kRegRounds; r++) codes[r] = code_at(32 * r + lane);
#pragma unroll
        for (int r = 0; r < kRegRounds; r++) count_round(codes[r], cc);
    } else {
        for (int base = 0; base < n; base += 32) count_round(code_at(base + lane), cc);
    }
    const int sum = cc.c[0] + cc.c[1] + cc.c[2] + cc.c[3];
    if (site < 0) {
        if (ref_bad || cc.depth == 0) return;                                 // SSVPL:298-305
        const double min_count = fmax(1.0, pool_frac * (double)sum);          // SSVPL:311
        bool supported = false;
        for (int b = 0; b < 4; b++) if (b != ref && (double)cc.c[b] >= min_count) supported = true;
        if (!supported) return;                                               // only the reference allele: no variant (SSVPL:322-331)
    }
    const unsigned int slots = 1u + (unsigned int)((max(0, sum - kPoolInline) + (int)sizeof(ngsep_snv_call) - 1) / (int)sizeof(ngsep_snv_call));
    unsigned int slot = 0;
    if (lane == 0) slot = atomicAdd(d.n_calls, slots);
    slot = __shfl_sync(kFull, slot, 0);
    if (slot + slots > d.cap_calls) return;   // host sees n_calls > cap, grows the buffer and reruns
    ngsep_snv_call* o = &d.calls[slot];
    if (lane == 0) {
        o->seq_idx = d.seq_idx; o->pos = pos;
        o->count[0] = cc.c[0]; o->count[1] = cc.c[1]; o->count[2] = cc.c[2]; o->count[3] = cc.c[3];
        o->depth = cc.depth;
        for (int b = 0; b < 4; b++) { o->strand[2 * b] = cc.sn[b]; o->strand[2 * b + 1] = cc.c[b] - cc.sn[b]; }
        o->post_ref = 0.0; o->post_max = (double)site; o->gq = 0; o->qual = 0; o->a0 = -1; o->a1 = -1;
        o->ref = (int8_t)ref; o->kind = (int8_t)NGSEP_CALL_POOL; o->sample = sum;
    }
    uint8_t* ev0 = (uint8_t*)o->logcond;            // the first kPoolInline bytes
    uint8_t* ev1 = (uint8_t*)(o + 1);               // the rest, in the slots behind
    int written = 0;
    auto emit_round = [&](uint32_t code) {
        const bool counted = (code & kCodeCounted) != 0u;
        const unsigned int m = __ballot_sync(kFull, counted);
        if (counted) {
            const int idx = written + __popc(m & ((1u << lane) - 1u));
            const uint8_t e = (uint8_t)(code & 0x7fu);                        // quality << 2 | base
            if (idx < kPoolInline) ev0[idx] = e; else ev1[idx - kPoolInline] = e;
        }
        written += __popc(m);
    };
    if (in_regs) {
#pragma unroll
        for (int r = 0; r < kRegRounds; r++) emit_round(codes[r]);
    } else {
        for (int base = 0; base < n; base += 32) emit_round(code_at(base + lane));
    }
}

__global__ void __launch_bounds__(256, 4) k_pool_columns(const TileDev* __restrict__ tiles, const ColumnQueue cq, const DevParams* __restrict__ gprm) {
    const DevParams& prm = *gprm;
    const bool ignore_lower = (prm.flags & NGSEP_F_IGNORE_LOWER_REF) != 0;
    const int lane = threadIdx.x & 31, max_qs = prm.max_base_qs;
    const double pool_frac = 0.5 / (double)prm.pool_ploidy;
    const unsigned int n_open = min(cq.ctr[2], cq.cap_open);
    const unsigned int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (unsigned int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_open; t += n_warps) {
        const uint4 oc = cq.open[t];
        const TileDev& d = tiles[oc.x];
        const int pos = (int)oc.y;
        const int64_t gi = (int64_t)pos - 1;
        if (ignore_lower && ((d.ref_l[gi >> 5] >> (gi & 31)) & 1u)) continue;           // SSVPL:198
        const bool ref_bad = (d.ref_n[gi >> 5] >> (gi & 31)) & 1u;
        const int ref = (int)((d.ref2[gi >> 4] >> ((gi & 15) * 2)) & 3u);
        Recs<true> R; R.hdr = d.hdr; R.pay = d.payload; R.pay_base = 0u; R.n = d.n_reads;
        int i_lo = (int)oc.z, i_hi = (int)oc.z + (int)oc.w;
        if (oc.w >= kOpenSearch) {
            i_lo = warp_lower_bound(R, pos - d.max_span + 1, lane);
            i_hi = warp_lower_bound(R, pos + d.max_back + 1, lane);
        }
        const QualView qual(d);
        pool_column([&](int k) {
            if (i_lo + k >= i_hi) return 0u;
            const ngsep_read_hdr h = R.header(i_lo + k);
            return lane_eval(h, R.payload(h), qual, pos, max_qs);
        }, i_hi - i_lo, d, pos, ref, ref_bad, lane, pool_frac);
    }
}

// ------------------------------------------------------------------------------------------------------------
// MultisampleVariantsDetector: the open columns of a multi-sample context, one CTA per column
// ------------------------------------------------------------------------------------------------------------
// MSVD.onPileup for a span-1 position (MSVD:521-537).  At staging the records of a tile are indexed BY READ GROUP
// (gidx / goff: a stable sort of the record indexes by group id, so every group's list is in pileup order).  For a
// column covered by records [i_lo, i_hi) the records of one read group are then a contiguous piece of its list (two
// binary searches), and ONE THREAD PER SAMPLE walks the pieces of its read groups — in the order
// PR.getAlleleCalls(span, Set) visits the groups (PR:104-111) — evaluating each record once (lane_eval =
// PR.getAlleleCalls(1) + the quality rules of CH.updateCounts) and adding it to the sample's counts and its 10 distinct
// log-likelihoods in fp64 (CH:231-249): every sample's logConditionalProbs are bit-identical to the Java doubles.  The
// records whose read group belongs to no sample (list n_groups) only enter the POOLED counts (discoverPopulationSNV,
// MSVD:586-587), which are the sum over all lists.  Phase C: createSNVVariantPool (SSVPL:297-332), the allele
// refinement loop (MSVD:590-595, 642-656), per-sample VD.genotypeSNV (VD:21-97) with the sample-level minQuality of
// SSVPL:388, QUAL = best GQ of the decided non-hom-ref samples (MSVD:685-691), the filter of MSVD:533.
struct SamplePost { double hom[4], het[6]; };   // posteriors: [i][i], and ONE ordered event of (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)

__device__ __forceinline__ int pair_index(int lo, int hi) { return (lo == 0) ? hi - 1 : ((lo == 1) ? hi + 1 : 5); }   // into het[6]

// CH.getPosteriorProbabilities (CH:410-443) + calculatePosteriorProbabilities (CH:472-495): 16 ordered events, the sum in
// the reference's order.  L: [i][i] for i = 0..3, then the six pairs.
__device__ __forceinline__ void sample_posteriors(const double* L, double lp_hom, double lp_het, SamplePost& P) {
    double ev[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        ev[4 * i] = L[i] + lp_hom;
        int k = 4 * i + 1;
#pragma unroll
        for (int j = 0; j < 4; j++) if (j != i) { ev[k++] = L[4 + pair_index(min(i, j), max(i, j))] + lp_het; }
    }
    double mx = ev[0];
#pragma unroll
    for (int k = 1; k < 16; k++) mx = fmax(mx, ev[k]);
    double total = 0.0;
#pragma unroll
    for (int k = 0; k < 16; k++) { const double e = ev[k] - mx; ev[k] = (e < -20.0) ? 0.0 : exp10(e); total += ev[k]; }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        P.hom[i] = ev[4 * i] / total;
#pragma unroll
        for (int j = i + 1; j < 4; j++) P.het[pair_index(i, j)] = ev[4 * i + j] / total;   // event (i,j), j > i, sits at 4i + j
    }
}

// One sample's genotype for the allele list A (base indexes, reference first).  Returns g0 | g1 << 8 | gq << 16 with
// g = 0xff for "undecided".
__device__ __forceinline__ uint32_t sample_genotype(const SamplePost& P, int depth, const int* A, int nA, int sample_min_quality) {
    if (depth == 0) return 0xffffu;                                    // VD:23-27: undecided, no quality
    int g0, g1, gq;
    if (nA == 2) {                                                     // variant instanceof SNV, VD:33-60
        const int r = A[0], a = A[1];
        double pmax = P.hom[r];
        g0 = g1 = 0;
        if (P.hom[a] > pmax + 0.01) { pmax = P.hom[a]; g0 = g1 = 1; }
        const double ph = P.het[pair_index(min(r, a), max(r, a))];
        const double phet = ph + ph;                                   // postProbs[r][a] + postProbs[a][r]: equal doubles
        if (phet > pmax + 0.01) { pmax = phet; g0 = 0; g1 = 1; }
        gq = dev_phred(1.0 - pmax);
        if (gq == 0) return 0xffffu;                                   // VD:50
    } else {                                                           // VD:61-95: arg-max over the sub-matrix, from (0,0)
        g0 = g1 = 0;
        double pmax = P.hom[A[0]];
        for (int i = 0; i < nA; i++)
            for (int j = i; j < nA; j++) {
                double gp;
                if (i == j) gp = P.hom[A[i]];
                else { const double ph = P.het[pair_index(min(A[i], A[j]), max(A[i], A[j]))]; gp = ph + ph; }
                if (gp > pmax + 0.01) { pmax = gp; g0 = i; g1 = j; }
            }
        double max_p;
        if (g0 == g1) max_p = P.hom[A[g0]];
        else { const double ph = P.het[pair_index(min(A[g0], A[g1]), max(A[g0], A[g1]))]; max_p = ph + ph; }
        gq = dev_phred(1.0 - max_p);
    }
    if (sample_min_quality > gq) return 0xffffu;                       // SSVPL:388 makeUndecided
    return (uint32_t)g0 | ((uint32_t)g1 << 8) | ((uint32_t)gq << 16);
}


// ------------------------------------------------------------------------------------------------------------
// -knownVariants: the SNV sites of a tile (ngsep_tile_sites), one warp per site
// ------------------------------------------------------------------------------------------------------------
// SingleSampleVariantPileupListener.onPileup with input variants (SSVPL:162-176) for a variant that is an SNV:
// genotypeVariantSample (SSVPL:361-391) = PR.getAlleleCalls(1) + CH.calculateCountsSNV over EVERY covering record in
// pileup order (no screen, no hom-ref proof: a known site is reported whatever its genotype) + VD.genotypeSNV (VD:21-97)
// for the site's own allele list + the minQuality rule of SSVPL:388.  Record t of the tile's call array belongs to site t.
__global__ void __launch_bounds__(256) k_genotype_sites(const TileDev* __restrict__ tiles, const DevParams* __restrict__ gprm) {
    __shared__ double s_tab2[kTab2Size];
    const DevParams& prm = *gprm;
    for (int i = threadIdx.x; i < kTab2Size; i += blockDim.x) s_tab2[i] = prm.tab2[i];
    __syncthreads();
    const TileDev& d = tiles[0];
    const int lane = threadIdx.x & 31;
    const int n_warps = (int)((gridDim.x * blockDim.x) >> 5);
    const int max_qs = prm.max_base_qs;
    if (blockIdx.x == 0 && threadIdx.x == 0) *d.n_calls = (unsigned int)d.n_sites;
    for (int t = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); t < d.n_sites; t += n_warps) {
        const int pos = d.site_pos[t];
        const uint32_t al = d.site_alleles[t];   // bits 0-7: up to four base indexes, reference first; bits 8-10: how many
        Recs<true> R; R.hdr = d.hdr; R.pay = d.payload; R.pay_base = 0u; R.n = d.n_reads;
        const int i_lo = warp_lower_bound(R, pos - d.max_span + 1, lane);
        const int i_hi = warp_lower_bound(R, pos + d.max_back + 1, lane);
        const QualView qual(d);
        double acc = 0.0;
        ColCounts cc;
        for (int b = 0; b < 4; b++) { cc.c[b] = 0; cc.sn[b] = 0; }
        cc.depth = 0;
        for (int base = i_lo; base < i_hi; base += 32) {
            uint32_t code = 0u;
            if (base + lane < i_hi) {
                const ngsep_read_hdr h = R.header(base + lane);
                code = lane_eval(h, R.payload(h), qual, pos, max_qs);
            }
            count_round(code, cc);
            accumulate_round(code, min(32, i_hi - base), acc, lane, s_tab2);
        }
        // logConditionalProbs[i][j] on lane 4 i + j (the matrix is symmetric: CH:231-249 with f == g == 250)
        const int i = (lane & 15) >> 2, j = lane & 3;
        const int lo = min(i, j), hi = max(i, j);
        const int owner = (i == j) ? i : ((lo == 0) ? 3 + hi : ((lo == 1) ? 5 + hi : 9));
        const double Lij = __shfl_sync(kFull, acc, owner);
        double L[10];
#pragma unroll
        for (int k = 0; k < 10; k++) L[k] = __shfl_sync(kFull, acc, k);
        if ((unsigned int)t >= d.cap_calls) continue;
        ngsep_snv_call* o = &d.calls[t];
        if (lane < 16) o->logcond[lane] = Lij;
        if (lane == 16) {
            o->seq_idx = d.seq_idx; o->pos = pos;
            o->count[0] = cc.c[0]; o->count[1] = cc.c[1]; o->count[2] = cc.c[2]; o->count[3] = cc.c[3];
            o->depth = cc.depth;
            for (int b = 0; b < 4; b++) { o->strand[2 * b] = cc.sn[b]; o->strand[2 * b + 1] = cc.c[b] - cc.sn[b]; }
        }
        if (lane == 0) {
            const int nA = (int)((al >> 8) & 7u);
            int A[4];
            for (int k = 0; k < 4; k++) A[k] = (int)((al >> (2 * k)) & 3u);
            uint32_t gt = 0xffffu;
            SamplePost P;
            P.hom[A[0]] = 0.0;
            if (cc.depth > 0) {
                sample_posteriors(L, prm.tables[93], prm.tables[94], P);
                gt = sample_genotype(P, cc.depth, A, nA, prm.min_quality);
            }
            const bool undecided = (gt & 0xffu) == 0xffu;
            o->post_ref = (cc.depth > 0) ? P.hom[A[0]] : 0.0;
            o->post_max = 0.0;
            o->gq = (int16_t)(undecided ? 0 : (int)(gt >> 16));
            o->qual = 0;
            o->a0 = (int8_t)(undecided ? -1 : (int)(gt & 0xffu));
            o->a1 = (int8_t)(undecided ? -1 : (int)((gt >> 8) & 0xffu));
            o->ref = (int8_t)A[0];
            o->kind = (int8_t)NGSEP_CALL_SITE;
            o->sample = t;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// RelativeAlleleCountsCalculator: allele counts of EVERY position of a tile (a PileupListener reducer)
// ------------------------------------------------------------------------------------------------------------
// discovery/RelativeAlleleCountsCalculator.java:246-290 per position: calls = PR.getAlleleCalls(1); fewer than minRD calls: nothing;
// the calls whose base quality is at least minBaseQualityScore are counted per allele; out = the largest count, the second largest
// and the number of alleles seen.  Read-centric: one CTA owns a window of positions and its count tile in shared memory
// (position x {A, C, G, T, other, calls}); every record that can reach the window scatters its callable bases into the tile
// (shared-memory atomics), then one thread per position reduces its row.  The callable-base rules are the detector's own
// (resolve_offset / the simple-record test of lane_eval = RA.getAlleleCall, RA:747-871).
constexpr int kAlleleT = 1024;
__global__ void __launch_bounds__(256) k_allele_counts(const TileDev* __restrict__ tiles, int min_rd, int min_bq, uint2* __restrict__ out) {
    __shared__ unsigned int s_cnt[kAlleleT * 6];
    __shared__ int s_range[2];
    const TileDev& d = tiles[0];
    const QualView qual(d);
    const int n_pos = d.tile_last - d.tile_first + 1;
    const int n_win = (n_pos + kAlleleT - 1) / kAlleleT;
    const int tid = threadIdx.x, lane = tid & 31;
    Recs<true> R; R.hdr = d.hdr; R.pay = d.payload; R.pay_base = 0u; R.n = d.n_reads;
    for (int w = blockIdx.x; w < n_win; w += gridDim.x) {
        const int A = d.tile_first + w * kAlleleT, B = min(d.tile_last, A + kAlleleT - 1);
        for (int i = tid; i < kAlleleT * 6; i += blockDim.x) s_cnt[i] = 0u;
        if (tid < 32) {
            const int lo = warp_lower_bound(R, A - d.max_span + 1, lane);
            const int hi = warp_lower_bound(R, B + d.max_back + 1, lane);
            if (lane == 0) { s_range[0] = lo; s_range[1] = hi; }
        }
        __syncthreads();
        const int i_lo = s_range[0], i_hi = s_range[1];
        for (int i = i_lo + tid; i < i_hi; i += blockDim.x) {
            const ngsep_read_hdr h = R.header(i);
            const uint32_t* pay = R.payload(h);
            if (h.meta & NGSEP_M_SIMPLE) {
                const int L = (int)(h.meta & 0xffffu);
                const uint32_t* b2 = pay + 1;
                const uint32_t qoff = b2[((L + 15) >> 4) + ((L + 31) >> 5)];
                const int o_lo = max(max((int)h.ign_start, 0), A - h.first);                 // RA:782-784
                const int o_hi = min(L - max((int)h.ign_end, 0), B - h.first + 1);
                for (int off = o_lo; off < o_hi; off++) {
                    unsigned int* row = s_cnt + (h.first + off - A) * 6;
                    atomicAdd(row + 5, 1u);
                    if ((int)qual.at(qoff, (uint32_t)off) - 33 >= min_bq) atomicAdd(row + ((b2[off >> 4] >> ((off & 15) * 2)) & 3u), 1u);
                }
            } else {
                const RecView r = view_record(h, pay);
                int s, e;
                record_interval(h, pay, s, e);
                s = max(s, A); e = min(e, B);
                for (int pos = s; pos <= e; pos++) {
                    const int off = resolve_offset(r, h.ign_start, h.ign_end, pos);
                    if (off < 0) continue;
                    unsigned int* row = s_cnt + (pos - A) * 6;
                    atomicAdd(row + 5, 1u);
                    if ((int)qual.at(r.qoff, (uint32_t)off) - 33 < min_bq) continue;
                    const bool other = (r.meta & NGSEP_M_NMASK) && ((r.nm[off >> 5] >> (off & 31)) & 1u);   // a base outside ACGT is an allele of its own
                    atomicAdd(row + (other ? 4u : ((r.b2[off >> 4] >> ((off & 15) * 2)) & 3u)), 1u);
                }
            }
        }
        __syncthreads();
        for (int p = tid; p <= B - A; p += blockDim.x) {
            const unsigned int* row = s_cnt + p * 6;
            uint2 o = make_uint2(0u, 0u);
            if ((int)row[5] >= min_rd) {
                unsigned int mx = 0u, second = 0u, n_al = 0u;
                for (int k = 0; k < 5; k++) {
                    const unsigned int c = row[k];
                    if (c) n_al++;
                    if (c > mx) { second = mx; mx = c; } else if (c > second) second = c;
                }
                o = make_uint2(mx | (n_al << 28), second);
            }
            out[(size_t)(A - d.tile_first) + (size_t)p] = o;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------
// Indel / STR allele tallies of the span > 1 events: one event per thread
// ------------------------------------------------------------------------------------------------------------
// The per-call part of CountsHelper.updateCountsIndel (CH:253-304), call by call in pileup order: calculateLogCond (CH:384-396) of
// the call against every allele of its length, the allele the call counts for (an exact match, else the best-fitting one), the
// allele and strand counts, the diagonal of logConditionalProbs.  Every value is a sum of host-supplied table entries in the
// reference's order, so the outputs are bit-identical to the Java doubles; what depends on libm (the mixture terms of the
// off-diagonal entries, the posteriors) is finished by the host from call_index / call_fit.  Events are a few per ten thousand
// positions and a few dozen calls each: the point is that these tallies are made where the SNV tallies are made, not speed.
__global__ void __launch_bounds__(128) k_indel_tallies(const IndelEventsDev ev, const DevParams* __restrict__ gprm, int* __restrict__ total, int* __restrict__ counts,
                                                       int* __restrict__ strand, double* __restrict__ diag, int* __restrict__ call_index, double* __restrict__ call_fit,
                                                       double* __restrict__ lc_scratch) {
    const int k = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (k >= ev.n_events) return;
    const double* tab = gprm->tables;   // tables[3 q] = log10(1 - e_q) (a match), tables[3 q + 1] = log10(e_q / 3) (a mismatch)
    const int a0 = ev.allele_first[k], n = ev.allele_first[k + 1] - a0;
    const int c0 = ev.call_first[k], c1 = ev.call_first[k + 1];
    double* lc = lc_scratch + a0;
    double* dg = diag + a0;
    int* cnt = counts + a0;
    int* snd = strand + 2 * a0;
    const double lerr = ev.log_err;
    for (int i = 0; i < n; i++) { cnt[i] = 0; snd[2 * i] = 0; snd[2 * i + 1] = 0; dg[i] = 0.0; }
    for (int c = c0; c < c1; c++) {
        const int co = ev.call_off[c], clen = ev.call_off[c + 1] - co;
        const char* cs = ev.call_chars + co;
        const unsigned char* cq = ev.call_quals + co;
        int index = -1, best = -1;
        for (int i = 0; i < n; i++) {
            const int ao = ev.allele_off[a0 + i], alen = ev.allele_off[a0 + i + 1] - ao;
            const char* as = ev.allele_chars + ao;
            lc[i] = lerr;
            if (alen != clen) continue;
            bool same = true;
            double v = 0.0;
            for (int b = 0; b < alen; b++) {
                const bool eq = as[b] == cs[b];
                same = same && eq;
                const int q = (int)(signed char)min(ev.max_qs, (int)cq[b] - 33);
                if (q < 3) continue;
                v += eq ? tab[3 * q] : tab[3 * q + 1];
            }
            if (same && index < 0) index = i;
            lc[i] = fmax(lerr, v);
            if (lc[i] > lerr && (best == -1 || lc[best] < lc[i])) best = i;
        }
        if (index >= 0 && best >= 0 && best != index) index = min(index, best);
        else if (index < 0 && best >= 0) index = best;
        if (index >= 0) { cnt[index]++; snd[2 * index + (ev.call_neg[c] ? 0 : 1)]++; }
        for (int i = 0; i < n; i++) dg[i] += lc[i];
        call_index[c] = index;
        call_fit[c] = (index >= 0) ? lc[index] : lerr;
    }
    total[k] = c1 - c0;
}

// Multi-sample staging: the evidence bytes of every SIMPLE record and the headers with their quality offsets (TileDev.mhdr / ev).
// A column of the sample axis evaluates every covering record once per column — 960 records x 100 columns for a GBS tag —, so what
// one evaluation touches is what the kernel's L2 traffic is made of: header + evidence byte = two 32-byte sectors instead of header +
// base word + qoff word + quality byte.
__global__ void k_build_evidence(const TileDev d, ngsep_read_hdr* __restrict__ mhdr, uint8_t* __restrict__ ev, int max_qs) {
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= d.n_reads) return;
    if (d.report && d.report[0]) return;   // the tile failed its staging check: nothing of it is walked
    const ngsep_read_hdr h = d.hdr[i];
    const uint32_t* pay = d.payload + h.off;
    ngsep_read_hdr m = h;
    if (h.meta & NGSEP_M_SIMPLE) {
        const int L = (int)(h.meta & 0xffffu);
        const uint32_t* b2 = pay + 1;
        const uint32_t qoff = b2[((L + 15) >> 4) + ((L + 31) >> 5)];
        m.off = qoff;
        const QualView qual(d);
        uint32_t* out = (uint32_t*)(ev + (size_t)qoff * d.ev_mul);   // ev_mul is a multiple of 4: word stores
        for (int k0 = 0; k0 < L; k0 += 4) {
            uint32_t w = 0u;
            for (int k = k0; k < min(L, k0 + 4); k++) {
                int q = (int)qual.at(qoff, (uint32_t)k) - 33;
                q = min(30, q);                                     // CH:91
                uint32_t e = 0u;                                    // CH:214-216: a call that is not counted
                if (q > 3) { if (q > max_qs) q = max_qs; e = 0x80u | ((uint32_t)q << 2) | ((b2[k >> 4] >> ((k & 15) * 2)) & 3u); }   // CH:217-219
                w |= e << (8 * (k - k0));
            }
            out[k0 >> 2] = w;
        }
    } else {
        m.off = view_record(h, pay).qoff;
    }
    mhdr[i] = m;
}

void launch_build_evidence(const TileDev& d, ngsep_read_hdr* mhdr, uint8_t* ev, int max_qs, cudaStream_t s) {
    if (d.n_reads > 0) k_build_evidence<<<(d.n_reads + 127) / 128, 128, 0, s>>>(d, mhdr, ev, max_qs);
}

// First entry of read-group list g that is >= want.  gblk[g][b] = the answer for want = b * kGroupBlock (built at
// staging), so the search only looks at the few list entries that fall into one block of record indexes.
constexpr int kGroupBlockShift = 8, kGroupBlock = 1 << kGroupBlockShift;
__device__ __forceinline__ int list_lower_bound(const uint32_t* __restrict__ lst, const uint32_t* __restrict__ blk, int n_blk, uint32_t want) {
    const int b = min((int)(want >> kGroupBlockShift), n_blk);
    int lo = (int)__ldg(blk + b), hi = (int)__ldg(blk + min(b + 1, n_blk));
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(lst + mid) < want) lo = mid + 1; else hi = mid; }
    return lo;
}

struct SampleAcc { double L[10]; int c0, c1, c2, c3, depth; };
constexpr int kInFlight = 6;

__device__ __forceinline__ void sample_add(SampleAcc& a, uint32_t code, const double* s_tab) {
    if (code & kCodeCall) a.depth++;                                   // CH:210
    if (!(code & kCodeCounted)) return;
    const int b = (int)(code & 3u), q = (int)((code >> 2) & 31u);
    a.c0 += (b == 0); a.c1 += (b == 1); a.c2 += (b == 2); a.c3 += (b == 3);   // CH:223
    const double ta = s_tab[3 * q], tb = s_tab[3 * q + 1], tc = s_tab[3 * q + 2];
#pragma unroll
    for (int i = 0; i < 4; i++) a.L[i] += (b == i) ? ta : tb;          // CH:231-235
    a.L[4] += (b == 0 || b == 1) ? tc : tb;                            // CH:236-249 with f == g == 250
    a.L[5] += (b == 0 || b == 2) ? tc : tb;
    a.L[6] += (b == 0 || b == 3) ? tc : tb;
    a.L[7] += (b == 1 || b == 2) ? tc : tb;
    a.L[8] += (b == 1 || b == 3) ? tc : tb;
    a.L[9] += (b == 2 || b == 3) ? tc : tb;
}

template <int kThr>
__global__ void __launch_bounds__(kThr, 1024 / kThr) k_genotype_columns_multi(const TileDev* __restrict__ tiles, const ColumnQueue cq, const DevParams* __restrict__ gprm) {
    __shared__ double s_tab[96];
    __shared__ int s_pool[5];            // pooled A, C, G, T counts and totalCount
    __shared__ int s_red[4];             // [0] called-allele mask, [1] best quality, [2] output slot
    const DevParams& prm = *gprm;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < NGSEP_N_TABLES; i += kThr) s_tab[i] = prm.tables[i];
    const bool ignore_lower = (prm.flags & NGSEP_F_IGNORE_LOWER_REF) != 0;
    const bool emit_all = (prm.flags & NGSEP_F_EMIT_ALL) != 0;
    const int n_samples = prm.n_samples, n_groups = prm.n_groups, max_qs = prm.max_base_qs;
    const int n_batches = (n_samples + 1 + kThr - 1) / kThr;   // + the list of the records that belong to no sample
    ngsep_sample_call* scratch = prm.scratch + (size_t)blockIdx.x * (size_t)n_samples;
    const unsigned int n_open = min(cq.ctr[2], cq.cap_open);
    // a CTA takes RUNS of kColumnRun neighbouring list entries: the window kernel lists a window's columns in position order,
    // so the columns of a run share most of their records and a sample's thread finds them in L1 again
    const unsigned int run = (unsigned int)max(1, prm.column_run);
    for (unsigned int t = blockIdx.x * run; t < n_open; t = ((t + 1) % run) ? t + 1 : t + 1 + (gridDim.x - 1) * run) {
        __syncthreads();   // every `continue` below is uniform over the CTA; the previous column's shared state is free from here on
        const uint4 oc = cq.open[t];   // tile, position, first record, number of records
        const TileDev& d = tiles[oc.x];
        const int pos = (int)oc.y;
        const int64_t gi = (int64_t)pos - 1;
        if (ignore_lower && ((d.ref_l[gi >> 5] >> (gi & 31)) & 1u)) continue;           // SSVPL:198
        const bool ref_bad = (d.ref_n[gi >> 5] >> (gi & 31)) & 1u;
        const int ref = (int)((d.ref2[gi >> 4] >> ((gi & 15) * 2)) & 3u);
        Recs<true> R; R.hdr = d.hdr; R.pay = d.payload; R.pay_base = 0u; R.n = d.n_reads;
        const QualView qv(d);
        int i_lo = (int)oc.z, i_hi = (int)oc.z + (int)oc.w;
        if (oc.w >= kOpenSearch) {
            i_lo = warp_lower_bound(R, pos - d.max_span + 1, lane);
            i_hi = warp_lower_bound(R, pos + d.max_back + 1, lane);
        }
        if (tid < 5) s_pool[tid] = 0;
        if (tid < 4) s_red[tid] = 0;
        __syncthreads();
        // ---- phases A + B: one thread per sample ------------------------------------------------------------------------
        for (int batch = 0; batch < n_batches; batch++) {
            const int smp = batch * kThr + tid;
            SampleAcc a;
#pragma unroll
            for (int k = 0; k < 10; k++) a.L[k] = 0.0;
            a.c0 = a.c1 = a.c2 = a.c3 = a.depth = 0;
            int g_lo = 0, g_n = 0;
            if (smp < n_samples) { g_lo = prm.sample_group_off[smp]; g_n = prm.sample_group_off[smp + 1] - g_lo; }
            else if (smp == n_samples) g_n = 1;   // the records outside every sample
            for (int gk = 0; gk < g_n; gk++) {
                const int g = (smp < n_samples) ? (int)prm.sample_groups[g_lo + gk] : n_groups;
                const uint32_t* lst = d.gidx + d.goff[g];
                const uint32_t* blk = d.gblk + (size_t)g * (size_t)(d.n_gblk + 1);
                int ka = list_lower_bound(lst, blk, d.n_gblk, (uint32_t)i_lo);
                const int kb = list_lower_bound(lst, blk, d.n_gblk, (uint32_t)i_hi);
                for (; ka < kb; ka += kInFlight) {   // several records in flight: their loads are independent, the sums stay in list order
                    uint32_t code[kInFlight];
#pragma unroll
                    for (int u = 0; u < kInFlight; u++) {
                        code[u] = 0u;
                        if (ka + u < kb) {
                            const int idx = (int)__ldg(lst + ka + u);
                            const int4 mh = __ldg((const int4*)(d.mhdr + idx));   // first | meta | ign_start, ign_end | qoff
                            if ((uint32_t)mh.y & NGSEP_M_SIMPLE) {
                                // one match block as long as the read (RA:842-871 needs no payload): the header and ONE evidence byte
                                const int L = (int)((uint32_t)mh.y & 0xffffu), off = pos - mh.x;
                                const int ign_s = (int)(int16_t)(mh.z & 0xffff), ign_e = (int)(int16_t)((uint32_t)mh.z >> 16);
                                if (off >= 0 && off < L && off >= ign_s && (L - off) > ign_e) {                  // RA:782-784
                                    const uint32_t e = __ldg(d.ev + (size_t)(uint32_t)mh.w * d.ev_mul + (size_t)off);
                                    code[u] = kCodeCall | ((e & 0x80u) ? (kCodeCounted | (e & 0x7fu)) : 0u) | (((uint32_t)mh.y & NGSEP_M_NEG) ? kCodeNeg : 0u);
                                }
                            } else {
                                const ngsep_read_hdr h = R.header(idx);
                                code[u] = lane_eval(h, R.payload(h), qv, pos, max_qs);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kInFlight; u++) sample_add(a, code[u], s_tab);
                }
            }
            if (smp <= n_samples) {   // pooled counts: every record of the column is in exactly one list
                if (a.c0) atomicAdd(&s_pool[0], a.c0);
                if (a.c1) atomicAdd(&s_pool[1], a.c1);
                if (a.c2) atomicAdd(&s_pool[2], a.c2);
                if (a.c3) atomicAdd(&s_pool[3], a.c3);
                if (a.depth) atomicAdd(&s_pool[4], a.depth);
            }
            if (smp < n_samples) {
                ngsep_sample_call& o = scratch[smp];
                o.count[0] = a.c0; o.count[1] = a.c1; o.count[2] = a.c2; o.count[3] = a.c3; o.depth = a.depth;
                o.gq = 0; o.g0 = -1; o.g1 = -1;
#pragma unroll
                for (int k = 0; k < 10; k++) o.logcond[k] = a.L[k];
            }
        }
        __syncthreads();
        // ---- phase C ---------------------------------------------------------------------------------------------------
        const int pc[4] = {s_pool[0], s_pool[1], s_pool[2], s_pool[3]};
        const int pool_depth = s_pool[4];
        if (emit_all && d.counts_dbg && tid == 0) {
            int32_t* o = d.counts_dbg + (int64_t)(pos - d.tile_first) * 5;
            o[0] = pc[0]; o[1] = pc[1]; o[2] = pc[2]; o[3] = pc[3]; o[4] = pool_depth;
        }
        if (ref_bad || pool_depth == 0) continue;                                  // SSVPL:298-305 (uniform over the CTA)
        // createSNVVariantPool (SSVPL:306-330): the reference allele, then every other base with count >= minCount
        int A[4], nA = 1;
        {
            const int sum = pc[0] + pc[1] + pc[2] + pc[3];
            const double min_count = fmax(1.0, prm.min_allele_depth_freq * (double)sum);
            A[0] = ref; A[1] = A[2] = A[3] = 0;
            for (int b = 0; b < 4; b++) if (b != ref && (double)pc[b] >= min_count) A[nA++] = b;
        }
        if (nA < 2) continue;
        const int nA0 = nA;
        const double lp_hom = s_tab[93], lp_het = s_tab[94];
        int qs = 0;
        bool final_pass = nA <= 2;   // the refinement loop runs only for more than two alleles (MSVD:590)
        for (;;) {
            // genotypeVariant (MSVD:674-693): every sample, with the current allele list
            unsigned int mask = 0u;
            int best = 0;
            for (int smp = tid; smp < n_samples; smp += kThr) {
                ngsep_sample_call& sc = scratch[smp];
                // A sample without one counted base of any alternative allele of the list is called homozygous reference
                // (or left undecided): every other genotype has at most the likelihood of (ref,ref) and a smaller or
                // equal prior.  It cannot add an allele or raise the variant quality, so its posterior is
                // only computed if the site turns out to be reported (g0 = -2: pending).
                if (prm.fast_homref) {
                    int alt = 0;
                    for (int k = 1; k < nA; k++) alt += sc.count[A[k]];
                    if (alt == 0) { if (final_pass) { sc.gq = 0; sc.g0 = -2; sc.g1 = -2; } continue; }
                }
                double L[10];
#pragma unroll
                for (int k = 0; k < 10; k++) L[k] = sc.logcond[k];
                uint32_t g = 0xffffu;
                if (sc.depth > 0) {
                    SamplePost P;
                    sample_posteriors(L, lp_hom, lp_het, P);
                    g = sample_genotype(P, sc.depth, A, nA, prm.sample_min_quality);
                }
                const int g0 = (int)(g & 0xffu), g1 = (int)((g >> 8) & 0xffu), gq = (int)(g >> 16);
                if (g0 != 0xff) {
                    mask |= (1u << A[g0]) | (1u << A[g1]);                          // getCalledAlleles
                    if (!(g0 == 0 && g1 == 0) && gq > best) best = gq;               // MSVD:685-687
                }
                if (final_pass) { sc.gq = (int16_t)gq; sc.g0 = (g0 == 0xff) ? (int8_t)-1 : (int8_t)g0; sc.g1 = (g1 == 0xff) ? (int8_t)-1 : (int8_t)g1; }
            }
            for (int o = 16; o; o >>= 1) { mask |= __shfl_xor_sync(kFull, mask, o); best = max(best, __shfl_xor_sync(kFull, best, o)); }
            if (lane == 0) { atomicOr((unsigned int*)&s_red[0], mask); atomicMax(&s_red[1], best); }
            __syncthreads();
            const unsigned int called = (unsigned int)s_red[0] | (1u << ref);
            qs = s_red[1];
            __syncthreads();
            if (tid == 0) { s_red[0] = 0; s_red[1] = 0; }
            __syncthreads();
            if (final_pass) break;
            // makeNewVariant (MSVD:642-656): reference + called alleles (TreeSet order = base order); unchanged size ends the loop
            if (__popc(called & 0xfu) != nA) {
                int B[4] = {ref, 0, 0, 0}, nB = 1;
                for (int b = 0; b < 4; b++) if (b != ref && ((called >> b) & 1u)) B[nB++] = b;
                nA = nB;
                for (int k = 0; k < 4; k++) A[k] = B[k];
                if (nA <= 2) final_pass = true;
            } else final_pass = true;
            if (nA < 2) break;
        }
        // a variant reduced to its reference allele is not an SNV any more; every sample is then called homozygous
        // reference or left undecided, so variantQS == 0 and MSVD:533 drops it
        if (nA < 2) continue;
        if (qs == 0 || qs < prm.min_quality) continue;                              // MSVD:533
        for (int smp = tid; smp < n_samples; smp += kThr) {   // the site is reported: the pending samples get their genotype now
            ngsep_sample_call& sc = scratch[smp];
            if (sc.g0 != -2) continue;
            double L[10];
#pragma unroll
            for (int k = 0; k < 10; k++) L[k] = sc.logcond[k];
            uint32_t g = 0xffffu;
            if (sc.depth > 0) {
                SamplePost P;
                sample_posteriors(L, lp_hom, lp_het, P);
                g = sample_genotype(P, sc.depth, A, nA, prm.sample_min_quality);
            }
            const int g0 = (int)(g & 0xffu), g1 = (int)((g >> 8) & 0xffu);
            sc.gq = (int16_t)(g >> 16); sc.g0 = (g0 == 0xff) ? (int8_t)-1 : (int8_t)g0; sc.g1 = (g1 == 0xff) ? (int8_t)-1 : (int8_t)g1;
        }
        if (tid == 0) s_red[2] = (int)atomicAdd(d.n_calls, 1u);
        __syncthreads();
        const unsigned int slot = (unsigned int)s_red[2];
        if (slot < d.cap_calls) {   // else: the host sees n_calls > cap, grows the buffers and reruns the tile
            if (tid == 0) {
                ngsep_pop_site& o = d.pop_sites[slot];
                o.seq_idx = d.seq_idx; o.pos = pos;
                o.count[0] = pc[0]; o.count[1] = pc[1]; o.count[2] = pc[2]; o.count[3] = pc[3]; o.depth = pool_depth;
                o.qs = (int16_t)qs; o.ref = (int8_t)ref; o.n_alleles = (int8_t)nA; o.n_alleles0 = (int8_t)nA0;
                for (int k = 0; k < 4; k++) o.allele[k] = (int8_t)((k < nA) ? A[k] : -1);
            }
            const uint2* src = (const uint2*)scratch;
            uint2* dst = (uint2*)(d.pop_calls + (size_t)slot * (size_t)n_samples);
            const int n_words = n_samples * (int)(sizeof(ngsep_sample_call) / 8);
            for (int k = tid; k < n_words; k += kThr) dst[k] = src[k];
        }
    }
}

// CTAs per SM that are resident at once: the column loop is grid-strided, so the grid must not exceed one wave
int genotype_multi_ctas_per_sm(int n_samples) {
    int n = 0;
    if (const char* e = getenv("NGSEP_MULTI_MAX_L1")) if (atoi(e)) {   // experiment: the unified cache goes to L1
        cudaFuncSetAttribute(k_genotype_columns_multi<32>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
        cudaFuncSetAttribute(k_genotype_columns_multi<64>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
        cudaFuncSetAttribute(k_genotype_columns_multi<128>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    }
    if (n_samples + 1 <= 32) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_genotype_columns_multi<32>, 32, 0);
    else if (n_samples + 1 <= 64) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_genotype_columns_multi<64>, 64, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_genotype_columns_multi<128>, 128, 0);
    return n;
}

void launch_genotype_multi(const TileDev* d_tiles, const ColumnQueue& cq, const DevParams* dprm, int n_samples, int grid, cudaStream_t s) {
    // the CTA is as wide as the sample axis (+ the list of the records outside every sample), up to 128 threads
    if (n_samples + 1 <= 32) k_genotype_columns_multi<32><<<grid, 32, 0, s>>>(d_tiles, cq, dprm);
    else if (n_samples + 1 <= 64) k_genotype_columns_multi<64><<<grid, 64, 0, s>>>(d_tiles, cq, dprm);
    else k_genotype_columns_multi<128><<<grid, 128, 0, s>>>(d_tiles, cq, dprm);
}

// ---- staging: records indexed by read group (stable) ---------------------------------------------------------------
__global__ void k_group_keys(const uint16_t* __restrict__ group, const uint8_t* __restrict__ in_sample, int n, int n_groups, uint16_t* keys, uint32_t* vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int g = (int)group[i];
    // ids outside the table (0xFFFF) and read groups that belong to no sample share the last list: pooled discovery only
    keys[i] = (uint16_t)((g < n_groups && in_sample[g]) ? g : n_groups);
    vals[i] = (uint32_t)i;
}
__global__ void k_group_offsets(const uint16_t* __restrict__ keys_sorted, int n, int n_groups, uint32_t* goff) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n_groups + 1) return;
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)keys_sorted[mid] < g) lo = mid + 1; else hi = mid; }
    goff[g] = (uint32_t)lo;
}

// gblk[g][b], b = 0..n_blk: position in list g of its first record index >= b * kGroupBlock
__global__ void k_group_blocks(const uint32_t* __restrict__ gidx, const uint32_t* __restrict__ goff, int n_groups, int n_blk, uint32_t* gblk) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)(n_groups + 1) * (n_blk + 1)) return;
    const int g = (int)(t / (n_blk + 1)), b = (int)(t % (n_blk + 1));
    const uint32_t* lst = gidx + goff[g];
    int lo = 0, hi = (int)(goff[g + 1] - goff[g]);
    const uint32_t want = (uint32_t)b << kGroupBlockShift;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (lst[mid] < want) lo = mid + 1; else hi = mid; }
    gblk[t] = (uint32_t)lo;
}

int group_blocks(int n) { return (n + kGroupBlock - 1) >> kGroupBlockShift; }

size_t group_index_temp_bytes(int n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint16_t*)nullptr, (uint16_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, n, 0, 16);
    return bytes;
}

// gidx[n]: record indexes sorted by group (stable: each group's list keeps the pileup order); goff[n_groups + 2]: list g is
// gidx[goff[g] .. goff[g+1]), list n_groups = the records outside every sample.  keys_tmp: 2 x n uint16, vals_tmp: n uint32.
cudaError_t launch_group_index(const uint16_t* group, const uint8_t* in_sample, int n, int n_groups, uint16_t* keys_tmp, uint32_t* vals_tmp, void* cub_tmp, size_t cub_bytes,
                               uint32_t* gidx, uint32_t* goff, uint32_t* gblk, cudaStream_t s) {
    if (n > 0) {
        k_group_keys<<<(n + 255) / 256, 256, 0, s>>>(group, in_sample, n, n_groups, keys_tmp, vals_tmp);
        int bits = 1;
        while ((1 << bits) <= n_groups) bits++;
        cudaError_t e = cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, (const uint16_t*)keys_tmp, keys_tmp + n, (const uint32_t*)vals_tmp, gidx, n, 0, bits, s);
        if (e != cudaSuccess) return e;
    }
    k_group_offsets<<<(n_groups + 2 + 127) / 128, 128, 0, s>>>(keys_tmp + n, n, n_groups, goff);
    const long long cells = (long long)(n_groups + 1) * (group_blocks(n) + 1);
    k_group_blocks<<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(gidx, goff, n_groups, group_blocks(n), gblk);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// window index: record range + payload span of every window (two binary searches over the sorted starts)
// ------------------------------------------------------------------------------------------------------------
__global__ void k_window_index(const TileDev d, WinRange* out, int tile_idx) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= d.n_windows) return;
    int64_t wbeg = (int64_t)(d.win0 + w) * d.T;
    if (d.report && d.report[0]) {   // the tile failed its staging check: an empty window, nothing is walked
        WinRange r; r.lo = r.hi = 0; r.off_lo = r.off_hi = 0u; r.tile = tile_idx; r.wbeg = (int)wbeg; r.g_pack = 0u; r.span_pack = 0u;
        out[w] = r;
        return;
    }
    int64_t p_lo = wbeg + 1 - d.max_span + 1;     // smallest ordering start that can still reach the window
    int64_t p_hi = wbeg + d.T + d.max_back;       // largest ordering start that can begin inside it
    int lo = 0, hi = d.n_reads;
    while (lo < hi) { int mid = (lo + hi) >> 1; if ((int64_t)d.hdr[mid].first < p_lo) lo = mid + 1; else hi = mid; }
    int a = lo;
    hi = d.n_reads;
    while (lo < hi) { int mid = (lo + hi) >> 1; if ((int64_t)d.hdr[mid].first <= p_hi) lo = mid + 1; else hi = mid; }
    WinRange r;
    r.lo = a; r.hi = lo;
    r.off_lo = d.hdr[a].off; r.off_hi = d.hdr[lo].off;
    r.tile = tile_idx; r.wbeg = (int)wbeg;
    // [g_lo, g_hi): 0-based indexes of the window inside the tile and the sequence, relative to wbeg
    const int64_t g_lo = max(wbeg, (int64_t)d.tile_first - 1), g_hi = max(g_lo, min(min(wbeg + d.T, (int64_t)d.tile_last), d.ref_len));
    r.g_pack = (uint32_t)(g_lo - wbeg) | ((uint32_t)(g_hi - wbeg) << 16);
    r.span_pack = (uint32_t)min(d.max_span, 0xffff) | ((uint32_t)min(d.max_back, 0xffff) << 16);
    out[w] = r;
}

// ------------------------------------------------------------------------------------------------------------
// reference packing: ASCII -> 2 bits/base + "not ACGT" + "lower case" bitmaps (32 bases per thread)
// ------------------------------------------------------------------------------------------------------------
__global__ void k_ref_pack(const char* __restrict__ ascii, int64_t len, uint32_t* ref2, uint32_t* ref_n, uint32_t* ref_l) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t base = t * 32;
    if (base >= len) return;
    uint32_t w0 = 0, w1 = 0, nb = 0, lb = 0;
    for (int i = 0; i < 32; i++) {
        int64_t k = base + i;
        uint32_t code = 0, isn = 1, isl = 0;
        if (k < len) {
            unsigned char c = (unsigned char)ascii[k];
            isl = (c >= 'a' && c <= 'z');
            unsigned char u = isl ? (unsigned char)(c - 32) : c;
            switch (u) { case 'A': code = 0; isn = 0; break; case 'C': code = 1; isn = 0; break;
                         case 'G': code = 2; isn = 0; break; case 'T': code = 3; isn = 0; break; default: break; }
        }
        if (i < 16) w0 |= code << (2 * i); else w1 |= code << (2 * (i - 16));
        nb |= isn << i;
        lb |= isl << i;
    }
    ref2[t * 2] = w0; ref2[t * 2 + 1] = w1; ref_n[t] = nb; ref_l[t] = lb;
}

void launch_window_index(const TileDev& d, WinRange* out, int tile_idx, cudaStream_t s) {
    int n = d.n_windows;
    if (n > 0) k_window_index<<<(n + 127) / 128, 128, 0, s>>>(d, out, tile_idx);
}

// Staging-time check of a tile, on the device copy of its headers (the host thread only enqueues copies):
//   report[0] |= 1  records not sorted by start      |= 2  payload offsets not monotone / out of range
//             |= 4  a record larger than one stage    |= 8  a record's body (CIGAR, bases, masks, qoff) does not fit its payload
//                                                          span, or its qualities reach beyond the quality array
//   report[2..3] += read lengths (aligned bases, 64 bit)
// A tile with report[0] != 0 gets an EMPTY window index (k_window_index), so the kernels never walk it; the host reports
// the error when it fetches the tile.  The packer's NGSEP_M_SIMPLE promise is verified here too: a record whose flag is
// not backed by its payload loses the flag (and goes through the general path, which is always correct).
__global__ void k_check_headers(ngsep_read_hdr* hdr, const uint32_t* __restrict__ payload, int n, unsigned long long n_words, unsigned long long n_qual,
                                unsigned int qual_bits, unsigned int pay_cap_words, unsigned int* report) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int err = 0u, L = 0u;
    if (i < n) {
        const ngsep_read_hdr h = hdr[i];
        const unsigned long long next = hdr[i + 1].off;   // the sentinel carries n_words
        if (i > 0 && h.first < hdr[i - 1].first) err |= 1u;
        if (next < h.off || next > n_words) err |= 2u;
        else if (next - h.off + 4ull > pay_cap_words) err |= 4u;
        L = h.meta & 0xffffu;
        if (!err) {
            const unsigned int nc = (h.meta >> 16) & 0xffu, lw = (L + 31u) >> 5;
            const unsigned long long body = ((h.meta & NGSEP_M_WINDOW) ? 3ull : 0ull) + nc + ((L + 15u) >> 4) + ((h.meta & NGSEP_M_NMASK) ? lw : 0u) + lw + 1ull;
            if (body > next - h.off) err |= 8u;
            else {
                const unsigned long long qoff = payload[h.off + body - 1ull];
                if (qoff * 4ull + ((unsigned long long)L * qual_bits + 7ull) / 8ull > n_qual) err |= 8u;
            }
        }
        if (!err && (h.meta & NGSEP_M_SIMPLE)) {
            const bool ok = (h.meta & (0x00ff0000u | NGSEP_M_NMASK | NGSEP_M_WINDOW)) == 0x00010000u && payload[h.off] == ((L << 3) | 3u);
            if (!ok) hdr[i].meta = h.meta & ~NGSEP_M_SIMPLE;
        }
    }
    for (int o = 16; o; o >>= 1) { L += __shfl_xor_sync(0xffffffffu, L, o); err |= __shfl_xor_sync(0xffffffffu, err, o); }
    if ((threadIdx.x & 31) == 0) {
        if (L) atomicAdd((unsigned long long*)(report + 2), (unsigned long long)L);
        if (err) atomicOr(report, err);
    }
}

// ------------------------------------------------------------------------------------------------------------
// NGSEP_LAYOUT_COMPACT: the wire stream -> the full payload layout (one scan between the two kernels)
// ------------------------------------------------------------------------------------------------------------
// What need not cross the bus does not: the CIGAR word of a single-block record is len*8+3 and its low-quality mask is a function
// of its qualities, so a record flagged NGSEP_M_SIMPLE travels as [bases2, qoff].  k_expand_sizes gives every record its length in
// the full layout (and checks the wire offsets it will follow), an exclusive scan turns the lengths into offsets, k_expand_records
// writes the full payload and the new hdr.off.  The staging checks (k_check_headers) then run on the result as for any other tile.
__global__ void k_expand_sizes(const ngsep_read_hdr* __restrict__ hdr, int n, unsigned long long n_wire, uint32_t* __restrict__ sizes, unsigned int* report) {
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i > n) return;
    if (i == n) { sizes[n] = 0u; return; }
    const ngsep_read_hdr h = hdr[i];
    const unsigned long long next = hdr[i + 1].off;   // the sentinel carries n_wire
    uint32_t sz = 0u;
    if (next < h.off || next > n_wire || next - h.off > 0x00ffffffull) atomicOr(report, 2u);
    else if (h.meta & NGSEP_M_SIMPLE) {
        const uint32_t L = h.meta & 0xffffu, nb = (L + 15u) >> 4, nw = (L + 31u) >> 5;
        const bool ok = (h.meta & (0x00ff0000u | NGSEP_M_NMASK | NGSEP_M_WINDOW)) == 0x00010000u && next - h.off == nb + 1ull;
        if (!ok) atomicOr(report, 8u); else sz = 1u + nb + nw + 1u;
    } else sz = (uint32_t)(next - h.off);
    sizes[i] = sz;
}

__global__ void k_expand_records(ngsep_read_hdr* __restrict__ hdr, int n, const uint32_t* __restrict__ wire, const uint32_t* __restrict__ sizes,
                                 const uint32_t* __restrict__ offs, uint32_t* __restrict__ payload, unsigned long long n_words_full, const uint8_t* __restrict__ qual,
                                 unsigned long long n_qual, uint32_t qual_bits, uint4 qual_dict, unsigned int* report) {
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i > n) return;
    if (report[0]) return;   // the wire offsets cannot be trusted (k_expand_sizes): nothing is expanded, the tile gets an empty window index
    if (i == n) {
        if ((unsigned long long)offs[n] != n_words_full) atomicOr(report, 2u);
        hdr[n].off = (uint32_t)n_words_full;
        return;
    }
    const ngsep_read_hdr h = hdr[i];
    const uint32_t sz = sizes[i], o = offs[i];
    if ((unsigned long long)o + sz > n_words_full) { atomicOr(report, 2u); return; }
    const uint32_t* src = wire + h.off;
    uint32_t* dst = payload + o;
    if (h.meta & NGSEP_M_SIMPLE) {
        const uint32_t L = h.meta & 0xffffu, nb = (L + 15u) >> 4, nw = (L + 31u) >> 5;
        dst[0] = (L << 3) | 3u;
        for (uint32_t k = 0; k < nb; k++) dst[1u + k] = src[k];
        const uint32_t qoff = src[nb];
        const bool q_ok = (unsigned long long)qoff * 4ull + ((unsigned long long)L * qual_bits + 7ull) / 8ull <= n_qual;
        if (!q_ok) atomicOr(report, 8u);
        const QualView qv(qual, qual_bits, qual_dict);
        for (uint32_t w = 0; w < nw; w++) {
            uint32_t m = 0u;
            if (q_ok) for (uint32_t k = 32u * w; k < min(L, 32u * w + 32u); k++) if ((int)qv.at(qoff, k) - 33 <= 3) m |= 1u << (k & 31u);   // CH:214
            dst[1u + nb + w] = m;
        }
        dst[1u + nb + nw] = qoff;
    } else {
        for (uint32_t k = 0; k < sz; k++) dst[k] = src[k];
    }
    hdr[i].off = o;
}

size_t expand_scan_temp_bytes(int n) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, n + 1);
    return bytes;
}

cudaError_t launch_expand_compact(ngsep_read_hdr* hdr, int n, const uint32_t* wire, unsigned long long n_wire, uint32_t* payload, unsigned long long n_words_full,
                                  const uint8_t* qual, unsigned long long n_qual, uint32_t qual_bits, const uint32_t* qual_dict4, uint32_t* sizes, uint32_t* offs,
                                  void* tmp, size_t tmp_bytes, unsigned int* report, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    k_expand_sizes<<<(n + 1 + 255) / 256, 256, 0, s>>>(hdr, n, n_wire, sizes, report);
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, (const uint32_t*)sizes, offs, n + 1, s);
    if (e != cudaSuccess) return e;
    const uint4 dict = make_uint4(qual_dict4[0], qual_dict4[1], qual_dict4[2], qual_dict4[3]);
    k_expand_records<<<(n + 1 + 127) / 128, 128, 0, s>>>(hdr, n, wire, sizes, offs, payload, n_words_full, qual, n_qual, qual_bits, dict, report);
    return cudaGetLastError();
}

void launch_check_headers(ngsep_read_hdr* hdr, const uint32_t* payload, int n, unsigned long long n_words, unsigned long long n_qual, unsigned int qual_bits,
                          unsigned int pay_cap_words, unsigned int* report, cudaStream_t s) {
    if (n > 0) k_check_headers<<<(n + 255) / 256, 256, 0, s>>>(hdr, payload, n, n_words, n_qual, qual_bits, pay_cap_words, report);
}

// The kernel is compiled for two CTA sizes and three register budgets each; the context picks one.
typedef void (*PileupKernel)(const KArgs);
static PileupKernel kernel_for(int threads, int min_blocks) {
    if (threads == 128) {
        if (min_blocks <= 8) return k_pileup_windows<128, 8>;
        if (min_blocks <= 10) return k_pileup_windows<128, 10>;
        return k_pileup_windows<128, 12>;
    }
    if (min_blocks <= 3) return k_pileup_windows<256, 3>;
    if (min_blocks == 4) return k_pileup_windows<256, 4>;
    if (min_blocks == 5) return k_pileup_windows<256, 5>;
    return k_pileup_windows<256, 6>;
}

int pileup_max_ctas_per_sm(size_t smem, int threads, int min_blocks) {
    PileupKernel k = kernel_for(threads, min_blocks);
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, threads, smem) != cudaSuccess) return 0;
    return n;
}

void launch_pileup(const TileDev* d_tiles, const WinRange* d_win, int n_windows, const DevParams* dprm, const ColumnQueue& cq, const LaunchCfg& cfg,
                   int grid, size_t smem, cudaStream_t s) {
    KArgs a;
    a.tiles = d_tiles; a.win = d_win; a.n_windows = n_windows; a.prm = dprm; a.cq = cq; a.cfg = cfg;
    a.lay = make_layout(cfg.T_lay, cfg.pay_cap_words, cfg.n_stages);
    kernel_for(cfg.threads, cfg.min_blocks)<<<grid, cfg.threads, smem, s>>>(a);   // grid = min(windows, resident CTAs): persistent CTAs
}

void launch_genotype(const TileDev* d_tiles, const ColumnQueue& cq, const DevParams* dprm, int grid, cudaStream_t s) {
    // one warp per open column, grid-stride; two register budgets (NGSEP_GENO_MIN_BLOCKS picks the other one for experiments)
    static const int min_blocks = [] { const char* e = getenv("NGSEP_GENO_MIN_BLOCKS"); return e ? atoi(e) : 4; }();
    if (min_blocks >= 6) k_genotype_columns<6><<<grid, 256, 0, s>>>(d_tiles, cq, dprm);
    else k_genotype_columns<4><<<grid, 256, 0, s>>>(d_tiles, cq, dprm);
}

// -knownVariants with ploidy >= 3: every known SNV site of the tile as an NGSEP_CALL_POOL column (one warp per site)
__global__ void __launch_bounds__(256, 4) k_pool_sites(const TileDev* __restrict__ tiles, const DevParams* __restrict__ gprm) {
    const DevParams& prm = *gprm;
    const TileDev& d = tiles[0];
    const int lane = threadIdx.x & 31, max_qs = prm.max_base_qs;
    const int n_warps = (int)((gridDim.x * blockDim.x) >> 5);
    for (int t = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); t < d.n_sites; t += n_warps) {
        const int pos = d.site_pos[t];
        Recs<true> R; R.hdr = d.hdr; R.pay = d.payload; R.pay_base = 0u; R.n = d.n_reads;
        const int i_lo = warp_lower_bound(R, pos - d.max_span + 1, lane);
        const int i_hi = warp_lower_bound(R, pos + d.max_back + 1, lane);
        const QualView qual(d);
        pool_column([&](int k) {
            if (i_lo + k >= i_hi) return 0u;
            const ngsep_read_hdr h = R.header(i_lo + k);
            return lane_eval(h, R.payload(h), qual, pos, max_qs);
        }, i_hi - i_lo, d, pos, (int)(d.site_alleles[t] & 3u), false, lane, 0.0, t);
    }
}

void launch_pool_sites(const TileDev* d_tile, const DevParams* dprm, int n_sites, int max_grid, cudaStream_t s) {
    k_pool_sites<<<max(1, min(max_grid, (n_sites + 7) / 8)), 256, 0, s>>>(d_tile, dprm);
}

void launch_pool_columns(const TileDev* d_tiles, const ColumnQueue& cq, const DevParams* dprm, int grid, cudaStream_t s) {
    k_pool_columns<<<grid, 256, 0, s>>>(d_tiles, cq, dprm);
}

void launch_genotype_sites(const TileDev* d_tile, const DevParams* dprm, int n_sites, int max_grid, cudaStream_t s) {
    const int grid = max(1, min(max_grid, (n_sites + 7) / 8));   // one warp per site
    k_genotype_sites<<<grid, 256, 0, s>>>(d_tile, dprm);
}

void launch_indel_tallies(const IndelEventsDev& ev, const DevParams* dprm, int* total, int* counts, int* strand, double* diag, int* call_index, double* call_fit,
                          double* lc_scratch, cudaStream_t s) {
    if (ev.n_events <= 0) return;
    k_indel_tallies<<<(ev.n_events + 127) / 128, 128, 0, s>>>(ev, dprm, total, counts, strand, diag, call_index, call_fit, lc_scratch);
}

void launch_allele_counts(const TileDev* d_tile, int n_pos, int min_rd, int min_bq, uint2* out, int max_grid, cudaStream_t s) {
    const int n_win = (n_pos + kAlleleT - 1) / kAlleleT;
    k_allele_counts<<<max(1, min(max_grid, n_win)), 256, 0, s>>>(d_tile, min_rd, min_bq, out);
}

void launch_ref_pack(const char* ascii, int64_t len, uint32_t* ref2, uint32_t* ref_n, uint32_t* ref_l, cudaStream_t s) {
    int64_t nt = (len + 31) / 32;
    k_ref_pack<<<(unsigned)((nt + 255) / 256), 256, 0, s>>>(ascii, len, ref2, ref_n, ref_l);
}

}  // namespace ngsep_b200
