// sm_100a kernels of libngsep_pileup.so — see pileup_kernels.cuh for the design summary.
// Reference citations are relative to /root/reference/src/ngsep/ (RA = alignments/ReadAlignment.java,
// PR = discovery/PileupRecord.java, CH = discovery/CountsHelper.java, VD = discovery/VariantDiscoverySNVQAlgorithm.java,
// SSVPL = discovery/SingleSampleVariantPileupListener.java).
#include "pileup_kernels.cuh"

#include <climits>

namespace ngsep_b200 {

// ------------------------------------------------------------------------------------------------------------
// shared-memory layout
// ------------------------------------------------------------------------------------------------------------
struct SmemLayout { unsigned bar, tab, ref2, refn, refl, cand, cov, list, misc, hdr, pay, total, max_reads; };

__host__ __device__ inline SmemLayout make_layout(int T, unsigned pay_cap_words) {
    SmemLayout L;
    unsigned o = 0;
    L.bar = o; o += 16;
    L.tab = o; o += 96 * 8;
    L.ref2 = o; o += (T / 16 + 2) * 4;
    L.refn = o; o += (T / 32 + 1) * 4;
    L.refl = o; o += (T / 32 + 1) * 4;
    L.cand = o; o += (T / 32 + 1) * 4;
    L.cov = o; o += (T / 32 + 1) * 4;
    L.list = o; o += T * 2;
    L.misc = o; o += 64;
    o = (o + 15u) & ~15u;
    L.max_reads = pay_cap_words / 16;
    L.hdr = o; o += (L.max_reads + 1) * 16;
    L.pay = o; o += (pay_cap_words + 8) * 4;
    L.total = o;
    return L;
}

size_t pileup_smem_bytes(int T, unsigned int pay_cap_words) { return make_layout(T, pay_cap_words).total; }

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
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// small bit helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t range_mask32(int lo, int hi) {  // bits [lo,hi), 0 <= lo < hi <= 32
    uint32_t m = (hi >= 32) ? 0xffffffffu : ((1u << hi) - 1u);
    return m & ~((1u << lo) - 1u);
}
__device__ __forceinline__ void mark_bits(uint32_t* bm, int a, int b) {  // set bits [a,b) of a shared bitmap
    if (a >= b) return;
    int w0 = a >> 5, w1 = (b - 1) >> 5;
    for (int w = w0; w <= w1; w++) {
        int lo = (w == w0) ? (a & 31) : 0;
        int hi = (w == w1) ? ((b - 1) & 31) + 1 : 32;
        atomicOr(&bm[w], range_mask32(lo, hi));
    }
}
__device__ __forceinline__ uint32_t extract16(const uint32_t* b2, int roff) {  // 16 packed bases starting at read offset roff
    if (roff >= 0) {
        int w = roff >> 4;
        return __funnelshift_r(b2[w], b2[w + 1], (roff & 15) * 2);
    }
    return b2[0] << ((-roff) * 2);
}

struct RecView {
    int first, vfrom, vto, L, nc;
    uint32_t meta;
    const uint32_t* cig;
    const uint32_t* b2;
    const uint32_t* nm;
    const uint8_t* q;
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
    r.q = (const uint8_t*)t;
    return r;
}

// ------------------------------------------------------------------------------------------------------------
// screen: candidate + coverage bits of one record inside the window
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void screen_record(const ngsep_read_hdr& h, const uint32_t* pay, int g_lo, int g_hi, int wbeg,
                                              const uint32_t* s_ref2, uint32_t* s_cand, uint32_t* s_cov) {
    RecView r = view_record(h, pay);
    // 0-based index range of this window/tile/version intersection
    int lo = max(g_lo, (r.vfrom == INT_MIN) ? INT_MIN : r.vfrom - 1);
    int hi = min(g_hi, (r.vto == INT_MAX) ? INT_MAX : r.vto - 1);
    if (lo >= hi) return;
    int ref_idx = r.first - 1, rpos = 0;
    const int cov_lo = ref_idx;
    for (int c = 0; c < r.nc; c++) {
        uint32_t v = r.cig[c];
        int len = (int)(v >> 3);
        if ((v & 3u) == 3u) {  // consumes reference and read: M(3) or X(7)
            int a = max(ref_idx, lo), b = min(ref_idx + len, hi);
            if (a < b) {
                const int delta = rpos - ref_idx;  // read offset = index + delta
                for (int k = a >> 4; k <= (b - 1) >> 4; k++) {
                    int gbase = k << 4;
                    uint32_t refw = s_ref2[k - (wbeg >> 4)];
                    uint32_t rd = extract16(r.b2, gbase + delta);
                    uint32_t x = rd ^ refw;
                    uint32_t m = (x | (x >> 1)) & 0x55555555u;
                    int vlo = max(a, gbase) - gbase, vhi = min(b, gbase + 16) - gbase;
                    m &= range_mask32(2 * vlo, 2 * vhi);
                    if (m) {
                        m = (m | (m >> 1)) & 0x33333333u;
                        m = (m | (m >> 2)) & 0x0f0f0f0fu;
                        m = (m | (m >> 4)) & 0x00ff00ffu;
                        m = (m | (m >> 8)) & 0x0000ffffu;
                        int bit = gbase - wbeg;
                        atomicOr(&s_cand[bit >> 5], m << (bit & 31));
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
    mark_bits(s_cov, max(cov_lo, lo) - wbeg, min(ref_idx, hi) - wbeg);  // PR.addAlignment: first <= pos <= last (PR:154-156)
}

// ------------------------------------------------------------------------------------------------------------
// exact: per-candidate state and the restatement of getAlleleCalls(1) + updateCounts
// ------------------------------------------------------------------------------------------------------------
struct ColState {
    int c0, c1, c2, c3, depth;
    int sn0, sn1, sn2, sn3, sp0, sp1, sp2, sp3;  // negative / positive strand per allele
    double d0, d1, d2, d3, o01, o02, o03, o12, o13, o23;
};
__device__ __forceinline__ void col_init(ColState& s) {
    s.c0 = s.c1 = s.c2 = s.c3 = s.depth = 0;
    s.sn0 = s.sn1 = s.sn2 = s.sn3 = s.sp0 = s.sp1 = s.sp2 = s.sp3 = 0;
    s.d0 = s.d1 = s.d2 = s.d3 = s.o01 = s.o02 = s.o03 = s.o12 = s.o13 = s.o23 = 0.0;
}

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
                int next_cons = 0;
                if (c + 1 < r.nc) {
                    uint32_t nv = r.cig[c + 1];
                    uint32_t nop = nv & 7u;
                    next_indel = (nop == 1u || nop == 2u);
                    next_cons = (nv & 2u) ? (int)(nv >> 3) : 0;
                }
                bool skip = rp < ign_s;
                skip = skip || (r.L - rp) <= ign_e;
                skip = skip || (prev_indel && j < 2);
                skip = skip || (next_indel && j < len - 1 && j >= len - 2);
                if (next_indel && j == len - 1) return -1;  // anchor: either skipped or an indel call (length > 1)
                (void)next_cons;
                return skip ? -1 : rp;
            }
        }
        if (v & 1u) ref_pos += len;
        if (v & 2u) read_pos += len;
        prev_indel = (op == 1u || op == 2u);
    }
    return -1;
}

__device__ __forceinline__ void col_add(ColState& s, const RecView& r, int off, const double* s_tab, int max_qs) {
    s.depth++;  // CH:210 totalCount counts every call returned by getAlleleCalls(1), whatever its quality or base
    int q = (int)r.q[off] - 33;
    q = min(30, q);                                   // CH:91
    if (q <= 3) return;                               // CH:214-216
    if (q > max_qs) q = max_qs;                       // CH:217-219
    if ((r.meta & NGSEP_M_NMASK) && ((r.nm[off >> 5] >> (off & 31)) & 1u)) return;  // allele not in ACGT: CH:220-221
    int b = (int)((r.b2[off >> 4] >> ((off & 15) * 2)) & 3u);
    bool neg = (r.meta & NGSEP_M_NEG) != 0;
    s.c0 += (b == 0); s.c1 += (b == 1); s.c2 += (b == 2); s.c3 += (b == 3);
    s.sn0 += (neg && b == 0); s.sn1 += (neg && b == 1); s.sn2 += (neg && b == 2); s.sn3 += (neg && b == 3);
    s.sp0 += (!neg && b == 0); s.sp1 += (!neg && b == 1); s.sp2 += (!neg && b == 2); s.sp3 += (!neg && b == 3);
    const double ta = s_tab[3 * q], tb = s_tab[3 * q + 1], tc = s_tab[3 * q + 2];
    // CH:231-249 with f == g == 250: one add per genotype per call, in pileup order
    s.d0 += (b == 0) ? ta : tb;
    s.d1 += (b == 1) ? ta : tb;
    s.d2 += (b == 2) ? ta : tb;
    s.d3 += (b == 3) ? ta : tb;
    s.o01 += (b == 0 || b == 1) ? tc : tb;
    s.o02 += (b == 0 || b == 2) ? tc : tb;
    s.o03 += (b == 0 || b == 3) ? tc : tb;
    s.o12 += (b == 1 || b == 2) ? tc : tb;
    s.o13 += (b == 1 || b == 3) ? tc : tb;
    s.o23 += (b == 2 || b == 3) ? tc : tb;
}

__device__ __forceinline__ int dev_phred(double p) {  // math/PhredScoreHelper.java:31-40
    if (p == 0.0) return 255;
    double score = -10.0 * log10(p);
    if (score > 255.0) return 255;
    double f = floor(score);
    return (int)((score - f >= 0.5) ? f + 1.0 : f);  // Math.round
}

// CH.getPosteriorProbabilities (CH:410-495) + VD.discoverSNV (VD:100-222) + the filter of SSVPL:223
__device__ void col_finalize(const ColState& s, int pos, int ref, const TileDev& d, const DevParams& prm, const double* s_tab) {
    double Lm[16];
    Lm[0] = s.d0; Lm[5] = s.d1; Lm[10] = s.d2; Lm[15] = s.d3;
    Lm[1] = Lm[4] = s.o01; Lm[2] = Lm[8] = s.o02; Lm[3] = Lm[12] = s.o03;
    Lm[6] = Lm[9] = s.o12; Lm[7] = Lm[13] = s.o13; Lm[11] = Lm[14] = s.o23;
    const double lp_hom = s_tab[93], lp_het = s_tab[94];
    double ev[16];
    {
        int k = 0;
        for (int i = 0; i < 4; i++) {
            ev[k++] = Lm[i * 5] + lp_hom;
            for (int j = 0; j < 4; j++) if (i != j) ev[k++] = Lm[i * 4 + j] + lp_het;
        }
    }
    double log_max = 1.0;  // CH:474-479, quirk kept: starts at 1 and is replaced by the first event
    for (int k = 0; k < 16; k++) if (log_max > 0.0 || log_max < ev[k]) log_max = ev[k];
    double total = 0.0;
    for (int k = 0; k < 16; k++) {
        double e = ev[k] - log_max;
        e = (e < -20.0) ? 0.0 : exp10(e);
        ev[k] = e;
        total += e;
    }
    double post[16];
    {
        int k = 0;
        for (int i = 0; i < 4; i++) {
            post[i * 5] = ev[k++] / total;
            for (int j = 0; j < 4; j++) if (i != j) post[i * 4 + j] = ev[k++] / total;
        }
    }
    // VD.getIndexesMaxGenotype (VD:223-243)
    int ii = ref, jj = ref;
    double pmax = post[ref * 5];
    for (int i = 0; i < 4; i++) for (int j = i; j < 4; j++) {
        double gp = post[i * 4 + j];
        if (i != j) gp += post[j * 4 + i];
        if (gp > pmax + 0.01) { pmax = gp; ii = i; jj = j; }
    }
    double max_p = post[ii * 4 + jj];
    if (ii != jj) max_p += post[jj * 4 + ii];
    int gq = dev_phred(1.0 - max_p);
    int kind, a0, a1;
    if (ii != jj && ii != ref && jj != ref) {  // VD:128-177
        kind = NGSEP_CALL_TRIALLELIC;
        if (post[ii * 5] > post[jj * 5] + 0.01) { a0 = ii; a1 = jj; } else { a0 = jj; a1 = ii; }
    } else if (ii != jj) {
        kind = NGSEP_CALL_HET; a0 = ref; a1 = (ref != ii) ? ii : jj;
    } else if (ii != ref) {
        kind = NGSEP_CALL_HOMALT; a0 = ii; a1 = ii;
    } else {
        return;  // homozygous reference: dropped by SSVPL:223
    }
    if (prm.min_quality > gq) return;  // SSVPL:223
    unsigned int slot = atomicAdd(d.n_calls, 1u);
    if (slot >= d.cap_calls) return;   // host sees n_calls > cap and reports NGSEP_EOVERFLOW
    ngsep_snv_call* o = &d.calls[slot];
    o->seq_idx = d.seq_idx; o->pos = pos;
    o->count[0] = s.c0; o->count[1] = s.c1; o->count[2] = s.c2; o->count[3] = s.c3;
    o->depth = s.depth;
    o->strand[0] = s.sn0; o->strand[1] = s.sp0; o->strand[2] = s.sn1; o->strand[3] = s.sp1;
    o->strand[4] = s.sn2; o->strand[5] = s.sp2; o->strand[6] = s.sn3; o->strand[7] = s.sp3;
    for (int k = 0; k < 16; k++) o->logcond[k] = Lm[k];
    o->post_ref = post[ref * 5];
    o->post_max = max_p;
    o->gq = (int16_t)gq;
    o->qual = (int16_t)dev_phred(post[ref * 5]);
    o->a0 = (int8_t)a0; o->a1 = (int8_t)a1; o->ref = (int8_t)ref; o->kind = (int8_t)kind;
    o->sample = 0;
}

// ------------------------------------------------------------------------------------------------------------
// the window kernel
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_pileup_windows(const TileDev d, const DevParams* __restrict__ gprm) {
    extern __shared__ __align__(128) unsigned char smem[];
    const SmemLayout lay = make_layout(d.T, d.pay_cap_words);
    uint64_t* bar = (uint64_t*)(smem + lay.bar);
    double* s_tab = (double*)(smem + lay.tab);
    uint32_t* s_ref2 = (uint32_t*)(smem + lay.ref2);
    uint32_t* s_refn = (uint32_t*)(smem + lay.refn);
    uint32_t* s_refl = (uint32_t*)(smem + lay.refl);
    uint32_t* s_cand = (uint32_t*)(smem + lay.cand);
    uint32_t* s_cov = (uint32_t*)(smem + lay.cov);
    uint16_t* s_list = (uint16_t*)(smem + lay.list);
    int* s_misc = (int*)(smem + lay.misc);
    ngsep_read_hdr* s_hdr = (ngsep_read_hdr*)(smem + lay.hdr);
    uint32_t* s_pay = (uint32_t*)(smem + lay.pay);

    const int tid = threadIdx.x;
    const int T = d.T;
    const DevParams& prm = *gprm;
    for (int i = tid; i < NGSEP_N_TABLES; i += kThreads) s_tab[i] = prm.tables[i];
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    uint32_t parity = 0;
    const bool emit_all = (prm.flags & NGSEP_F_EMIT_ALL) != 0;
    const bool ignore_lower = (prm.flags & NGSEP_F_IGNORE_LOWER_REF) != 0;

    for (int w = blockIdx.x; w < d.n_windows; w += gridDim.x) {
        const int wbeg = (d.win0 + w) * T;                       // 0-based index of the window's first position
        const int g_lo = max(wbeg, d.tile_first - 1);            // [g_lo, g_hi): indexes of this window inside the tile
        const int g_hi = (int)min((int64_t)min(wbeg + T, d.tile_last), d.ref_len);
        const WinRange wr = d.win[w];
        // stage the reference window + clear the bitmaps
        for (int i = tid; i < T / 16 + 1; i += kThreads) {
            int64_t gw = (int64_t)(wbeg >> 4) + i;
            s_ref2[i] = (gw * 16 < d.ref_len) ? d.ref2[gw] : 0u;
        }
        for (int i = tid; i < T / 32; i += kThreads) {
            int64_t gw = (int64_t)(wbeg >> 5) + i;
            bool in = gw * 32 < d.ref_len;
            s_refn[i] = in ? d.ref_n[gw] : 0xffffffffu;
            s_refl[i] = in ? d.ref_l[gw] : 0u;
            s_cand[i] = 0u;
            s_cov[i] = 0u;
        }
        __syncthreads();

        const bool single_chunk = (wr.hi - wr.lo) <= (int)lay.max_reads && (wr.off_hi - (wr.off_lo & ~3u)) <= d.pay_cap_words;
        int n_chunk = 0;         // records in the currently staged chunk
        uint32_t pay_base = 0;   // payload word offset staged at s_pay[0]
        int chunk_lo = 0;        // first record of the staged chunk

        // ---- chunk loader (all threads call; thread 0 issues the bulk copies) ------------------------------------
        auto load_chunk = [&](int r0) {
            if (tid == 0) {
                int r1;
                if (single_chunk) r1 = wr.hi;
                else {
                    uint32_t base = d.hdr[r0].off & ~3u;
                    int lo_r = r0 + 1, hi_r = min(wr.hi, r0 + (int)lay.max_reads);
                    r1 = r0;  // largest r1 in (r0, hi_r] with off[r1] - base <= cap
                    while (lo_r <= hi_r) {
                        int mid = (lo_r + hi_r) >> 1;
                        if (d.hdr[mid].off - base <= d.pay_cap_words) { r1 = mid; lo_r = mid + 1; } else hi_r = mid - 1;
                    }
                    if (r1 == r0) r1 = r0 + 1;  // cannot happen: tile_submit rejects records larger than a chunk
                }
                uint32_t o0 = (single_chunk ? wr.off_lo : d.hdr[r0].off) & ~3u;
                uint32_t o1 = single_chunk ? wr.off_hi : d.hdr[r1].off;
                uint32_t hb = (uint32_t)(r1 - r0 + 1) * 16u;                 // headers incl. the next record's (sentinel) header
                uint32_t pb = ((o1 - o0) * 4u + 15u) & ~15u;
                s_misc[0] = r1;
                s_misc[1] = (int)o0;
                if (r1 > r0) {
                    mbar_expect_tx(bar, hb + pb);
                    bulk_g2s(s_hdr, d.hdr + r0, hb, bar);
                    if (pb) bulk_g2s(s_pay, d.payload + o0, pb, bar);
                }
            }
            __syncthreads();
            int r1 = s_misc[0];
            if (r1 > r0) { mbar_wait(bar, parity); parity ^= 1u; }
            n_chunk = r1 - r0;
            pay_base = (uint32_t)s_misc[1];
            chunk_lo = r0;
            return r1;
        };

        // ---- phase A: screen ---------------------------------------------------------------------------------
        {
            int r0 = wr.lo;
            while (r0 < wr.hi) {
                int r1 = load_chunk(r0);
                for (int i = tid; i < n_chunk; i += kThreads)
                    screen_record(s_hdr[i], s_pay + (s_hdr[i].off - pay_base), g_lo, g_hi, wbeg, s_ref2, s_cand, s_cov);
                __syncthreads();
                r0 = r1;
            }
        }
        // ---- candidate list (one warp) -----------------------------------------------------------------------
        if (tid < 32) {
            int total = 0, covered = 0;
            for (int base_w = 0; base_w < T / 32; base_w += 32) {
                int wi = base_w + tid;
                uint32_t cw = 0, vw = 0;
                if (wi < T / 32) {
                    vw = s_cov[wi];
                    cw = emit_all ? vw : (s_cand[wi] & vw & ~s_refn[wi]);
                    if (ignore_lower) cw &= ~s_refl[wi];
                }
                int c = __popc(cw);
                int incl = c;
                for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += t; }
                int start = total + incl - c;
                while (cw) { int b = __ffs(cw) - 1; cw &= cw - 1; s_list[start++] = (uint16_t)(wi * 32 + b); }
                total += __shfl_sync(0xffffffffu, incl, 31);
                int cv = __popc(vw);
                for (int o = 16; o > 0; o >>= 1) cv += __shfl_xor_sync(0xffffffffu, cv, o);
                covered += cv;
            }
            if (tid == 0) {
                s_misc[2] = total;
                if (covered) atomicAdd(&d.counters[0], (unsigned long long)covered);
                if (total) atomicAdd(&d.counters[1], (unsigned long long)total);
            }
        }
        __syncthreads();
        const int n_cand = s_misc[2];

        // ---- phase B: exact evaluation, kThreads candidates per round ----------------------------------------------
        for (int base = 0; base < n_cand; base += kThreads) {
            const bool active = base + tid < n_cand;
            const int pos = active ? wbeg + (int)s_list[base + tid] + 1 : 0;  // 1-based
            ColState st;
            col_init(st);
            int r0 = wr.lo;
            while (r0 < wr.hi) {
                int r1;
                if (single_chunk) r1 = wr.hi;   // still resident from phase A
                else r1 = load_chunk(r0);
                if (active) {
                    // records whose ordering start lies in [pos - max_span + 1, pos + max_back]
                    int want = pos - d.max_span + 1;
                    int lo_i = 0, hi_i = n_chunk;
                    while (lo_i < hi_i) { int mid = (lo_i + hi_i) >> 1; if (s_hdr[mid].first < want) lo_i = mid + 1; else hi_i = mid; }
                    const int last_first = pos + d.max_back;
                    for (int i = lo_i; i < n_chunk; i++) {
                        const ngsep_read_hdr h = s_hdr[i];
                        if (h.first > last_first) break;
                        RecView r = view_record(h, s_pay + (h.off - pay_base));
                        int off = resolve_offset(r, h.ign_start, h.ign_end, pos);
                        if (off >= 0) col_add(st, r, off, s_tab, prm.max_base_qs);
                    }
                }
                if (!single_chunk) __syncthreads();
                r0 = r1;
            }
            if (active) {
                const int li = pos - 1 - wbeg;
                const bool ref_bad = (s_refn[li >> 5] >> (li & 31)) & 1u;
                const int ref = (int)((s_ref2[li >> 4] >> ((li & 15) * 2)) & 3u);
                if (emit_all && d.counts_dbg) {
                    int32_t* o = d.counts_dbg + (int64_t)(pos - d.tile_first) * 5;
                    o[0] = st.c0; o[1] = st.c1; o[2] = st.c2; o[3] = st.c3; o[4] = st.depth;
                }
                bool lower = ignore_lower && ((s_refl[li >> 5] >> (li & 31)) & 1u);
                if (!ref_bad && !lower && st.depth > 0) {  // VD:101-107, SSVPL:198
                    int nonref = st.c0 + st.c1 + st.c2 + st.c3 - ((ref == 0) ? st.c0 : (ref == 1) ? st.c1 : (ref == 2) ? st.c2 : st.c3);
                    if (!(prm.fast_homref && nonref == 0)) col_finalize(st, pos, ref, d, prm, s_tab);
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------
// window index: record range + payload span of every window (two binary searches over the sorted starts)
// ------------------------------------------------------------------------------------------------------------
__global__ void k_window_index(const TileDev d, WinRange* out) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= d.n_windows) return;
    int64_t wbeg = (int64_t)(d.win0 + w) * d.T;
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

void launch_window_index(const TileDev& d, cudaStream_t s) {
    int n = d.n_windows;
    k_window_index<<<(n + 127) / 128, 128, 0, s>>>(d, const_cast<WinRange*>(d.win));
}

void launch_pileup(const TileDev& d, const DevParams* dprm, int grid, size_t smem, cudaStream_t s) {
    static size_t configured = 0;
    if (smem > configured) {
        cudaFuncSetAttribute(k_pileup_windows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    k_pileup_windows<<<grid, kThreads, smem, s>>>(d, dprm);
}

void launch_ref_pack(const char* ascii, int64_t len, uint32_t* ref2, uint32_t* ref_n, uint32_t* ref_l, cudaStream_t s) {
    int64_t nt = (len + 31) / 32;
    k_ref_pack<<<(unsigned)((nt + 255) / 256), 256, 0, s>>>(ascii, len, ref2, ref_n, ref_l);
}

}  // namespace ngsep_b200
