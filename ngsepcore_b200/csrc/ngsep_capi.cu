// C ABI of libngsep_pileup.so (declared in include/ngsep_pileup.h): context, reference upload, tile submit/fetch.
// Host-side plumbing only — all arithmetic of the path lives in pileup_kernels.cu.  There is no CPU fallback:
// every entry point fails with NGSEP_ENODEVICE / NGSEP_ECUDA when the device cannot do the work.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pileup_kernels.cuh"

using namespace ngsep_b200;

namespace {

struct RefSeq {   // ref2 / ref_n / ref_l point INSIDE the allocations (front padding for the TMA halo of window 0)
    uint32_t* ref2 = nullptr; uint32_t* ref_n = nullptr; uint32_t* ref_l = nullptr; int64_t len = 0;
    uint32_t *a_ref2 = nullptr, *a_refn = nullptr, *a_refl = nullptr;
    void release() { if (a_ref2) cudaFree(a_ref2); if (a_refn) cudaFree(a_refn); if (a_refl) cudaFree(a_refl); *this = RefSeq(); }
};
constexpr size_t kRefFrontWords = 16;                       // 64 B in front of word 0
constexpr size_t kRefTailWords2 = kMaxWindow / 16 + 64;     // the last window is staged whole, whatever the sequence length
constexpr size_t kRefTailWords1 = kMaxWindow / 32 + 64;

template <typename T>
struct DevBuf {
    T* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 64;
        cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

constexpr size_t kMetaStride = 32;      // d_meta: [0] column-queue counters (2 x u32), then per slot: counters[2] (u64) | n_calls (u32) | pad
constexpr size_t kMaxSlots = 65536;
constexpr size_t kRing = 4;             // tiles that may be in flight between ngsep_tile_submit and ngsep_tile_fetch

struct Slot {
    DevBuf<ngsep_read_hdr> d_hdr;
    DevBuf<uint32_t> d_pay;
    DevBuf<uint32_t> d_wire;                  // NGSEP_LAYOUT_COMPACT: the payload as it crossed the bus, expanded into d_pay at staging
    DevBuf<uint8_t> d_qual;
    DevBuf<WinRange> d_win;
    DevBuf<ngsep_snv_call> d_calls;
    DevBuf<uint4> d_open;                     // single-tile launches: the tile's own open-column list (its counter sits in d_counters)
    DevBuf<uint2> d_ent;                      //                       ... and the resolved entries of its open columns
    DevBuf<uint16_t> d_group;                 // multi-sample contexts: read-group id per record
    DevBuf<uint32_t> d_gidx, d_goff, d_gblk;          //                        records indexed by read group (launch_group_index)
    DevBuf<ngsep_read_hdr> d_mhdr; DevBuf<uint8_t> d_ev;   //                     headers with quality offsets + evidence bytes (launch_build_evidence)
    DevBuf<ngsep_pop_site> d_sites;           //                        population sites (cap = tile.cap_calls) ...
    DevBuf<ngsep_sample_call> d_scalls;       //                        ... x n_samples sample calls
    DevBuf<int32_t> d_counts;
    DevBuf<int32_t> d_site_pos; DevBuf<uint16_t> d_site_alleles;   // -knownVariants: the sites of the staged tile (ngsep_tile_sites)
    std::vector<int32_t> h_site_pos; std::vector<uint16_t> h_site_alleles;
    TileDev* d_tile = nullptr;                // device copy of `tile` for single-tile launches, followed by the 16-byte staging report
    bool report_read = false;                 // the staging report has been checked on the host
    unsigned int* d_report() const { return (unsigned int*)(d_tile + 1); }
    unsigned int* d_ncalls = nullptr;         // inside Ctx::d_meta
    unsigned long long* d_counters = nullptr; // inside Ctx::d_meta
    cudaEvent_t e0 = nullptr, em = nullptr, e1 = nullptr;   // before the pileup kernel / between the two kernels / after
    cudaEvent_t e_copy = nullptr;                           // the tile's host -> device copies (copy stream) are done
    TileDev tile{};
    struct Pinned { ngsep_read_hdr sentinel; TileDev tile; unsigned char meta[32]; unsigned int rep[4]; }* pin = nullptr;   // page-locked: small copies stay asynchronous
    // A pipelined tile's results come back behind its kernels, on their stream: counters and staging report into `pin`, the first
    // kPrefetchCalls call records into pin_calls.  ngsep_tile_fetch then waits for the tile's event and reads host memory: a
    // synchronous copy at fetch time queues behind the bulk copies of the tiles submitted since (0.2 ms each, measured).
    ngsep_snv_call* pin_calls = nullptr;
    bool prefetched = false;
    bool staged = false, launched = false, in_batch = false, pipelined = false;
    int grid = 0;
    uint64_t aligned_bases = 0;
    uint64_t h2d_bytes = 0;                   // bytes of the tile copied to the device at staging
    uint64_t n_words = 0, n_qual = 0;         // payload words / quality bytes of the staged tile
    bool qual_on_host = false;                // quality bytes stay in the caller's page-locked buffer (read on demand by the genotype kernel)
    unsigned int launches = 0;
    void release() {
        d_hdr.release(); d_pay.release(); d_wire.release(); d_qual.release(); d_win.release(); d_calls.release(); d_counts.release();
        d_site_pos.release(); d_site_alleles.release();
        d_group.release(); d_sites.release(); d_scalls.release(); d_gidx.release(); d_goff.release(); d_gblk.release(); d_open.release(); d_ent.release();
        d_mhdr.release(); d_ev.release();
        if (d_tile) cudaFree(d_tile);
        if (pin) cudaFreeHost(pin);
        pin = nullptr;
        if (pin_calls) cudaFreeHost(pin_calls);
        pin_calls = nullptr;
        if (e_copy) cudaEventDestroy(e_copy);
        if (e0) cudaEventDestroy(e0);
        if (em) cudaEventDestroy(em);
        if (e1) cudaEventDestroy(e1);
        d_tile = nullptr; e0 = em = e1 = nullptr;
    }
};

struct Ctx {
    bool prefetch_results = true;   // NGSEP_PREFETCH=0: a pipelined tile's results are copied when they are fetched (A/B switch)
    double tm[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // NGSEP_TIMING: host seconds in stage / launch (submit) and event wait / counters / report / calls (fetch)
    int device = 0;
    cudaStream_t stream = nullptr;            // every kernel of the context
    cudaStream_t copy_stream = nullptr;       // bulk host -> device copies of staged tiles
    cudaStream_t geno_stream = nullptr;       // genotype kernels of single-tile launches
    cudaEvent_t t0 = nullptr, t1 = nullptr, tg = nullptr;   // stopwatch (tg joins the genotype stream)
    cudaEvent_t b0 = nullptr, bm = nullptr, b1 = nullptr;   // around / between the kernels of the batched run
    ngsep_params params{};
    DevParams* d_prm = nullptr;
    DevParams hp{};                           // host copy of *d_prm
    // sample axis (ngsep_samples_set)
    bool samples_set = false;
    DevBuf<int> d_sgo; DevBuf<uint16_t> d_sgroups; DevBuf<ngsep_sample_call> d_scratch; DevBuf<uint8_t> d_gins;
    DevBuf<uint16_t> d_gkeys; DevBuf<uint32_t> d_gvals; DevBuf<unsigned char> d_cub;   // staging scratch of launch_group_index (stream-ordered reuse)
    std::vector<int32_t> next_site_pos; std::vector<uint16_t> next_site_alleles;   // ngsep_tile_sites: sites of the next submitted tile
    bool next_has_sites = false;
    int multi_grid = 0;                       // CTAs of k_genotype_columns_multi
    bool lazy_qual = false;                   // NGSEP_F_QUALITY_IN_PLACE (NGSEP_LAZY_QUAL=0/1 overrides)
    unsigned char* d_meta = nullptr;          // kMaxSlots x kMetaStride: per-slot counters, cleared with ONE memset per batch
    std::vector<RefSeq> refs;
    std::vector<Slot*> slots;                 // slot 0 = first submit/fetch slot, 1.. = resident tiles
    std::vector<Slot*> ring;                  // submit/fetch pipeline: ring[0] == slots[0]; tiles are fetched in submission order
    size_t ring_head = 0, ring_tail = 0, ring_count = 0;
    // batched launch over every resident slot
    DevBuf<TileDev> d_btiles;
    DevBuf<WinRange> d_bwin;
    // column queue shared by every launch of this context (launches are stream-ordered)
    DevBuf<uint4> d_open;
    DevBuf<int32_t> d_cov_first, d_cov_last; DevBuf<uint8_t> d_cov_unique; DevBuf<int2> d_cov_diff; DevBuf<unsigned char> d_cov_tmp;   // ngsep_coverage_histogram
    DevBuf<unsigned long long> d_cov_hist;
    DevBuf<uint2> d_allele_out;   // ngsep_tile_allele_counts
    DevBuf<uint32_t> d_xsize, d_xoff; DevBuf<unsigned char> d_xcub;   // NGSEP_LAYOUT_COMPACT: record lengths / offsets of the full layout, scan scratch (stream-ordered reuse)
    DevBuf<unsigned char> d_ev_in; DevBuf<unsigned char> d_ev_out;   // ngsep_indel_tallies: inputs / outputs + scratch, each one allocation
    DevBuf<uint2> d_ent;     // resolved entries of the open columns (a full list only costs speed: such columns go out as record ranges)
    size_t q_min_cols = 0;   // learned from overflows
    int b_windows = 0, b_grid = 0;
    bool batch_dirty = true, batch_ran = false;
    int sm_count = 0;
    LaunchCfg cfg{1280, 6400u, 1, 4, 256, 0};   // L2 prefetch of the window after next: measured 1 % slower, off (NGSEP_L2_PREFETCH=1)
    size_t smem = 0;
    int ctas_per_sm = 1;
    int geno_grid = 0;                        // blocks of k_genotype_columns
    std::string err;
};

int fail(Ctx* c, int code, const std::string& msg) { if (c) c->err = msg; return code; }
int cuda_fail(Ctx* c, cudaError_t e, const char* what) {
    return fail(c, NGSEP_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call, what) do { cudaError_t _e = (call); if (_e != cudaSuccess) return cuda_fail(c, _e, what); } while (0)

std::string g_global_err;

int new_slot(Ctx* c, Slot** out, size_t id) {   // id = index of the slot's 32-byte counter block in d_meta
    if (c->slots.size() >= kMaxSlots - kRing) return fail(c, NGSEP_ENOMEM, "too many resident tiles");
    Slot* s = new Slot();
    s->d_counters = (unsigned long long*)(c->d_meta + (1 + id) * kMetaStride);
    s->d_ncalls = (unsigned int*)(c->d_meta + (1 + id) * kMetaStride + 16);
    cudaError_t e = cudaMalloc((void**)&s->d_tile, sizeof(TileDev) + 16);   // + the staging report (k_check_headers)
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&s->pin, sizeof(Slot::Pinned), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->e_copy, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&s->e0);
    if (e == cudaSuccess) e = cudaEventCreate(&s->em);
    if (e == cudaSuccess) e = cudaEventCreate(&s->e1);
    if (e != cudaSuccess) { s->release(); delete s; return cuda_fail(c, e, "slot setup"); }
    *out = s;
    return NGSEP_OK;
}

int clear_counts_dbg(Ctx* c, Slot* s) {
    if (s->tile.counts_dbg) {
        size_t n = (size_t)(s->tile.tile_last - s->tile.tile_first + 1) * 5;
        CK(cudaMemsetAsync(s->tile.counts_dbg, 0, n * sizeof(int32_t), c->stream), "memset counts");
    }
    return NGSEP_OK;
}

// Open-column capacity a tile asks for: a small fraction of the positions stays open after the window kernel (every
// position with the EMIT_ALL test hook).  Overflow is detected after the run: the list grows and the run is repeated.
void queue_need(const Ctx* c, const Slot* s, size_t& cols) {
    const size_t len = (size_t)(s->tile.tile_last - s->tile.tile_first + 1);
    if ((c->params.flags & NGSEP_F_EMIT_ALL) || c->params.n_samples > 1 || c->params.ploidy >= 3) cols += len + 1024;   // sample axis / pools: every candidate stays open
    else cols += len / 32 + 16384;
}
// Entry-list capacity: open columns are a fraction of a percent of the positions and hold `depth` entries each, i.e. a few
// tenths of an entry per record.  Only single-sample production runs use the list.
size_t entries_need(const Ctx* c, const Slot* s) {
    // Measured (profiles/README.md, round 2): resolving the entries in the window kernel costs it more (a warp-serial piece of
    // work in front of a block barrier, ~0.045 ms per yeast pass) than the genotype kernel saves (~0.013 ms): off unless asked for.
    static const bool on = [] { const char* e = std::getenv("NGSEP_ENTRIES"); return e && std::atoi(e) != 0; }();
    if (!on) return 0;
    if ((c->params.flags & NGSEP_F_EMIT_ALL) || c->params.n_samples > 1) return 0;
    // + the pieces the resident warps of the window kernel reserve ahead (kEntChunk = 1024 entries each)
    return std::min<size_t>((size_t)s->tile.n_reads / 2 + (size_t)c->sm_count * 64 * 1024, 0x7fffffffull);
}

int ensure_queue(Ctx* c, size_t cols) {
    cols = std::max(cols, c->q_min_cols);
    if (cols > 0x7ffffff0ull) return fail(c, NGSEP_ENOMEM, "open-column list too large (split the tile)");
    CK(c->d_open.reserve(cols), "open-column list alloc");
    return NGSEP_OK;
}

ColumnQueue queue_view(Ctx* c) {   // the batched launch over the resident tiles
    ColumnQueue q;
    q.open = c->d_open.p; q.ctr = (unsigned int*)c->d_meta; q.cap_open = (unsigned int)c->d_open.cap;
    q.ent = c->d_ent.p; q.cap_ent = (unsigned int)std::min<size_t>(c->d_ent.cap, 0x7fffffffull);
    return q;
}
// A single-tile launch has its own lists: its kernels may overlap those of the next tile.  The kernels only touch ctr[2]
// and ctr[3], which are placed on the n_open and n_ent words (bytes 20, 24) of the slot's counter block.
ColumnQueue queue_view(Slot* s) {
    ColumnQueue q;
    q.open = s->d_open.p; q.ctr = (unsigned int*)((unsigned char*)s->d_counters + 12); q.cap_open = (unsigned int)s->d_open.cap;
    q.ent = s->d_ent.p; q.cap_ent = (unsigned int)std::min<size_t>(s->d_ent.cap, 0x7fffffffull);
    return q;
}

// After a run: did the open-column list overflow?  (stream must be idle)  grew = 1: it was enlarged, repeat the run.
int queue_overflowed(Ctx* c, int* grew) {
    unsigned int ctr[3] = {0, 0, 0};
    CK(cudaMemcpy(ctr, c->d_meta, sizeof(ctr), cudaMemcpyDeviceToHost), "queue counters copy");
    *grew = 0;
    if (ctr[2] > c->d_open.cap) {
        c->q_min_cols = std::max(c->q_min_cols, (size_t)ctr[2] + (size_t)ctr[2] / 4 + 1024);
        *grew = 1;
    }
    return NGSEP_OK;
}

constexpr size_t kPrefetchCalls = 2048;   // call records of a pipelined tile that travel behind its kernels (a 2 Mb tile holds about 2000 SNV calls)

// one tile: pileup kernel + genotype kernel
int launch_slot(Ctx* c, Slot* s) {
    if (!s->staged) return fail(c, NGSEP_ESTATE, "no tile staged");
    size_t cols = 0;
    queue_need(c, s, cols);
    cols = std::max(cols, c->q_min_cols);
    if (cols > 0x7ffffff0ull) return fail(c, NGSEP_ENOMEM, "open-column list too large (split the tile)");
    CK(s->d_open.reserve(cols), "open-column list alloc");   // the slot is idle here (its previous tile has been fetched)
    CK(s->d_ent.reserve(entries_need(c, s)), "column entry list alloc");
    CK(cudaMemsetAsync(s->d_counters, 0, kMetaStride, c->stream), "memset counters");
    s->launches = 0;
    int rc = clear_counts_dbg(c, s);
    if (rc != NGSEP_OK) return rc;
    CK(cudaEventRecord(s->e0, c->stream), "event record");
    if (s->tile.n_windows > 0) launch_pileup(s->d_tile, s->d_win.p, s->tile.n_windows, c->d_prm, queue_view(s), c->cfg, s->grid, c->smem, c->stream);
    CK(cudaEventRecord(s->em, c->stream), "event record");
    // The genotype kernel runs on its own stream behind the tile's window kernel: the context stream goes on with the next
    // tile (staging kernels, window kernel) meanwhile.  Genotype launches are ordered among themselves (shared scratch).
    CK(cudaStreamWaitEvent(c->geno_stream, s->em, 0), "wait for the window kernel");
    if (s->tile.site_pos) {   // known sites: genotyped, nothing is discovered (the window kernel has only counted the covered positions)
        if (c->hp.pool_ploidy) launch_pool_sites(s->d_tile, c->d_prm, s->tile.n_sites, c->geno_grid, c->geno_stream);   // pools: the sites' evidence for the host's genotypeVariantPool
        else launch_genotype_sites(s->d_tile, c->d_prm, s->tile.n_sites, c->geno_grid, c->geno_stream);
        s->launches = s->tile.n_windows > 0 ? 2 : 1;
    } else if (s->tile.n_windows > 0) {
        if (c->params.n_samples > 1) launch_genotype_multi(s->d_tile, queue_view(s), c->d_prm, c->params.n_samples, c->multi_grid, c->geno_stream);
        else if (c->hp.pool_ploidy) launch_pool_columns(s->d_tile, queue_view(s), c->d_prm, c->geno_grid, c->geno_stream);
        else launch_genotype(s->d_tile, queue_view(s), c->d_prm, c->geno_grid, c->geno_stream);
        s->launches = 2;
    }
    s->prefetched = false;
    if (s->pipelined && c->prefetch_results) {
        static_assert(sizeof(s->pin->meta) == kMetaStride, "meta layout");
        CK(cudaMemcpyAsync(s->pin->meta, s->d_counters, kMetaStride, cudaMemcpyDeviceToHost, c->geno_stream), "counters copy");
        CK(cudaMemcpyAsync(s->pin->rep, s->d_report(), 16, cudaMemcpyDeviceToHost, c->geno_stream), "report copy");
        if (c->params.n_samples == 1 && s->tile.calls) {
            if (!s->pin_calls) CK(cudaHostAlloc((void**)&s->pin_calls, kPrefetchCalls * sizeof(ngsep_snv_call), cudaHostAllocDefault), "call prefetch buffer");
            const size_t k = std::min<size_t>(kPrefetchCalls, s->tile.cap_calls);
            CK(cudaMemcpyAsync(s->pin_calls, s->tile.calls, k * sizeof(ngsep_snv_call), cudaMemcpyDeviceToHost, c->geno_stream), "call prefetch");
        }
        s->prefetched = true;
    }
    CK(cudaEventRecord(s->e1, c->geno_stream), "event record");
    CK(cudaGetLastError(), "kernel launch");
    s->launched = true;
    s->in_batch = false;
    return NGSEP_OK;
}

// every resident tile in ONE pileup launch (the windows of all tiles form one list walked by persistent CTAs) followed
// by ONE genotype launch over the queued columns
int launch_batch(Ctx* c) {
    const size_t n = c->slots.size();
    if (n <= 1) return fail(c, NGSEP_ESTATE, "no resident tiles");
    if (c->batch_dirty) {
        std::vector<TileDev> tiles;
        size_t n_win = 0;
        for (size_t i = 1; i < n; i++) { tiles.push_back(c->slots[i]->tile); n_win += (size_t)c->slots[i]->tile.n_windows; }
        if (n_win > 0x7fffffffull) return fail(c, NGSEP_EINVAL, "too many windows in one batch");
        CK(c->d_btiles.reserve(tiles.size()), "batch tiles alloc");
        CK(c->d_bwin.reserve(n_win + 1), "batch windows alloc");
        CK(cudaMemcpyAsync(c->d_btiles.p, tiles.data(), tiles.size() * sizeof(TileDev), cudaMemcpyHostToDevice, c->stream), "batch tiles copy");
        size_t off = 0;
        for (size_t i = 1; i < n; i++) {
            launch_window_index(c->slots[i]->tile, c->d_bwin.p + off, (int)(i - 1), c->stream);
            off += (size_t)c->slots[i]->tile.n_windows;
        }
        CK(cudaGetLastError(), "window index launch");
        CK(cudaStreamSynchronize(c->stream), "batch staging");   // `tiles` lives on this stack frame
        c->b_windows = (int)n_win;
        c->b_grid = (int)std::min<size_t>(n_win, (size_t)c->sm_count * (size_t)c->ctas_per_sm);
        c->batch_dirty = false;
    }
    size_t cols = 0;
    for (size_t i = 1; i < n; i++) queue_need(c, c->slots[i], cols);
    int rc = ensure_queue(c, cols);
    if (rc != NGSEP_OK) return rc;
    {
        size_t ents = 0;
        for (size_t i = 1; i < n; i++) ents += entries_need(c, c->slots[i]) ? (size_t)c->slots[i]->tile.n_reads / 2 : 0;
        if (ents) ents += (size_t)c->sm_count * 64 * 1024;
        CK(c->d_ent.reserve(std::min<size_t>(ents, 0x7fffffffull)), "column entry list alloc");
    }
    CK(cudaMemsetAsync(c->d_meta, 0, (1 + n) * kMetaStride, c->stream), "memset counters");   // queue + every slot: ONE memset
    for (size_t i = 1; i < n; i++) { rc = clear_counts_dbg(c, c->slots[i]); if (rc != NGSEP_OK) return rc; }
    CK(cudaEventRecord(c->b0, c->stream), "event record");
    if (c->b_windows > 0) launch_pileup(c->d_btiles.p, c->d_bwin.p, c->b_windows, c->d_prm, queue_view(c), c->cfg, c->b_grid, c->smem, c->stream);
    CK(cudaEventRecord(c->bm, c->stream), "event record");
    if (c->b_windows > 0) {
        if (c->params.n_samples > 1) launch_genotype_multi(c->d_btiles.p, queue_view(c), c->d_prm, c->params.n_samples, c->multi_grid, c->stream);
        else if (c->hp.pool_ploidy) launch_pool_columns(c->d_btiles.p, queue_view(c), c->d_prm, c->geno_grid, c->stream);
        else launch_genotype(c->d_btiles.p, queue_view(c), c->d_prm, c->geno_grid, c->stream);
    }
    CK(cudaEventRecord(c->b1, c->stream), "event record");
    CK(cudaGetLastError(), "kernel launch");
    for (size_t i = 1; i < n; i++) { c->slots[i]->launched = true; c->slots[i]->in_batch = true; c->slots[i]->launches = 0; }
    c->slots[0]->launched = false;   // the batch memset cleared the first pipeline slot's counters as well
    c->ring_head = c->ring_tail = c->ring_count = 0;
    c->batch_ran = true;
    return NGSEP_OK;
}

inline double host_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Reads the staging report of a tile whose staging work has completed on the device.
int read_report(Ctx* c, Slot* s, const unsigned int* host_copy = nullptr) {
    if (s->report_read) return NGSEP_OK;
    struct { unsigned int err, pad; unsigned long long bases; } rep;
    if (host_copy) std::memcpy(&rep, host_copy, sizeof(rep));
    else CK(cudaMemcpy(&rep, s->d_report(), sizeof(rep), cudaMemcpyDeviceToHost), "report copy");
    s->report_read = true;
    s->aligned_bases = rep.bases;
    if (rep.err & 1u) return fail(c, NGSEP_EINVAL, "records are not sorted by start");
    if (rep.err & 2u) return fail(c, NGSEP_EINVAL, "payload offsets are not monotone");
    if (rep.err & 4u) return fail(c, NGSEP_EINVAL, "record payload exceeds the device window (read too long)");
    if (rep.err & 8u) return fail(c, NGSEP_EINVAL, "a record's body does not fit its payload span, or its qualities lie outside the quality array");
    return NGSEP_OK;
}

// Is [p, p + bytes) page-locked host memory the device can read in place?  (cudaHostRegister / cudaHostAlloc; the
// device alias is returned.)
const uint8_t* mapped_host_alias(const void* p, size_t bytes) {
    if (!p || !bytes) return nullptr;
    cudaPointerAttributes a0{}, a1{};
    if (cudaPointerGetAttributes(&a0, p) != cudaSuccess || cudaPointerGetAttributes(&a1, (const char*)p + bytes - 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a0.type != cudaMemoryTypeHost || a1.type != cudaMemoryTypeHost || !a0.devicePointer || !a1.devicePointer) return nullptr;
    if ((const char*)a1.devicePointer - (const char*)a0.devicePointer != (ptrdiff_t)(bytes - 1)) return nullptr;   // one contiguous mapping
    return (const uint8_t*)a0.devicePointer;
}

int stage_slot(Ctx* c, Slot* s, const ngsep_tile* t, bool allow_host_qual = false, const SynthSpec* synth = nullptr, int64_t synth_g_lo = 0) {
    if (!t) return fail(c, NGSEP_EINVAL, "null argument");
    if (t->n_reads < 0 || t->tile_first < 1 || t->tile_last < t->tile_first) return fail(c, NGSEP_EINVAL, "bad tile extents");
    if (t->seq_idx < 0 || (size_t)t->seq_idx >= c->refs.size() || !c->refs[t->seq_idx].ref2) return fail(c, NGSEP_ESTATE, "reference sequence not uploaded");
    if (!synth && t->n_reads > 0 && (!t->hdr || !t->payload || !t->qual)) return fail(c, NGSEP_EINVAL, "null tile buffers");
    if (t->n_qual > 0x3fffffff0ull) return fail(c, NGSEP_EINVAL, "tile qualities exceed 16 GiB");
    const bool multi = c->params.n_samples > 1;
    if (multi && !c->samples_set) return fail(c, NGSEP_ESTATE, "multi-sample context: call ngsep_samples_set before the first tile");
    if (multi && t->n_reads > 0 && !t->group) return fail(c, NGSEP_EINVAL, "multi-sample context: ngsep_tile.group is required");
    if (t->n_words > 0xfffffff0ull) return fail(c, NGSEP_EINVAL, "tile payload exceeds 16 GiB");
    if (t->qual_bits != 0u && t->qual_bits != 2u && t->qual_bits != 4u && t->qual_bits != 8u) return fail(c, NGSEP_EINVAL, "ngsep_tile.qual_bits must be 0, 2, 4 or 8");
    if (t->layout != NGSEP_LAYOUT_FULL && t->layout != NGSEP_LAYOUT_COMPACT) return fail(c, NGSEP_EINVAL, "ngsep_tile.layout must be NGSEP_LAYOUT_FULL or NGSEP_LAYOUT_COMPACT");
    const bool compact = t->layout == NGSEP_LAYOUT_COMPACT && !synth;
    if (compact && t->n_wire_words > t->n_words) return fail(c, NGSEP_EINVAL, "ngsep_tile.n_wire_words exceeds n_words");
    const size_t wire_words = compact ? (size_t)t->n_wire_words : (size_t)t->n_words;   // what `payload` holds on the caller's side
    const unsigned int qual_bits = t->qual_bits ? t->qual_bits : 8u;
    CK(cudaSetDevice(c->device), "set device");
    const RefSeq& ref = c->refs[t->seq_idx];
    // Quality bytes are only read for the positions that reach k_genotype_columns (a fraction of a percent of a
    // single-sample tile): when the caller's quality buffer is page-locked (ngsep_host_register) a pipelined tile leaves
    // it where it is and the kernel reads the few bytes it needs in place, over PCIe.  Multi-sample contexts and the
    // EMIT_ALL test hook evaluate every column: they always copy.
    const uint8_t* host_qual = nullptr;
    if (allow_host_qual && c->lazy_qual && !multi && !compact && !(c->params.flags & NGSEP_F_EMIT_ALL) && t->n_reads > 0) host_qual = mapped_host_alias(t->qual, (size_t)t->n_qual);
    s->qual_on_host = host_qual != nullptr;
    {   // pipeline slots grow together: a slot takes the largest capacity any of them has needed so far, so that cycling
        // tiles of different sizes through the ring stops reallocating after the first pass
        size_t need_h = (size_t)t->n_reads + 1, need_p = (size_t)t->n_words + 16, need_q = host_qual ? 0 : (size_t)t->n_qual + 16;
        if (s->pipelined) for (Slot* r : c->ring) { need_h = std::max(need_h, r->d_hdr.cap); need_p = std::max(need_p, r->d_pay.cap); need_q = std::max(need_q, r->d_qual.cap); }
        if (need_h > s->d_hdr.cap) { s->d_hdr.release(); CK(cudaMalloc((void**)&s->d_hdr.p, need_h * sizeof(ngsep_read_hdr)), "hdr alloc"); s->d_hdr.cap = need_h; }
        if (need_p > s->d_pay.cap) { s->d_pay.release(); CK(cudaMalloc((void**)&s->d_pay.p, need_p * sizeof(uint32_t)), "payload alloc"); s->d_pay.cap = need_p; }
        if (need_q > s->d_qual.cap) { s->d_qual.release(); CK(cudaMalloc((void**)&s->d_qual.p, need_q), "quality alloc"); s->d_qual.cap = need_q; }
    }
    if (multi) CK(s->d_group.reserve((size_t)t->n_reads + 16), "read-group alloc");
    // The bulk copies run on the COPY stream, the kernels on the context stream: the copies of the next tile overlap the
    // kernels of this one (a slot's buffers are only reused after its tile has been fetched).  Pageable sources are staged by
    // the driver synchronously with the host thread anyway: they stay on the context stream.
    const bool pinned_src = !synth && t->n_reads > 0 && mapped_host_alias(t->payload, wire_words * 4) != nullptr;
    if (compact) CK(s->d_wire.reserve(wire_words + 16), "wire payload alloc");
    cudaStream_t copy_stream = pinned_src ? c->copy_stream : c->stream;
    if (multi && t->n_reads > 0) CK(cudaMemcpyAsync(s->d_group.p, t->group, (size_t)t->n_reads * sizeof(uint16_t), cudaMemcpyHostToDevice, copy_stream), "read-group copy");
    if (synth) launch_synth_tile(*synth, synth_g_lo, t->n_reads, ref.ref2, s->d_hdr.p, s->d_pay.p, s->d_qual.p, copy_stream);   // generated in place: nothing crosses PCIe
    else if (t->n_reads > 0) {
        CK(cudaMemcpyAsync(s->d_hdr.p, t->hdr, (size_t)t->n_reads * sizeof(ngsep_read_hdr), cudaMemcpyHostToDevice, copy_stream), "hdr copy");
        CK(cudaMemcpyAsync(compact ? s->d_wire.p : s->d_pay.p, t->payload, wire_words * 4, cudaMemcpyHostToDevice, copy_stream), "payload copy");
        if (!host_qual) CK(cudaMemcpyAsync(s->d_qual.p, t->qual, (size_t)t->n_qual, cudaMemcpyHostToDevice, copy_stream), "quality copy");
    }
    s->pin->sentinel = ngsep_read_hdr{INT_MAX, 0u, 0, 0, (uint32_t)wire_words};   // (compact: rewritten to n_words by the expansion)
    CK(cudaMemcpyAsync(s->d_hdr.p + t->n_reads, &s->pin->sentinel, sizeof(ngsep_read_hdr), cudaMemcpyHostToDevice, copy_stream), "sentinel copy");
    if (pinned_src) {
        CK(cudaEventRecord(s->e_copy, c->copy_stream), "copy event");
        CK(cudaStreamWaitEvent(c->stream, s->e_copy, 0), "wait for the copies");
    }
    if (multi) {   // records indexed by read group; the scratch buffers are shared by the context (growing them waits for the stream)
        const size_t n = (size_t)t->n_reads;
        const size_t cub_bytes = group_index_temp_bytes(t->n_reads);
        if (c->d_gkeys.cap < 2 * n + 16 || c->d_gvals.cap < n + 16 || c->d_cub.cap < cub_bytes + 256) {
            CK(cudaStreamSynchronize(c->stream), "drain before growing the staging scratch");
            CK(c->d_gkeys.reserve(2 * n + 16), "group key alloc");
            CK(c->d_gvals.reserve(n + 16), "group value alloc");
            CK(c->d_cub.reserve(cub_bytes + 256), "sort scratch alloc");
        }
        CK(s->d_gidx.reserve(n + 16), "group index alloc");
        CK(s->d_mhdr.reserve(n + 16), "evidence header alloc");
        CK(s->d_ev.reserve((size_t)t->n_qual * (8u / qual_bits) + 64), "evidence alloc");
        CK(s->d_goff.reserve((size_t)c->hp.n_groups + 4), "group offsets alloc");
        CK(s->d_gblk.reserve(((size_t)c->hp.n_groups + 1) * ((size_t)group_blocks(t->n_reads) + 1) + 16), "group block table alloc");
        CK(launch_group_index(s->d_group.p, c->d_gins.p, t->n_reads, c->hp.n_groups, c->d_gkeys.p, c->d_gvals.p, c->d_cub.p, c->d_cub.cap, s->d_gidx.p, s->d_goff.p, s->d_gblk.p, c->stream),
           "group index launch");
    }
    s->h2d_bytes = synth ? 0 : (uint64_t)t->n_reads * sizeof(ngsep_read_hdr) + (uint64_t)wire_words * 4 + (host_qual ? 0 : (uint64_t)t->n_qual) +
                   (multi ? (uint64_t)t->n_reads * 2 : 0);
    s->n_words = (uint64_t)t->n_words; s->n_qual = (uint64_t)t->n_qual;
    // ordering, offsets, stage fit and the NGSEP_M_SIMPLE promise are checked on the device copy (the host thread only
    // enqueues work: the next tile can be packed / copied meanwhile); a tile that fails gets an empty window index
    CK(cudaMemsetAsync(s->d_report(), 0, 16, c->stream), "report memset");
    if (compact && t->n_reads > 0) {   // the wire stream -> the full layout, before anything looks at the payload
        const size_t n = (size_t)t->n_reads, tmp = expand_scan_temp_bytes(t->n_reads);
        if (c->d_xsize.cap < n + 2 || c->d_xoff.cap < n + 2 || c->d_xcub.cap < tmp + 256) {
            CK(cudaStreamSynchronize(c->stream), "drain before growing the expansion scratch");
            CK(c->d_xsize.reserve(n + 2), "expansion alloc"); CK(c->d_xoff.reserve(n + 2), "expansion alloc"); CK(c->d_xcub.reserve(tmp + 256), "expansion alloc");
        }
        uint32_t dict4[4];
        std::memcpy(dict4, t->qual_dict, 16);
        CK(launch_expand_compact(s->d_hdr.p, t->n_reads, s->d_wire.p, (unsigned long long)wire_words, s->d_pay.p, (unsigned long long)t->n_words, s->d_qual.p,
                                 (unsigned long long)t->n_qual, qual_bits, dict4, c->d_xsize.p, c->d_xoff.p, c->d_xcub.p, c->d_xcub.cap, s->d_report(), c->stream),
           "payload expansion");
    }
    launch_check_headers(s->d_hdr.p, s->d_pay.p, t->n_reads, (unsigned long long)t->n_words, (unsigned long long)t->n_qual, qual_bits, c->cfg.pay_cap_words, s->d_report(), c->stream);
    s->report_read = false;
    s->aligned_bases = 0;
    // window length: a multiple of 128, shrunk while an average window would not fit one stage
    int T = c->cfg.T_lay;
    int64_t tile_len = (int64_t)t->tile_last - t->tile_first + 1;
    double words_per_pos = (double)t->n_words / (double)tile_len;
    double recs_per_pos = (double)t->n_reads / (double)tile_len;
    const int span = t->max_span < 1 ? 1 : t->max_span;
    while (T > 128 && (words_per_pos * (T + span) > 0.85 * c->cfg.pay_cap_words || recs_per_pos * (T + span) > 0.85 * (c->cfg.pay_cap_words / 12)))
        T -= 128;   // the largest multiple of 128 whose average window fits (60x: 512, not the 256 that halving gave)
    TileDev& d = s->tile;
    d = TileDev{};
    d.hdr = s->d_hdr.p; d.payload = s->d_pay.p; d.group = multi ? s->d_group.p : nullptr; d.qual = host_qual ? host_qual : s->d_qual.p;
    d.qual_bits = qual_bits;
    std::memcpy(d.qual_dict, t->qual_dict, 16);
    d.gidx = multi ? s->d_gidx.p : nullptr; d.goff = multi ? s->d_goff.p : nullptr;
    d.gblk = multi ? s->d_gblk.p : nullptr; d.n_gblk = multi ? group_blocks(t->n_reads) : 0;
    d.mhdr = multi ? s->d_mhdr.p : nullptr; d.ev = multi ? s->d_ev.p : nullptr; d.ev_mul = 32u / qual_bits;
    d.n_reads = t->n_reads; d.max_span = span; d.max_back = t->max_back < 0 ? 0 : t->max_back;
    d.tile_first = t->tile_first; d.tile_last = t->tile_last; d.seq_idx = t->seq_idx;
    d.ref2 = ref.ref2; d.ref_n = ref.ref_n; d.ref_l = ref.ref_l; d.ref_len = ref.len;
    d.T = T;
    d.win0 = (t->tile_first - 1) / T;
    d.n_windows = (t->n_reads == 0) ? 0 : ((t->tile_last - 1) / T - d.win0 + 1);
    CK(s->d_win.reserve((size_t)d.n_windows + 1), "window alloc");
    if (multi) {   // population sites + n_samples sample calls each; grown (and the tile rerun) when a tile needs more
        // (a site costs n_samples x 104 bytes: start at 0.2 % of the positions, not more — the fetch grows the buffers and
        // reruns the tile in the rare case a tile reports more sites)
        size_t cap_sites = std::max<size_t>(s->d_sites.cap, (size_t)(tile_len / 512) + 4096);
        if (s->d_sites.cap < cap_sites) CK(s->d_sites.reserve(cap_sites), "site alloc");
        CK(s->d_scalls.reserve(s->d_sites.cap * (size_t)c->params.n_samples), "sample call alloc");
        d.pop_sites = s->d_sites.p; d.pop_calls = s->d_scalls.p; d.calls = nullptr; d.n_calls = s->d_ncalls; d.cap_calls = (unsigned)s->d_sites.cap;
    } else {
        size_t cap_calls = (size_t)(tile_len / 64) + 16384;   // ~20x the expected SNV density; a tile that needs more grows the buffer and reruns
        if (s->d_calls.cap < cap_calls) CK(s->d_calls.reserve(cap_calls), "calls alloc");
        d.calls = s->d_calls.p; d.n_calls = s->d_ncalls; d.cap_calls = (unsigned)s->d_calls.cap;
    }
    d.site_pos = nullptr; d.site_alleles = nullptr; d.n_sites = 0;
    if (s->pipelined && c->next_has_sites) {   // ngsep_tile_sites: this tile is genotyped at the listed sites
        s->h_site_pos.swap(c->next_site_pos); s->h_site_alleles.swap(c->next_site_alleles);
        c->next_site_pos.clear(); c->next_site_alleles.clear(); c->next_has_sites = false;
        const size_t ns = s->h_site_pos.size();
        for (size_t k = 0; k < ns; k++)
            if (s->h_site_pos[k] < t->tile_first || s->h_site_pos[k] > t->tile_last) return fail(c, NGSEP_EINVAL, "ngsep_tile_sites: a site lies outside the tile");
        CK(s->d_site_pos.reserve(ns + 1), "site alloc");
        CK(s->d_site_alleles.reserve(ns + 1), "site alloc");
        if (ns) {
            CK(cudaMemcpyAsync(s->d_site_pos.p, s->h_site_pos.data(), ns * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream), "site copy");
            CK(cudaMemcpyAsync(s->d_site_alleles.p, s->h_site_alleles.data(), ns * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream), "site copy");
        }
        if (s->d_calls.cap < ns + 16) { CK(s->d_calls.reserve(ns + 16), "calls alloc"); d.calls = s->d_calls.p; d.cap_calls = (unsigned)s->d_calls.cap; }
        d.site_pos = s->d_site_pos.p; d.site_alleles = s->d_site_alleles.p; d.n_sites = (int)ns;
        s->h2d_bytes += (uint64_t)ns * 6;
    }
    d.counts_dbg = nullptr;
    if (c->params.flags & NGSEP_F_EMIT_ALL) {
        CK(s->d_counts.reserve((size_t)tile_len * 5), "counts alloc");
        d.counts_dbg = s->d_counts.p;
    }
    d.counters = s->d_counters;
    d.report = s->d_report();
    s->grid = std::min(d.n_windows, c->sm_count * c->ctas_per_sm);   // persistent CTAs
    s->pin->tile = s->tile;
    CK(cudaMemcpyAsync(s->d_tile, &s->pin->tile, sizeof(TileDev), cudaMemcpyHostToDevice, c->stream), "tile descriptor copy");
    if (multi) launch_build_evidence(d, s->d_mhdr.p, s->d_ev.p, c->params.max_base_qs, c->stream);   // (after the staging check: it trusts the offsets)
    launch_window_index(d, s->d_win.p, 0, c->stream);   // depends on the staged records only: built once per staging
    CK(cudaGetLastError(), "window index launch");
    s->staged = true;
    s->launched = false;
    return NGSEP_OK;
}

// Single-sample contexts return ngsep_snv_call records in `out`; multi-sample contexts return population sites in
// `sites` and their n_samples sample calls each in `scalls`.
int fetch_slot(Ctx* c, Slot* s, ngsep_snv_call* out, size_t cap, size_t* n_out, int32_t* counts_dbg, ngsep_tile_stats* stats,
               ngsep_pop_site* sites = nullptr, ngsep_sample_call* scalls = nullptr) {
    if (!n_out) return fail(c, NGSEP_EINVAL, "null argument");
    const bool multi = c->params.n_samples > 1;
    if (multi && out) return fail(c, NGSEP_ESTATE, "multi-sample context: use ngsep_tile_fetch_pop / ngsep_resident_fetch_pop");
    if (!multi && (sites || scalls)) return fail(c, NGSEP_ESTATE, "single-sample context: use ngsep_tile_fetch / ngsep_resident_fetch");
    if (!s->launched) return fail(c, NGSEP_ESTATE, "no tile in flight");
    CK(cudaSetDevice(c->device), "set device");
    struct Meta { unsigned long long ctr[2]; unsigned int n_calls, n_open, pad[2]; } meta;
    static_assert(sizeof(Meta) == kMetaStride, "meta layout");
    for (int attempt = 0; attempt < 4; attempt++) {
        // a pipelined tile waits for ITS kernels only: later tiles may already be copying on the same stream
        const double tf0 = host_now();
        if (s->in_batch) CK(cudaStreamSynchronize(c->stream), "tile kernels");
        else CK(cudaEventSynchronize(s->e1), "tile kernels");
        const double tf1 = host_now();
        const bool pre = s->prefetched && !s->in_batch;   // counters, report and the first calls are in page-locked host memory already
        if (pre) std::memcpy(&meta, s->pin->meta, sizeof(meta));
        else CK(cudaMemcpy(&meta, s->d_counters, sizeof(meta), cudaMemcpyDeviceToHost), "counters copy");
        const double tf2 = host_now();
        { int rrc = read_report(c, s, pre ? s->pin->rep : nullptr); if (rrc != NGSEP_OK) return rrc; }
        c->tm[2] += tf1 - tf0; c->tm[3] += tf2 - tf1; c->tm[4] += host_now() - tf2;
        int grew = 0;
        if (s->in_batch) {
            int rc = queue_overflowed(c, &grew);
            if (rc != NGSEP_OK) return rc;
        } else if (meta.n_open > s->d_open.cap) {
            c->q_min_cols = std::max(c->q_min_cols, (size_t)meta.n_open + (size_t)meta.n_open / 4 + 1024);
            grew = 1;
        }
        if (meta.n_calls <= s->tile.cap_calls && !grew) break;
        if (attempt == 3) return fail(c, NGSEP_EOVERFLOW, "call / column buffer overflow");
        if (meta.n_calls > s->tile.cap_calls) {   // grow and rerun the staged tile
            if (!s->in_batch) CK(cudaStreamSynchronize(c->stream), "drain before growing the call buffers");
            if (multi) {
                CK(s->d_sites.reserve((size_t)meta.n_calls + 1024), "site realloc");
                CK(s->d_scalls.reserve(s->d_sites.cap * (size_t)c->params.n_samples), "sample call realloc");
                s->tile.pop_sites = s->d_sites.p; s->tile.pop_calls = s->d_scalls.p; s->tile.cap_calls = (unsigned)s->d_sites.cap;
            } else {
                CK(s->d_calls.reserve((size_t)meta.n_calls + 1024), "calls realloc");
                s->tile.calls = s->d_calls.p; s->tile.cap_calls = (unsigned)s->d_calls.cap;
            }
            CK(cudaMemcpy(s->d_tile, &s->tile, sizeof(TileDev), cudaMemcpyHostToDevice), "tile descriptor copy");
            c->batch_dirty = true;
        }
        int rc = s->in_batch ? launch_batch(c) : launch_slot(c, s);
        if (rc != NGSEP_OK) return rc;
    }
    const unsigned int n = meta.n_calls;
    // device -> host bytes of this fetch: counters + staging report, and the call records (a pipelined tile's first slots came behind its kernels)
    size_t d2h = kMetaStride + 16;
    if (s->prefetched && !s->in_batch && s->pin_calls) d2h += std::min<size_t>(kPrefetchCalls, s->tile.cap_calls) * sizeof(ngsep_snv_call);
    if (stats) {
        float ms = 0.f, pms = 0.f;
        if (s->in_batch) {   // the whole batch: its tiles share the launches
            CK(cudaEventElapsedTime(&ms, c->b0, c->b1), "event elapsed");
            CK(cudaEventElapsedTime(&pms, c->b0, c->bm), "event elapsed");
        } else {
            CK(cudaEventElapsedTime(&ms, s->e0, s->e1), "event elapsed");
            CK(cudaEventElapsedTime(&pms, s->e0, s->em), "event elapsed");
        }
        stats->covered_positions = meta.ctr[0];
        stats->candidate_positions = meta.ctr[1];
        stats->calls = n;
        stats->aligned_bases = s->aligned_bases;
        stats->h2d_bytes = s->h2d_bytes;
        stats->quality_in_place = s->qual_on_host ? 1u : 0u;
        stats->d2h_bytes = 0;   // (completed below, once the call records have been brought back)
        stats->kernel_ms = ms;
        stats->pileup_ms = pms;
        stats->kernel_launches = s->launches;
        if (s->in_batch) { unsigned int qc[3] = {0, 0, 0}; CK(cudaMemcpy(qc, c->d_meta, sizeof(qc), cudaMemcpyDeviceToHost), "queue counters copy"); stats->open_columns = qc[2]; }
        else stats->open_columns = meta.n_open;
    }
    *n_out = n;
    if (n > cap) return fail(c, NGSEP_EOVERFLOW, "caller's call buffer is too small; n_out holds the required capacity");
    if (n > 0 && multi) {
        if (!sites || !scalls) return fail(c, NGSEP_EINVAL, "null output buffer");
        CK(cudaMemcpy(sites, s->d_sites.p, (size_t)n * sizeof(ngsep_pop_site), cudaMemcpyDeviceToHost), "site copy");
        CK(cudaMemcpy(scalls, s->d_scalls.p, (size_t)n * (size_t)c->params.n_samples * sizeof(ngsep_sample_call), cudaMemcpyDeviceToHost), "sample call copy");
        d2h += (size_t)n * (sizeof(ngsep_pop_site) + (size_t)c->params.n_samples * sizeof(ngsep_sample_call));
    } else if (n > 0) {
        if (!out) return fail(c, NGSEP_EINVAL, "null output buffer");
        const double tc0 = host_now();
        size_t have = 0;
        if (s->prefetched && !s->in_batch && s->pin_calls) { have = std::min<size_t>(n, kPrefetchCalls); std::memcpy(out, s->pin_calls, have * sizeof(ngsep_snv_call)); }
        if (n > have) CK(cudaMemcpy(out + have, s->d_calls.p + have, ((size_t)n - have) * sizeof(ngsep_snv_call), cudaMemcpyDeviceToHost), "calls copy");
        d2h += ((size_t)n - have) * sizeof(ngsep_snv_call);
        c->tm[5] += host_now() - tc0;
    }
    if (counts_dbg && s->tile.counts_dbg) {
        size_t m = (size_t)(s->tile.tile_last - s->tile.tile_first + 1) * 5;
        CK(cudaMemcpy(counts_dbg, s->tile.counts_dbg, m * sizeof(int32_t), cudaMemcpyDeviceToHost), "counts copy");
        d2h += m * sizeof(int32_t);
    }
    if (stats) stats->d2h_bytes = (uint32_t)std::min<size_t>(d2h, 0xffffffffull);
    return NGSEP_OK;
}

}  // namespace

extern "C" {

int ngsep_abi_version(void) { return 2004; }
int ngsep_built_sm(void) { return 100; }

const char* ngsep_last_error(void* ctx) {
    if (!ctx) return g_global_err.c_str();
    return ((Ctx*)ctx)->err.c_str();
}

int ngsep_ctx_create(int device, const ngsep_params* params, const double* tables, size_t n_tables, void** out) {
    if (!out || !params || !tables) { g_global_err = "null argument"; return NGSEP_EINVAL; }
    *out = nullptr;
    if (n_tables != NGSEP_N_TABLES) { g_global_err = "tables must hold NGSEP_N_TABLES doubles"; return NGSEP_EINVAL; }
    if (params->ploidy >= 3 && params->n_samples > 1) { g_global_err = "ploidy >= 3 (pool algorithm, SSVPL:240-254) is built for single-sample contexts only"; return NGSEP_EINVAL; }
    if (params->n_samples < 1 || params->n_samples > 65535) { g_global_err = "n_samples must be in 1..65535"; return NGSEP_EINVAL; }
    if (params->max_base_qs < 4 || params->max_base_qs > 30) { g_global_err = "max_base_qs must be in 4..30 (CountsHelper tables have 31 rows)"; return NGSEP_EINVAL; }
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) { g_global_err = std::string("no CUDA device: ") + cudaGetErrorString(e); return NGSEP_ENODEVICE; }
    if (device < 0 || device >= n_dev) { g_global_err = "device index out of range"; return NGSEP_ENODEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { g_global_err = "cannot query device"; return NGSEP_ENODEVICE; }
    if (prop.major != 10) { g_global_err = "libngsep_pileup.so is built for sm_100a only; found sm_" + std::to_string(prop.major * 10 + prop.minor); return NGSEP_ENODEVICE; }
    Ctx* c = new Ctx();
    c->device = device;
    c->params = *params;
    c->sm_count = prop.multiProcessorCount;
    // launch shape (tunable for experiments; the defaults are what bench.py measures)
    if (const char* s = std::getenv("NGSEP_WINDOW")) { int v = std::atoi(s); if (v >= 128 && (v % 128) == 0 && v <= kMaxWindow) c->cfg.T_lay = v; }
    if (const char* s = std::getenv("NGSEP_PAY_CAP_WORDS")) { int v = std::atoi(s); if (v >= 1024 && v <= 50000) c->cfg.pay_cap_words = (unsigned)(v & ~3); }
    if (const char* s = std::getenv("NGSEP_STAGES")) { int v = std::atoi(s); if (v == 1 || v == 2) c->cfg.n_stages = v; }
    if (const char* s = std::getenv("NGSEP_L2_PREFETCH")) c->cfg.l2_prefetch = std::atoi(s) != 0;
    if (const char* s = std::getenv("NGSEP_PREFETCH")) c->prefetch_results = std::atoi(s) != 0;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->geno_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->t0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->t1);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->tg, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&c->b0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->bm);
    if (e == cudaSuccess) e = cudaEventCreate(&c->b1);
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_prm, sizeof(DevParams));
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_meta, (kMaxSlots + 1) * kMetaStride);
    if (e == cudaSuccess) e = cudaMemset(c->d_meta, 0, (kMaxSlots + 1) * kMetaStride);
    if (e != cudaSuccess) { g_global_err = std::string("context setup: ") + cudaGetErrorString(e); delete c; return NGSEP_ECUDA; }
    c->smem = pileup_smem_bytes(c->cfg);
    if (c->smem > (size_t)prop.sharedMemPerBlockOptin) { g_global_err = "window / stage configuration exceeds the shared memory of one SM"; delete c; return NGSEP_EINVAL; }
    if (const char* s = std::getenv("NGSEP_THREADS")) { int v = std::atoi(s); if (v == 128 || v == 256) c->cfg.threads = v; }
    if (c->cfg.threads == 128) c->cfg.min_blocks = 8;
    if (const char* s = std::getenv("NGSEP_MIN_BLOCKS")) { int v = std::atoi(s); if (v >= 3 && v <= 12) c->cfg.min_blocks = v; }
    c->ctas_per_sm = pileup_max_ctas_per_sm(c->smem, c->cfg.threads, c->cfg.min_blocks);
    if (c->ctas_per_sm < 1) { g_global_err = "the pileup kernel cannot be made resident with this configuration"; delete c; return NGSEP_ECUDA; }
    if (const char* s = std::getenv("NGSEP_CTAS_PER_SM")) { int v = std::atoi(s); if (v >= 1 && v < c->ctas_per_sm) c->ctas_per_sm = v; }
    c->lazy_qual = (params->flags & NGSEP_F_QUALITY_IN_PLACE) != 0;
    if (const char* s = std::getenv("NGSEP_LAZY_QUAL")) c->lazy_qual = std::atoi(s) != 0;
    c->geno_grid = c->sm_count * 4;   // one wave of resident CTAs (64 registers): a second wave pays the table prologue again (measured: 0.0462 -> 0.0443 ms)
    if (const char* s = std::getenv("NGSEP_GENO_BLOCKS_PER_SM")) { int v = std::atoi(s); if (v >= 1 && v <= 256) c->geno_grid = c->sm_count * v; }
    Slot* s0 = nullptr;
    if (new_slot(c, &s0, 0) != NGSEP_OK) { g_global_err = c->err; delete c; return NGSEP_ECUDA; }
    c->slots.push_back(s0);
    s0->pipelined = true;
    c->ring.push_back(s0);
    for (size_t k = 1; k < kRing; k++) {   // counter blocks of the other ring slots sit at the end of d_meta
        Slot* sk = nullptr;
        if (new_slot(c, &sk, kMaxSlots - k) != NGSEP_OK) { g_global_err = c->err; delete c; return NGSEP_ECUDA; }
        sk->pipelined = true;
        c->ring.push_back(sk);
    }
    DevParams hp{};
    hp.max_base_qs = params->max_base_qs;
    hp.min_quality = params->min_quality;
    hp.flags = params->flags;
    std::memcpy(hp.tables, tables, sizeof(double) * NGSEP_N_TABLES);
    // Columns without any counted non-reference base are homozygous reference whenever the (ref,ref) event
    // dominates: a_q > c_q > b_q for every q in 4..30 and log prior(hom) - log prior(het) comfortably positive.
    hp.fast_homref = (tables[93] - tables[94] > 1.0) ? 1 : 0;
    for (int q = 4; q <= 30 && hp.fast_homref; q++) if (!(tables[3 * q] > tables[3 * q + 2] && tables[3 * q + 2] > tables[3 * q + 1])) hp.fast_homref = 0;
    // Proven skip: with k counted reference bases and m counted non-reference bases, the 15 events other
    // than (ref,ref) sum, relative to it, to at most
    //   B(k,m) = 6*10^(-k*dca + m*gcb + D) + 3*10^(-k*dab + m*gab) + 6*10^(-k*dab + m*gcb + D)
    // with dca = min_q(a_q - c_q), gcb = max_q(c_q - b_q), dab = min_q(a_q - b_q), gab = max_q(a_q - b_q) over the
    // qualities that can be counted (4..maxBaseQS) and D = log prior(het) - log prior(hom).  B <= 0.9 gives
    // P(ref,ref) >= 0.526 and every other genotype <= 0.474 < P(ref,ref) + 0.01: VD.getIndexesMaxGenotype keeps (ref,ref).
    {
        double dca = 1e300, gcb = -1e300, dab = 1e300, gab = -1e300;
        for (int q = 4; q <= params->max_base_qs; q++) {
            double a = tables[3 * q], b = tables[3 * q + 1], cc = tables[3 * q + 2];
            dca = std::min(dca, a - cc); gcb = std::max(gcb, cc - b); dab = std::min(dab, a - b); gab = std::max(gab, a - b);
        }
        const double D = tables[94] - tables[93];
        for (int m = 0; m < 16; m++) {
            hp.k_min[m] = INT_MAX;
            if (!hp.fast_homref || !(dca > 0 && dab > 0)) continue;
            for (int k = 0; k <= 200000; k++) {
                double B = 6 * std::pow(10.0, -k * dca + m * gcb + D) + 3 * std::pow(10.0, -k * dab + m * gab) + 6 * std::pow(10.0, -k * dab + m * gcb + D);
                if (B <= 0.9) { hp.k_min[m] = k; break; }
            }
        }
    }
    for (int i = 0; i < kTab2Size; i++) {   // DevParams.tab2: the genotype kernels' per-lane table, from the host's tables
        const int l = i % kTabLanes, row = i / kTabLanes, q = row >> 2, b = row & 3;
        const int pa = (l < 4) ? l : ((l < 7) ? 0 : ((l < 9) ? 1 : 2));
        const int pb = (l < 4) ? l : ((l < 7) ? l - 3 : ((l < 9) ? l - 5 : 3));
        double v = 0.0;
        if (row < kTabZeroRow) v = (b == pa || b == pb) ? ((l < 4) ? tables[3 * q] : tables[3 * q + 2]) : tables[3 * q + 1];
        hp.tab2[i] = v;
    }
    hp.pool_ploidy = params->ploidy >= 3 ? params->ploidy : 0;
    hp.n_samples = params->n_samples;
    hp.column_run = 1;
    if (const char* s = std::getenv("NGSEP_MULTI_RUN")) { int v = std::atoi(s); if (v >= 1 && v <= 256) hp.column_run = v; }
    hp.multi_ordered = 1;
    if (const char* s = std::getenv("NGSEP_MULTI_ORDERED")) hp.multi_ordered = std::atoi(s) != 0;
    hp.sample_min_quality = 40;   // the per-sample listener of MSVD.genotypeVariant is a fresh one: SSVPL.DEF_MIN_QUALITY (MSVD:676-678, SSVPL:50)
    c->hp = hp;
    c->multi_grid = c->sm_count * 8;
    if (const char* s = std::getenv("NGSEP_MULTI_BLOCKS_PER_SM")) { int v = std::atoi(s); if (v >= 1 && v <= 16) c->multi_grid = c->sm_count * v; }
    e = cudaMemcpy(c->d_prm, &hp, sizeof(hp), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { g_global_err = std::string("params upload: ") + cudaGetErrorString(e); delete c; return NGSEP_ECUDA; }
    *out = c;
    return NGSEP_OK;
}

int ngsep_ctx_destroy(void* ctx) {
    Ctx* c = (Ctx*)ctx;
    if (!c) return NGSEP_EINVAL;
    if (std::getenv("NGSEP_TIMING"))
        std::fprintf(stderr, "context: host seconds in submit: stage %.4f launch %.4f; in fetch: event wait %.4f counters %.4f report %.4f calls %.4f\n", c->tm[0], c->tm[1], c->tm[2], c->tm[3],
                     c->tm[4], c->tm[5]);
    cudaSetDevice(c->device);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    if (c->geno_stream) cudaStreamSynchronize(c->geno_stream);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto& r : c->refs) r.release();
    for (Slot* s : c->slots) { s->release(); delete s; }
    for (size_t k = 1; k < c->ring.size(); k++) { c->ring[k]->release(); delete c->ring[k]; }
    c->d_btiles.release(); c->d_bwin.release(); c->d_open.release(); c->d_ent.release();
    c->d_cov_first.release(); c->d_cov_last.release(); c->d_cov_unique.release(); c->d_cov_diff.release(); c->d_cov_tmp.release(); c->d_cov_hist.release();
    c->d_sgo.release(); c->d_sgroups.release(); c->d_scratch.release(); c->d_gins.release(); c->d_gkeys.release(); c->d_gvals.release(); c->d_cub.release();
    if (c->d_prm) cudaFree(c->d_prm);
    if (c->d_meta) cudaFree(c->d_meta);
    if (c->t0) cudaEventDestroy(c->t0);
    if (c->t1) cudaEventDestroy(c->t1);
    if (c->tg) cudaEventDestroy(c->tg);
    if (c->b0) cudaEventDestroy(c->b0);
    if (c->bm) cudaEventDestroy(c->bm);
    if (c->b1) cudaEventDestroy(c->b1);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->geno_stream) cudaStreamDestroy(c->geno_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return NGSEP_OK;
}

int ngsep_ref_upload(void* ctx, int32_t seq_idx, const char* bases, int64_t length) {
    Ctx* c = (Ctx*)ctx;
    if (!c || !bases || seq_idx < 0 || length <= 0) return fail(c, NGSEP_EINVAL, "bad reference arguments");
    CK(cudaSetDevice(c->device), "set device");
    if ((size_t)seq_idx >= c->refs.size()) c->refs.resize(seq_idx + 1);
    RefSeq& r = c->refs[seq_idx];
    r.release();
    c->batch_dirty = true;
    int64_t n32 = (length + 31) / 32;
    const size_t w2 = kRefFrontWords + (size_t)n32 * 2 + kRefTailWords2, w1 = kRefFrontWords + (size_t)n32 + kRefTailWords1;
    char* d_ascii = nullptr;
    CK(cudaMalloc((void**)&d_ascii, (size_t)length), "ref ascii alloc");
    cudaError_t e = cudaMalloc((void**)&r.a_ref2, w2 * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&r.a_refn, w1 * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&r.a_refl, w1 * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_ascii, bases, (size_t)length, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(r.a_ref2, 0, w2 * 4, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(r.a_refn, 0xff, w1 * 4, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(r.a_refl, 0, w1 * 4, c->stream);
    if (e == cudaSuccess) {
        r.ref2 = r.a_ref2 + kRefFrontWords; r.ref_n = r.a_refn + kRefFrontWords; r.ref_l = r.a_refl + kRefFrontWords;
        launch_ref_pack(d_ascii, length, r.ref2, r.ref_n, r.ref_l, c->stream);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_ascii);
    if (e != cudaSuccess) { r.release(); return cuda_fail(c, e, "reference upload"); }
    r.len = length;
    return NGSEP_OK;
}

int ngsep_tile_submit(void* ctx, const ngsep_tile* t) {
    Ctx* c = (Ctx*)ctx;
    if (!c) return NGSEP_EINVAL;
    if (c->ring_count == c->ring.size()) return fail(c, NGSEP_ESTATE, "too many tiles in flight: call ngsep_tile_fetch first");
    Slot* s = c->ring[c->ring_head];
    const double ts0 = host_now();
    int rc = stage_slot(c, s, t, true);
    if (rc != NGSEP_OK) return rc;
    const double ts1 = host_now();
    rc = launch_slot(c, s);
    if (rc != NGSEP_OK) return rc;
    c->tm[0] += ts1 - ts0; c->tm[1] += host_now() - ts1;
    c->ring_head = (c->ring_head + 1) % c->ring.size();
    c->ring_count++;
    return NGSEP_OK;
}

int ngsep_tile_sites(void* ctx, int32_t n_sites, const int32_t* pos, const uint16_t* alleles) {
    Ctx* c = (Ctx*)ctx;
    if (!c) return NGSEP_EINVAL;
    if (c->params.n_samples > 1) return fail(c, NGSEP_ESTATE, "ngsep_tile_sites: single-sample contexts only");
    if (n_sites < 0 || (n_sites > 0 && (!pos || !alleles))) return fail(c, NGSEP_EINVAL, "ngsep_tile_sites: bad arguments");
    for (int32_t k = 0; k < n_sites; k++) {
        const unsigned n_al = (alleles[k] >> 8) & 7u;
        if (n_al < 2 || n_al > 4 || (alleles[k] >> 11)) return fail(c, NGSEP_EINVAL, "ngsep_tile_sites: a site needs 2 to 4 alleles");
    }
    c->next_site_pos.assign(pos, pos + n_sites);
    c->next_site_alleles.assign(alleles, alleles + n_sites);
    c->next_has_sites = true;
    return NGSEP_OK;
}

int ngsep_tile_fetch(void* ctx, ngsep_snv_call* out, size_t cap, size_t* n_out, int32_t* counts_dbg, ngsep_tile_stats* stats) {
    Ctx* c = (Ctx*)ctx;
    if (!c) return NGSEP_EINVAL;
    if (c->params.n_samples > 1) return fail(c, NGSEP_ESTATE, "multi-sample context: use ngsep_tile_fetch_pop");   // the tile stays in flight
    if (c->ring_count == 0) return fail(c, NGSEP_ESTATE, "no tile in flight");
    Slot* s = c->ring[c->ring_tail];
    int rc = fetch_slot(c, s, out, cap, n_out, counts_dbg, stats);
    if (rc == NGSEP_EOVERFLOW && n_out && *n_out > cap) return rc;   // the caller retries with a larger buffer: the tile stays at the tail
    c->ring_tail = (c->ring_tail + 1) % c->ring.size();
    c->ring_count--;
    return rc;
}

int ngsep_samples_set(void* ctx, int32_t n_groups, const int32_t* sample_group_off, const uint16_t* sample_groups, double min_allele_depth_freq) {
    Ctx* c = (Ctx*)ctx;
    if (!c) return NGSEP_EINVAL;
    const int n_samples = c->params.n_samples;
    if (n_samples < 2) return fail(c, NGSEP_ESTATE, "ngsep_samples_set needs a context created with n_samples > 1");
    if (!sample_group_off || !sample_groups || n_groups < 1 || n_groups > 65535) return fail(c, NGSEP_EINVAL, "bad read-group table");
    if (!(min_allele_depth_freq >= 0.0 && min_allele_depth_freq <= 1.0)) return fail(c, NGSEP_EINVAL, "min_allele_depth_freq must be in [0,1]");
    if (c->ring_count != 0) return fail(c, NGSEP_ESTATE, "tiles in flight");
    int max_g = 0;
    if (sample_group_off[0] != 0) return fail(c, NGSEP_EINVAL, "sample_group_off[0] must be 0");
    for (int s = 0; s < n_samples; s++) {
        const int n = sample_group_off[s + 1] - sample_group_off[s];
        if (n < 0) return fail(c, NGSEP_EINVAL, "sample_group_off must be non-decreasing");
        max_g = std::max(max_g, n);
    }
    const int total = sample_group_off[n_samples];
    std::vector<char> seen((size_t)n_groups, 0);
    for (int k = 0; k < total; k++) {
        if (sample_groups[k] >= n_groups) return fail(c, NGSEP_EINVAL, "read-group id out of range");
        if (seen[sample_groups[k]]) return fail(c, NGSEP_EINVAL, "a read group belongs to two samples (MSVD:500-502)");
        seen[sample_groups[k]] = 1;
    }
    CK(cudaSetDevice(c->device), "set device");
    CK(cudaStreamSynchronize(c->stream), "drain");
    {   // one wave of persistent CTAs (the kernel's column loop is grid-strided)
        int per_sm = genotype_multi_ctas_per_sm(n_samples);
        if (per_sm < 1) return fail(c, NGSEP_ECUDA, "the multi-sample kernel cannot be made resident");
        if (const char* e = std::getenv("NGSEP_MULTI_BLOCKS_PER_SM")) { int v = std::atoi(e); if (v >= 1 && v < per_sm) per_sm = v; }
        c->multi_grid = c->sm_count * per_sm;
    }
    CK(c->d_sgo.reserve((size_t)n_samples + 1), "sample table alloc");
    CK(c->d_sgroups.reserve((size_t)std::max(total, 1)), "sample table alloc");
    CK(c->d_scratch.reserve((size_t)c->multi_grid * (size_t)n_samples), "scratch alloc");
    CK(cudaMemcpy(c->d_sgo.p, sample_group_off, ((size_t)n_samples + 1) * sizeof(int), cudaMemcpyHostToDevice), "sample table copy");
    if (total > 0) CK(cudaMemcpy(c->d_sgroups.p, sample_groups, (size_t)total * sizeof(uint16_t), cudaMemcpyHostToDevice), "sample table copy");
    CK(c->d_gins.reserve((size_t)n_groups), "sample table alloc");   // which read groups belong to a sample at all
    CK(cudaMemcpy(c->d_gins.p, seen.data(), (size_t)n_groups, cudaMemcpyHostToDevice), "sample table copy");
    c->hp.n_groups = n_groups;
    c->hp.max_groups_per_sample = max_g;
    c->hp.min_allele_depth_freq = min_allele_depth_freq;
    c->hp.sample_group_off = c->d_sgo.p;
    c->hp.sample_groups = c->d_sgroups.p;
    c->hp.scratch = c->d_scratch.p;
    CK(cudaMemcpy(c->d_prm, &c->hp, sizeof(DevParams), cudaMemcpyHostToDevice), "params copy");
    c->samples_set = true;
    return NGSEP_OK;
}

int ngsep_tile_fetch_pop(void* ctx, ngsep_pop_site* sites, size_t cap_sites, size_t* n_sites, ngsep_sample_call* calls, int32_t* counts_dbg,
                         ngsep_tile_stats* stats) {
    Ctx* c = (Ctx*)ctx;
    if (!c) return NGSEP_EINVAL;
    if (c->params.n_samples < 2) return fail(c, NGSEP_ESTATE, "single-sample context: use ngsep_tile_fetch");
    if (c->ring_count == 0) return fail(c, NGSEP_ESTATE, "no tile in flight");
    Slot* s = c->ring[c->ring_tail];
    int rc = fetch_slot(c, s, nullptr, cap_sites, n_sites, counts_dbg, stats, sites, calls);
    if (rc == NGSEP_EOVERFLOW && n_sites && *n_sites > cap_sites) return rc;   // the tile stays at the tail: retry with room for *n_sites
    c->ring_tail = (c->ring_tail + 1) % c->ring.size();
    c->ring_count--;
    return rc;
}

int ngsep_resident_fetch_pop(void* ctx, int slot, ngsep_pop_site* sites, size_t cap_sites, size_t* n_sites, ngsep_sample_call* calls,
                             int32_t* counts_dbg, ngsep_tile_stats* stats) {
    Ctx* c = (Ctx*)ctx;
    if (!c) return NGSEP_EINVAL;
    if (slot < 1 || (size_t)slot >= c->slots.size()) return fail(c, NGSEP_EINVAL, "bad slot");
    return fetch_slot(c, c->slots[slot], nullptr, cap_sites, n_sites, counts_dbg, stats, sites, calls);
}

int ngsep_resident_add(void* ctx, const ngsep_tile* t, int* slot) {
    Ctx* c = (Ctx*)ctx;
    if (!c || !slot) return NGSEP_EINVAL;
    if (c->next_has_sites) return fail(c, NGSEP_ESTATE, "ngsep_tile_sites declares the sites of a submitted tile: resident tiles have none");
    CK(cudaSetDevice(c->device), "set device");
    Slot* s = nullptr;
    int rc = new_slot(c, &s, c->slots.size());
    if (rc != NGSEP_OK) return rc;
    rc = stage_slot(c, s, t);
    if (rc != NGSEP_OK) { s->release(); delete s; return rc; }
    c->slots.push_back(s);
    c->batch_dirty = true;
    *slot = (int)c->slots.size() - 1;
    return NGSEP_OK;
}

// ---- CoverageStatisticsCalculator: read-depth histograms of one region tile (coverage_kernels.cu) ---------------------------------
int ngsep_coverage_histogram(void* ctx, int32_t tile_first, int32_t tile_last, int64_t n, const int32_t* first, const int32_t* last, const uint8_t* unique,
                             int32_t max_coverage, uint64_t* hist_all, uint64_t* hist_unique) {
    Ctx* c = (Ctx*)ctx;
    if (!c) return NGSEP_EINVAL;
    if (tile_first < 1 || tile_last < tile_first || n < 0 || max_coverage < 1 || max_coverage > 4095 || !hist_all || !hist_unique || (n > 0 && (!first || !last || !unique)))
        return fail(c, NGSEP_EINVAL, "bad coverage arguments");
    CK(cudaSetDevice(c->device), "set device");
    const int64_t n_pos = (int64_t)tile_last - tile_first + 1;
    const size_t bins = (size_t)max_coverage + 1;
    CK(c->d_cov_first.reserve((size_t)n + 1), "coverage alloc"); CK(c->d_cov_last.reserve((size_t)n + 1), "coverage alloc"); CK(c->d_cov_unique.reserve((size_t)n + 1), "coverage alloc");
    CK(c->d_cov_diff.reserve((size_t)n_pos + 2), "coverage alloc");
    const size_t tmp = coverage_scan_temp_bytes(n_pos);
    CK(c->d_cov_tmp.reserve(tmp + 256), "coverage alloc");
    CK(c->d_cov_hist.reserve(2 * bins), "coverage alloc");
    if (n > 0) {
        CK(cudaMemcpyAsync(c->d_cov_first.p, first, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream), "coverage copy");
        CK(cudaMemcpyAsync(c->d_cov_last.p, last, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream), "coverage copy");
        CK(cudaMemcpyAsync(c->d_cov_unique.p, unique, (size_t)n, cudaMemcpyHostToDevice, c->stream), "coverage copy");
    }
    CK(cudaMemsetAsync(c->d_cov_hist.p, 0, 2 * bins * sizeof(unsigned long long), c->stream), "coverage memset");
    CK(launch_coverage(c->d_cov_first.p, c->d_cov_last.p, c->d_cov_unique.p, n, tile_first, tile_last, max_coverage, c->d_cov_diff.p, c->d_cov_tmp.p, c->d_cov_tmp.cap,
                       c->d_cov_hist.p, c->d_cov_hist.p + bins, c->sm_count, c->stream), "coverage kernels");
    std::vector<unsigned long long> h(2 * bins);
    CK(cudaMemcpyAsync(h.data(), c->d_cov_hist.p, 2 * bins * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream), "coverage copy");
    CK(cudaStreamSynchronize(c->stream), "coverage sync");
    for (size_t k = 0; k < bins; k++) { hist_all[k] = h[k]; hist_unique[k] = h[bins + k]; }
    return NGSEP_OK;
}

// ---- indel / STR allele tallies of the span > 1 events (k_indel_tallies) ------------------------------------------------------------
int ngsep_indel_tallies(void* ctx, const ngsep_indel_events* e, int32_t* total, int32_t* counts, int32_t* strand, double* diag, int32_t* call_index, double* call_fit) {
    Ctx* c = (Ctx*)ctx;
    if (!c) return NGSEP_EINVAL;
    if (!e || e->n_events < 0 || e->n_alleles < 0 || e->n_calls < 0) return fail(c, NGSEP_EINVAL, "ngsep_indel_tallies: bad arguments");
    if (e->n_events == 0) return NGSEP_OK;
    if (!total || !counts || !strand || !diag || !e->allele_first || !e->call_first || !e->allele_off || !e->call_off || !e->allele_chars ||
        (e->n_calls > 0 && (!e->call_chars || !e->call_quals || !e->call_neg || !call_index || !call_fit)))
        return fail(c, NGSEP_EINVAL, "ngsep_indel_tallies: null argument");
    // the layout is checked on the host: monotone offsets that cover the arrays, at least the reference allele per event
    if (e->allele_first[0] != 0 || e->call_first[0] != 0 || e->allele_first[e->n_events] != e->n_alleles || e->call_first[e->n_events] != e->n_calls)
        return fail(c, NGSEP_EINVAL, "ngsep_indel_tallies: event offsets do not cover the alleles / calls");
    for (int k = 0; k < e->n_events; k++)
        if (e->allele_first[k + 1] <= e->allele_first[k] || e->call_first[k + 1] < e->call_first[k]) return fail(c, NGSEP_EINVAL, "ngsep_indel_tallies: bad event");
    if (e->allele_off[0] != 0 || e->call_off[0] != 0) return fail(c, NGSEP_EINVAL, "ngsep_indel_tallies: offsets must start at 0");
    for (int a = 0; a < e->n_alleles; a++) if (e->allele_off[a + 1] < e->allele_off[a]) return fail(c, NGSEP_EINVAL, "ngsep_indel_tallies: allele offsets are not monotone");
    for (int k = 0; k < e->n_calls; k++) if (e->call_off[k + 1] < e->call_off[k]) return fail(c, NGSEP_EINVAL, "ngsep_indel_tallies: call offsets are not monotone");
    CK(cudaSetDevice(c->device), "set device");
    const size_t n_ev = (size_t)e->n_events, n_al = (size_t)e->n_alleles, n_ca = (size_t)e->n_calls;
    const size_t n_achars = (size_t)e->allele_off[n_al], n_cchars = (size_t)e->call_off[n_ca];
    auto up = [](size_t v) { return (v + 15) & ~(size_t)15; };
    size_t o = 0;   // one input allocation: offsets and strings
    const size_t o_af = o; o += up((n_ev + 1) * 4);
    const size_t o_cf = o; o += up((n_ev + 1) * 4);
    const size_t o_ao = o; o += up((n_al + 1) * 4);
    const size_t o_co = o; o += up((n_ca + 1) * 4);
    const size_t o_ac = o; o += up(n_achars + 1);
    const size_t o_cc = o; o += up(n_cchars + 1);
    const size_t o_cq = o; o += up(n_cchars + 1);
    const size_t o_cn = o; o += up(n_ca + 1);
    CK(c->d_ev_in.reserve(o), "indel event alloc");
    std::vector<unsigned char> h(o, 0);
    std::memcpy(h.data() + o_af, e->allele_first, (n_ev + 1) * 4);
    std::memcpy(h.data() + o_cf, e->call_first, (n_ev + 1) * 4);
    std::memcpy(h.data() + o_ao, e->allele_off, (n_al + 1) * 4);
    std::memcpy(h.data() + o_co, e->call_off, (n_ca + 1) * 4);
    std::memcpy(h.data() + o_ac, e->allele_chars, n_achars);
    if (n_cchars) { std::memcpy(h.data() + o_cc, e->call_chars, n_cchars); std::memcpy(h.data() + o_cq, e->call_quals, n_cchars); }
    if (n_ca) std::memcpy(h.data() + o_cn, e->call_neg, n_ca);
    CK(cudaMemcpyAsync(c->d_ev_in.p, h.data(), o, cudaMemcpyHostToDevice, c->stream), "indel event copy");
    size_t q = 0;   // one output allocation: totals, counts, strand counts, diagonals, per-call results + the kernel's scratch
    const size_t q_t = q; q += up(n_ev * 4);
    const size_t q_c = q; q += up(n_al * 4);
    const size_t q_s = q; q += up(2 * n_al * 4);
    const size_t q_d = q; q += up(n_al * 8);
    const size_t q_i = q; q += up(n_ca * 4);
    const size_t q_f = q; q += up(n_ca * 8);
    const size_t q_out = q;
    const size_t q_lc = q; q += up(n_al * 8);
    CK(c->d_ev_out.reserve(q), "indel tally alloc");
    IndelEventsDev d;
    unsigned char* base = c->d_ev_in.p;
    d.n_events = e->n_events; d.max_qs = c->params.max_base_qs;
    d.allele_first = (const int*)(base + o_af); d.call_first = (const int*)(base + o_cf); d.allele_off = (const int*)(base + o_ao); d.call_off = (const int*)(base + o_co);
    d.allele_chars = (const char*)(base + o_ac); d.call_chars = (const char*)(base + o_cc); d.call_quals = base + o_cq; d.call_neg = base + o_cn;
    d.log_err = e->log_err;
    unsigned char* ob = c->d_ev_out.p;
    launch_indel_tallies(d, c->d_prm, (int*)(ob + q_t), (int*)(ob + q_c), (int*)(ob + q_s), (double*)(ob + q_d), (int*)(ob + q_i), (double*)(ob + q_f), (double*)(ob + q_lc), c->stream);
    CK(cudaGetLastError(), "indel tally launch");
    std::vector<unsigned char> r(q_out);
    CK(cudaMemcpyAsync(r.data(), ob, q_out, cudaMemcpyDeviceToHost, c->stream), "indel tally copy");
    CK(cudaStreamSynchronize(c->stream), "indel tallies");
    std::memcpy(total, r.data() + q_t, n_ev * 4);
    std::memcpy(counts, r.data() + q_c, n_al * 4);
    std::memcpy(strand, r.data() + q_s, 2 * n_al * 4);
    std::memcpy(diag, r.data() + q_d, n_al * 8);
    if (n_ca) { std::memcpy(call_index, r.data() + q_i, n_ca * 4); std::memcpy(call_fit, r.data() + q_f, n_ca * 8); }
    return NGSEP_OK;
}

// ---- RelativeAlleleCountsCalculator: the allele counts of every position of one region tile (k_allele_counts) --------------------
int ngsep_tile_allele_counts(void* ctx, const ngsep_tile* t, int32_t min_rd, int32_t min_base_qs, uint32_t* out) {
    Ctx* c = (Ctx*)ctx;
    if (!c) return NGSEP_EINVAL;
    if (!t || !out) return fail(c, NGSEP_EINVAL, "null argument");
    if (c->params.n_samples > 1) return fail(c, NGSEP_ESTATE, "ngsep_tile_allele_counts: single-sample contexts only");
    if (c->ring_count != 0) return fail(c, NGSEP_ESTATE, "ngsep_tile_allele_counts: fetch the tiles in flight first");
    // coded qualities keep min(30, q) with every q <= 3 alike (ngsep_tile.qual_bits): enough to compare with a threshold of 4..30 only
    if (t->qual_bits != 0u && t->qual_bits != 8u && (min_base_qs < 4 || min_base_qs > 30))
        return fail(c, NGSEP_EINVAL, "ngsep_tile_allele_counts: this base-quality threshold needs the quality bytes themselves (qual_bits 8)");
    CK(cudaSetDevice(c->device), "set device");
    Slot* s = c->ring[0];
    c->next_has_sites = false;
    int rc = stage_slot(c, s, t);
    if (rc != NGSEP_OK) return rc;
    const size_t n_pos = (size_t)(t->tile_last - t->tile_first + 1);
    CK(c->d_allele_out.reserve(n_pos), "allele count alloc");
    CK(cudaStreamSynchronize(c->stream), "staging");
    rc = read_report(c, s);
    if (rc != NGSEP_OK) return rc;
    launch_allele_counts(s->d_tile, (int)n_pos, min_rd, min_base_qs, c->d_allele_out.p, c->sm_count * 8, c->stream);
    CK(cudaGetLastError(), "allele count launch");
    CK(cudaMemcpyAsync(out, c->d_allele_out.p, n_pos * sizeof(uint2), cudaMemcpyDeviceToHost, c->stream), "allele count copy");
    CK(cudaStreamSynchronize(c->stream), "allele counts");
    s->staged = false;
    return NGSEP_OK;
}

// ---- synthetic inputs generated on the device (synth_kernels.cu) ------------------------------------------------------------
int ngsep_synth_reference(void* ctx, int32_t seq_idx, int64_t length, uint64_t seed, double gc) {
    Ctx* c = (Ctx*)ctx;
    if (!c || seq_idx < 0 || length <= 0 || !(gc >= 0.0 && gc <= 1.0)) return fail(c, NGSEP_EINVAL, "bad reference arguments");
    CK(cudaSetDevice(c->device), "set device");
    if ((size_t)seq_idx >= c->refs.size()) c->refs.resize(seq_idx + 1);
    RefSeq& r = c->refs[seq_idx];
    r.release();
    c->batch_dirty = true;
    const int64_t n32 = (length + 31) / 32;
    const size_t w2 = kRefFrontWords + (size_t)n32 * 2 + kRefTailWords2, w1 = kRefFrontWords + (size_t)n32 + kRefTailWords1;
    char* d_ascii = nullptr;
    CK(cudaMalloc((void**)&d_ascii, (size_t)length), "ref ascii alloc");
    cudaError_t e = cudaMalloc((void**)&r.a_ref2, w2 * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&r.a_refn, w1 * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&r.a_refl, w1 * 4);
    if (e == cudaSuccess) e = cudaMemsetAsync(r.a_ref2, 0, w2 * 4, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(r.a_refn, 0xff, w1 * 4, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(r.a_refl, 0, w1 * 4, c->stream);
    if (e == cudaSuccess) {
        r.ref2 = r.a_ref2 + kRefFrontWords; r.ref_n = r.a_refn + kRefFrontWords; r.ref_l = r.a_refl + kRefFrontWords;
        launch_synth_ref(d_ascii, length, seed + (uint64_t)seq_idx * 0x100000001b3ull, gc, c->stream);
        launch_ref_pack(d_ascii, length, r.ref2, r.ref_n, r.ref_l, c->stream);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_ascii);
    if (e != cudaSuccess) { r.release(); return cuda_fail(c, e, "reference generation"); }
    r.len = length;
    return NGSEP_OK;
}

int ngsep_ref_download(void* ctx, int32_t seq_idx, char* bases, int64_t length) {
    Ctx* c = (Ctx*)ctx;
    if (!c || !bases || seq_idx < 0 || (size_t)seq_idx >= c->refs.size() || !c->refs[seq_idx].ref2 || length != c->refs[seq_idx].len) return fail(c, NGSEP_EINVAL, "bad reference arguments");
    CK(cudaSetDevice(c->device), "set device");
    const RefSeq& r = c->refs[seq_idx];
    char* d_ascii = nullptr;
    CK(cudaMalloc((void**)&d_ascii, (size_t)length), "ref ascii alloc");
    launch_ref_unpack(r.ref2, r.ref_n, r.ref_l, length, d_ascii, c->stream);
    cudaError_t e = cudaMemcpyAsync(bases, d_ascii, (size_t)length, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_ascii);
    if (e != cudaSuccess) return cuda_fail(c, e, "reference download");
    return NGSEP_OK;
}

int ngsep_synth_tile(void* ctx, int32_t seq_idx, int32_t tile_first, int32_t tile_last, const ngsep_synth_params* p, int* slot) {
    Ctx* c = (Ctx*)ctx;
    if (!c || !p || !slot) return NGSEP_EINVAL;
    if (c->params.n_samples > 1) return fail(c, NGSEP_ESTATE, "synthetic tiles are single-sample");
    if (seq_idx < 0 || (size_t)seq_idx >= c->refs.size() || !c->refs[seq_idx].ref2) return fail(c, NGSEP_ESTATE, "reference sequence not uploaded");
    if (p->read_len < 16 || p->read_len > 1000 || !(p->depth > 0.0) || !(p->snv_rate >= 0.0 && p->snv_rate < 1.0) || !(p->error_rate >= 0.0 && p->error_rate < 1.0))
        return fail(c, NGSEP_EINVAL, "bad synthetic-read parameters");
    CK(cudaSetDevice(c->device), "set device");
    SynthSpec sp;
    sp.seed = p->seed + (uint64_t)seq_idx * 0x100000001b3ull;
    sp.step = (double)p->read_len / p->depth;
    sp.read_len = p->read_len;
    sp.snv_thr = (uint32_t)(p->snv_rate * 4294967296.0);
    sp.err_thr = (uint32_t)(p->error_rate * 4294967296.0);
    int64_t g_lo = 0, n = 0;
    synth_tile_shape(sp, c->refs[seq_idx].len, tile_first, tile_last, &g_lo, &n);
    if (n > 0x7ffffff0ll) return fail(c, NGSEP_EINVAL, "tile holds too many reads (split it)");
    const int L = p->read_len, W = 1 + ((L + 15) >> 4) + ((L + 31) >> 5) + 1, Q = (((L * 2 + 7) >> 3) + 3) >> 2;
    ngsep_tile t;
    std::memset(&t, 0, sizeof t);
    t.seq_idx = seq_idx; t.tile_first = tile_first; t.tile_last = tile_last; t.n_reads = (int32_t)n;
    t.max_span = L; t.max_back = 0;
    t.n_words = (uint64_t)n * (uint64_t)W;
    t.n_qual = (uint64_t)n * (uint64_t)Q * 4ull;
    t.qual_bits = 2;
    t.qual_dict[0] = 35; t.qual_dict[1] = 43; t.qual_dict[2] = 53; t.qual_dict[3] = 63;   // q <= 3 (never counted), 10, 20, >= 30
    Slot* s = nullptr;
    int rc = new_slot(c, &s, c->slots.size());
    if (rc != NGSEP_OK) return rc;
    rc = stage_slot(c, s, &t, false, &sp, g_lo);
    if (rc != NGSEP_OK) { s->release(); delete s; return rc; }
    c->slots.push_back(s);
    c->batch_dirty = true;
    *slot = (int)c->slots.size() - 1;
    return NGSEP_OK;
}

int ngsep_resident_download(void* ctx, int slot, ngsep_tile* shape, ngsep_read_hdr* hdr, uint32_t* payload, uint8_t* qual) {
    Ctx* c = (Ctx*)ctx;
    if (!c || !shape) return NGSEP_EINVAL;
    if (slot < 1 || (size_t)slot >= c->slots.size()) return fail(c, NGSEP_EINVAL, "bad slot");
    CK(cudaSetDevice(c->device), "set device");
    Slot* s = c->slots[slot];
    CK(cudaStreamSynchronize(c->copy_stream), "sync");
    CK(cudaStreamSynchronize(c->stream), "sync");
    std::memset(shape, 0, sizeof *shape);
    shape->seq_idx = s->tile.seq_idx; shape->tile_first = s->tile.tile_first; shape->tile_last = s->tile.tile_last; shape->n_reads = s->tile.n_reads;
    shape->max_span = s->tile.max_span; shape->max_back = s->tile.max_back; shape->n_words = s->n_words; shape->n_qual = s->n_qual;
    shape->qual_bits = s->tile.qual_bits;
    std::memcpy(shape->qual_dict, s->tile.qual_dict, 16);
    if (hdr) CK(cudaMemcpy(hdr, s->d_hdr.p, (size_t)s->tile.n_reads * sizeof(ngsep_read_hdr), cudaMemcpyDeviceToHost), "hdr download");
    if (payload) CK(cudaMemcpy(payload, s->d_pay.p, (size_t)s->n_words * 4, cudaMemcpyDeviceToHost), "payload download");
    if (qual && !s->qual_on_host) CK(cudaMemcpy(qual, s->d_qual.p, (size_t)s->n_qual, cudaMemcpyDeviceToHost), "quality download");
    return NGSEP_OK;
}

int ngsep_resident_clear(void* ctx) {
    Ctx* c = (Ctx*)ctx;
    if (!c) return NGSEP_EINVAL;
    CK(cudaSetDevice(c->device), "set device");
    CK(cudaStreamSynchronize(c->stream), "sync");
    while (c->slots.size() > 1) { c->slots.back()->release(); delete c->slots.back(); c->slots.pop_back(); }
    c->batch_dirty = true;
    c->batch_ran = false;
    return NGSEP_OK;
}

int ngsep_resident_run(void* ctx, int slot) {
    Ctx* c = (Ctx*)ctx;
    if (!c) return NGSEP_EINVAL;
    CK(cudaSetDevice(c->device), "set device");
    if (slot >= 0) {
        if ((size_t)slot >= c->slots.size()) return fail(c, NGSEP_EINVAL, "bad slot");
        return launch_slot(c, c->slots[slot]);
    }
    return launch_batch(c);
}

int ngsep_resident_stats(void* ctx, ngsep_tile_stats* stats) {
    Ctx* c = (Ctx*)ctx;
    if (!c || !stats) return NGSEP_EINVAL;
    if (!c->batch_ran) return fail(c, NGSEP_ESTATE, "no batched run to report");
    CK(cudaSetDevice(c->device), "set device");
    const size_t n = c->slots.size();
    for (int attempt = 0; attempt < 4; attempt++) {
        CK(cudaStreamSynchronize(c->stream), "tile kernels");
        int grew = 0;
        int rc = queue_overflowed(c, &grew);
        if (rc != NGSEP_OK) return rc;
        if (!grew) break;
        if (attempt == 3) return fail(c, NGSEP_EOVERFLOW, "column buffer overflow");
        rc = launch_batch(c);
        if (rc != NGSEP_OK) return rc;
    }
    std::vector<unsigned char> meta((1 + n) * kMetaStride);
    CK(cudaMemcpy(meta.data(), c->d_meta, meta.size(), cudaMemcpyDeviceToHost), "counters copy");
    *stats = ngsep_tile_stats{};
    for (size_t i = 1; i < n; i++) {
        const unsigned char* m = meta.data() + (1 + i) * kMetaStride;
        const unsigned long long* ctr = (const unsigned long long*)m;
        stats->covered_positions += ctr[0];
        stats->candidate_positions += ctr[1];
        stats->calls += *(const unsigned int*)(m + 16);
        stats->aligned_bases += c->slots[i]->aligned_bases;
    }
    float ms = 0.f, pms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->b0, c->b1), "event elapsed");
    CK(cudaEventElapsedTime(&pms, c->b0, c->bm), "event elapsed");
    stats->kernel_ms = ms;
    stats->pileup_ms = pms;
    stats->kernel_launches = c->b_windows > 0 ? 2u : 0u;
    stats->open_columns = *(const unsigned int*)(meta.data() + 8);
    return NGSEP_OK;
}

int ngsep_resident_fetch(void* ctx, int slot, ngsep_snv_call* out, size_t cap, size_t* n_out, int32_t* counts_dbg, ngsep_tile_stats* stats) {
    Ctx* c = (Ctx*)ctx;
    if (!c) return NGSEP_EINVAL;
    if (slot < 0 || (size_t)slot >= c->slots.size()) return fail(c, NGSEP_EINVAL, "bad slot");
    if (slot == 0) return ngsep_tile_fetch(ctx, out, cap, n_out, counts_dbg, stats);   // slot 0 = the submit/fetch pipeline
    return fetch_slot(c, c->slots[slot], out, cap, n_out, counts_dbg, stats);
}

int ngsep_timer_start(void* ctx) {
    Ctx* c = (Ctx*)ctx;
    if (!c) return NGSEP_EINVAL;
    CK(cudaSetDevice(c->device), "set device");
    CK(cudaEventRecord(c->t0, c->stream), "timer start");
    return NGSEP_OK;
}

int ngsep_timer_stop(void* ctx, float* ms) {
    Ctx* c = (Ctx*)ctx;
    if (!c || !ms) return NGSEP_EINVAL;
    CK(cudaSetDevice(c->device), "set device");
    // single-tile launches run their genotype kernel on its own stream: the stopwatch covers it as well
    CK(cudaEventRecord(c->tg, c->geno_stream), "timer join");
    CK(cudaStreamWaitEvent(c->stream, c->tg, 0), "timer join");
    CK(cudaEventRecord(c->t1, c->stream), "timer stop");
    CK(cudaEventSynchronize(c->t1), "timer sync");
    CK(cudaEventElapsedTime(ms, c->t0, c->t1), "timer elapsed");
    return NGSEP_OK;
}

int ngsep_host_register(void* ptr, size_t bytes) {
    if (!ptr || !bytes) return NGSEP_EINVAL;
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) { g_global_err = std::string("cudaHostRegister: ") + cudaGetErrorString(e); cudaGetLastError(); return NGSEP_ECUDA; }
    return NGSEP_OK;
}

int ngsep_host_unregister(void* ptr) {
    if (!ptr) return NGSEP_EINVAL;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { g_global_err = std::string("cudaHostUnregister: ") + cudaGetErrorString(e); cudaGetLastError(); return NGSEP_ECUDA; }
    return NGSEP_OK;
}

}  // extern "C"
