// C ABI of libngsep_pileup.so (declared in include/ngsep_pileup.h): context, reference upload, tile submit/fetch.
// Host-side plumbing only — all arithmetic of the path lives in pileup_kernels.cu.  There is no CPU fallback:
// every entry point fails with NGSEP_ENODEVICE / NGSEP_ECUDA when the device cannot do the work.
#include <cuda_runtime.h>

#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pileup_kernels.cuh"

using namespace ngsep_b200;

namespace {

struct RefSeq { uint32_t* ref2 = nullptr; uint32_t* ref_n = nullptr; uint32_t* ref_l = nullptr; int64_t len = 0; };

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

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    ngsep_params params{};
    DevParams* d_prm = nullptr;
    std::vector<RefSeq> refs;
    DevBuf<ngsep_read_hdr> d_hdr;
    DevBuf<uint32_t> d_pay;
    DevBuf<uint16_t> d_sample;
    DevBuf<WinRange> d_win;
    DevBuf<ngsep_snv_call> d_calls;
    DevBuf<int32_t> d_counts;
    unsigned int* d_ncalls = nullptr;         // [1]
    unsigned long long* d_counters = nullptr; // [4]
    TileDev tile{};
    bool staged = false, launched = false;
    int grid = 0; size_t smem = 0;
    uint64_t aligned_bases = 0;
    unsigned int launches = 0;
    int sm_count = 0;
    int window = 1024; unsigned int pay_cap_words = 14336;
    std::string err;
};

int fail(Ctx* c, int code, const std::string& msg) { if (c) c->err = msg; return code; }
int cuda_fail(Ctx* c, cudaError_t e, const char* what) {
    return fail(c, NGSEP_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call, what) do { cudaError_t _e = (call); if (_e != cudaSuccess) return cuda_fail(c, _e, what); } while (0)

std::string g_global_err;

int launch_staged(Ctx* c) {
    if (!c->staged) return fail(c, NGSEP_ESTATE, "no tile staged");
    CK(cudaMemsetAsync(c->d_ncalls, 0, sizeof(unsigned int), c->stream), "memset n_calls");
    CK(cudaMemsetAsync(c->d_counters, 0, 4 * sizeof(unsigned long long), c->stream), "memset counters");
    c->launches = 0;
    if (c->tile.counts_dbg) {
        size_t n = (size_t)(c->tile.tile_last - c->tile.tile_first + 1) * 5;
        CK(cudaMemsetAsync(c->tile.counts_dbg, 0, n * sizeof(int32_t), c->stream), "memset counts");
    }
    CK(cudaEventRecord(c->ev0, c->stream), "event record");
    if (c->tile.n_windows > 0) {
        launch_window_index(c->tile, c->stream);
        launch_pileup(c->tile, c->d_prm, c->grid, c->smem, c->stream);
        c->launches = 2;
    }
    CK(cudaEventRecord(c->ev1, c->stream), "event record");
    CK(cudaGetLastError(), "kernel launch");
    c->launched = true;
    return NGSEP_OK;
}

}  // namespace

extern "C" {

int ngsep_abi_version(void) { return 1000; }
int ngsep_built_sm(void) { return 100; }

const char* ngsep_last_error(void* ctx) {
    if (!ctx) return g_global_err.c_str();
    return ((Ctx*)ctx)->err.c_str();
}

int ngsep_ctx_create(int device, const ngsep_params* params, const double* tables, size_t n_tables, void** out) {
    if (!out || !params || !tables) { g_global_err = "null argument"; return NGSEP_EINVAL; }
    *out = nullptr;
    if (n_tables != NGSEP_N_TABLES) { g_global_err = "tables must hold NGSEP_N_TABLES doubles"; return NGSEP_EINVAL; }
    if (params->ploidy >= 3) { g_global_err = "ploidy >= 3 (pool algorithm, SSVPL:240-254) is not on the device path"; return NGSEP_EINVAL; }
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
    if (const char* s = std::getenv("NGSEP_WINDOW")) { int v = std::atoi(s); if (v >= 64 && (v & (v - 1)) == 0 && v <= 8192) c->window = v; }
    if (const char* s = std::getenv("NGSEP_PAY_CAP_WORDS")) { int v = std::atoi(s); if (v >= 1024 && v <= 50000) c->pay_cap_words = (unsigned)v; }
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_prm, sizeof(DevParams));
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_ncalls, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_counters, 4 * sizeof(unsigned long long));
    if (e != cudaSuccess) { g_global_err = std::string("context setup: ") + cudaGetErrorString(e); delete c; return NGSEP_ECUDA; }
    DevParams hp{};
    hp.max_base_qs = params->max_base_qs;
    hp.min_quality = params->min_quality;
    hp.flags = params->flags;
    std::memcpy(hp.tables, tables, sizeof(double) * NGSEP_N_TABLES);
    // Columns without any counted non-reference base are homozygous reference whenever the (ref,ref) event
    // dominates: a_q > c_q > b_q for every q in 4..30 and log prior(hom) - log prior(het) comfortably positive.
    hp.fast_homref = (tables[93] - tables[94] > 1.0) ? 1 : 0;
    for (int q = 4; q <= 30 && hp.fast_homref; q++) if (!(tables[3 * q] > tables[3 * q + 2] && tables[3 * q + 2] > tables[3 * q + 1])) hp.fast_homref = 0;
    e = cudaMemcpy(c->d_prm, &hp, sizeof(hp), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { g_global_err = std::string("params upload: ") + cudaGetErrorString(e); delete c; return NGSEP_ECUDA; }
    *out = c;
    return NGSEP_OK;
}

int ngsep_ctx_destroy(void* ctx) {
    Ctx* c = (Ctx*)ctx;
    if (!c) return NGSEP_EINVAL;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto& r : c->refs) { if (r.ref2) cudaFree(r.ref2); if (r.ref_n) cudaFree(r.ref_n); if (r.ref_l) cudaFree(r.ref_l); }
    c->d_hdr.release(); c->d_pay.release(); c->d_sample.release(); c->d_win.release(); c->d_calls.release(); c->d_counts.release();
    if (c->d_prm) cudaFree(c->d_prm);
    if (c->d_ncalls) cudaFree(c->d_ncalls);
    if (c->d_counters) cudaFree(c->d_counters);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
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
    if (r.ref2) { cudaFree(r.ref2); cudaFree(r.ref_n); cudaFree(r.ref_l); r = RefSeq(); }
    int64_t n32 = (length + 31) / 32;
    char* d_ascii = nullptr;
    CK(cudaMalloc((void**)&d_ascii, (size_t)length), "ref ascii alloc");
    cudaError_t e = cudaMalloc((void**)&r.ref2, (size_t)(n32 * 2 + 4) * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&r.ref_n, (size_t)(n32 + 4) * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&r.ref_l, (size_t)(n32 + 4) * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_ascii, bases, (size_t)length, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(r.ref2, 0, (size_t)(n32 * 2 + 4) * 4, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(r.ref_n, 0xff, (size_t)(n32 + 4) * 4, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(r.ref_l, 0, (size_t)(n32 + 4) * 4, c->stream);
    if (e == cudaSuccess) { launch_ref_pack(d_ascii, length, r.ref2, r.ref_n, r.ref_l, c->stream); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_ascii);
    if (e != cudaSuccess) return cuda_fail(c, e, "reference upload");
    r.len = length;
    return NGSEP_OK;
}

int ngsep_tile_stage(void* ctx, const ngsep_tile* t) {
    Ctx* c = (Ctx*)ctx;
    if (!c || !t) return fail(c, NGSEP_EINVAL, "null argument");
    if (t->n_reads < 0 || t->tile_first < 1 || t->tile_last < t->tile_first) return fail(c, NGSEP_EINVAL, "bad tile extents");
    if (t->seq_idx < 0 || (size_t)t->seq_idx >= c->refs.size() || !c->refs[t->seq_idx].ref2) return fail(c, NGSEP_ESTATE, "reference sequence not uploaded");
    if (t->n_reads > 0 && (!t->hdr || !t->payload)) return fail(c, NGSEP_EINVAL, "null tile buffers");
    if (c->params.n_samples > 1) return fail(c, NGSEP_EINVAL, "multi-sample tiles are not supported by this build");
    CK(cudaSetDevice(c->device), "set device");
    const RefSeq& ref = c->refs[t->seq_idx];
    // validate ordering + chunk fit (a record must fit one staged chunk)
    uint64_t bases = 0;
    for (int i = 0; i < t->n_reads; i++) {
        const ngsep_read_hdr& h = t->hdr[i];
        if (i > 0 && h.first < t->hdr[i - 1].first) return fail(c, NGSEP_EINVAL, "records are not sorted by start");
        uint64_t next = (i + 1 < t->n_reads) ? t->hdr[i + 1].off : t->n_words;
        if (next < h.off || next > t->n_words) return fail(c, NGSEP_EINVAL, "payload offsets are not monotone");
        if (next - h.off + 4 > c->pay_cap_words) return fail(c, NGSEP_EINVAL, "record payload exceeds the device window (read too long)");
        bases += h.meta & 0xffffu;
    }
    c->aligned_bases = bases;
    if (t->n_words > 0xfffffff0ull) return fail(c, NGSEP_EINVAL, "tile payload exceeds 16 GiB");
    // window size: shrink while an average window would not fit one chunk
    int T = c->window;
    int64_t tile_len = (int64_t)t->tile_last - t->tile_first + 1;
    double words_per_pos = (double)t->n_words / (double)tile_len;
    while (T > 64 && words_per_pos * (T + t->max_span) > 0.8 * c->pay_cap_words) T >>= 1;
    TileDev& d = c->tile;
    d = TileDev{};
    CK(c->d_hdr.reserve((size_t)t->n_reads + 1), "hdr alloc");
    CK(c->d_pay.reserve((size_t)t->n_words + 16), "payload alloc");
    if (t->n_reads > 0) {
        CK(cudaMemcpyAsync(c->d_hdr.p, t->hdr, (size_t)t->n_reads * sizeof(ngsep_read_hdr), cudaMemcpyHostToDevice, c->stream), "hdr copy");
        CK(cudaMemcpyAsync(c->d_pay.p, t->payload, (size_t)t->n_words * 4, cudaMemcpyHostToDevice, c->stream), "payload copy");
    }
    ngsep_read_hdr sentinel{INT_MAX, 0u, 0, 0, (uint32_t)t->n_words};
    CK(cudaMemcpyAsync(c->d_hdr.p + t->n_reads, &sentinel, sizeof(sentinel), cudaMemcpyHostToDevice, c->stream), "sentinel copy");
    CK(cudaStreamSynchronize(c->stream), "stage sync");  // sentinel lives on this stack frame
    d.hdr = c->d_hdr.p; d.payload = c->d_pay.p; d.sample = nullptr;
    d.n_reads = t->n_reads; d.max_span = t->max_span < 1 ? 1 : t->max_span; d.max_back = t->max_back < 0 ? 0 : t->max_back;
    d.tile_first = t->tile_first; d.tile_last = t->tile_last; d.seq_idx = t->seq_idx;
    d.ref2 = ref.ref2; d.ref_n = ref.ref_n; d.ref_l = ref.ref_l; d.ref_len = ref.len;
    d.T = T;
    d.win0 = (t->tile_first - 1) / T;
    d.n_windows = (t->n_reads == 0) ? 0 : ((t->tile_last - 1) / T - d.win0 + 1);
    CK(c->d_win.reserve((size_t)d.n_windows + 1), "window alloc");
    d.win = c->d_win.p;
    size_t cap_calls = (size_t)(tile_len / 16) + 65536;
    if (c->d_calls.cap < cap_calls) CK(c->d_calls.reserve(cap_calls), "calls alloc");
    d.calls = c->d_calls.p; d.n_calls = c->d_ncalls; d.cap_calls = (unsigned)c->d_calls.cap;
    d.counts_dbg = nullptr;
    if (c->params.flags & NGSEP_F_EMIT_ALL) {
        CK(c->d_counts.reserve((size_t)tile_len * 5), "counts alloc");
        d.counts_dbg = c->d_counts.p;
    }
    d.counters = c->d_counters;
    d.pay_cap_words = c->pay_cap_words;
    c->smem = pileup_smem_bytes(T, c->pay_cap_words);
    c->grid = d.n_windows;
    c->staged = true;
    c->launched = false;
    return NGSEP_OK;
}

int ngsep_tile_run_staged(void* ctx) {
    Ctx* c = (Ctx*)ctx;
    if (!c) return NGSEP_EINVAL;
    CK(cudaSetDevice(c->device), "set device");
    return launch_staged(c);
}

int ngsep_tile_submit(void* ctx, const ngsep_tile* t) {
    int rc = ngsep_tile_stage(ctx, t);
    if (rc != NGSEP_OK) return rc;
    return launch_staged((Ctx*)ctx);
}

int ngsep_tile_fetch(void* ctx, ngsep_snv_call* out, size_t cap, size_t* n_out, int32_t* counts_dbg, ngsep_tile_stats* stats) {
    Ctx* c = (Ctx*)ctx;
    if (!c || !n_out) return fail(c, NGSEP_EINVAL, "null argument");
    if (!c->launched) return fail(c, NGSEP_ESTATE, "no tile in flight");
    CK(cudaSetDevice(c->device), "set device");
    unsigned int n = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        CK(cudaStreamSynchronize(c->stream), "tile kernels");
        CK(cudaMemcpy(&n, c->d_ncalls, sizeof(n), cudaMemcpyDeviceToHost), "n_calls copy");
        if (n <= c->tile.cap_calls) break;
        if (attempt == 1) return fail(c, NGSEP_EOVERFLOW, "call buffer overflow");
        CK(c->d_calls.reserve((size_t)n + 1024), "calls realloc");   // grow and rerun the staged tile
        c->tile.calls = c->d_calls.p; c->tile.cap_calls = (unsigned)c->d_calls.cap;
        int rc = launch_staged(c);
        if (rc != NGSEP_OK) return rc;
    }
    *n_out = n;
    if (n > cap) return fail(c, NGSEP_EOVERFLOW, "caller's call buffer is too small; n_out holds the required capacity");
    if (n > 0) {
        if (!out) return fail(c, NGSEP_EINVAL, "null output buffer");
        CK(cudaMemcpy(out, c->d_calls.p, (size_t)n * sizeof(ngsep_snv_call), cudaMemcpyDeviceToHost), "calls copy");
    }
    if (counts_dbg && c->tile.counts_dbg) {
        size_t m = (size_t)(c->tile.tile_last - c->tile.tile_first + 1) * 5;
        CK(cudaMemcpy(counts_dbg, c->tile.counts_dbg, m * sizeof(int32_t), cudaMemcpyDeviceToHost), "counts copy");
    }
    if (stats) {
        unsigned long long ctr[4] = {0, 0, 0, 0};
        CK(cudaMemcpy(ctr, c->d_counters, sizeof(ctr), cudaMemcpyDeviceToHost), "counters copy");
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1), "event elapsed");
        stats->covered_positions = ctr[0];
        stats->candidate_positions = ctr[1];
        stats->calls = n;
        stats->aligned_bases = c->aligned_bases;
        stats->kernel_ms = ms;
        stats->kernel_launches = c->launches;
    }
    return NGSEP_OK;
}

}  // extern "C"
