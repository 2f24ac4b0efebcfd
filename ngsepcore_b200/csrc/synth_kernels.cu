// Synthetic inputs generated ON THE DEVICE (BASELINE configs[2] and [4]: a 64 Mb / 3.1 Gb genome at 30-60x is 5-90 GB of
// packed reads that have no reason to cross PCIe for a throughput measurement).  A reference sequence and the packed tile of
// any region — record headers, payload words and 2-bit quality codes in the layout of include/ngsep_pileup.h — are pure
// functions of (seed, sequence, position / read index), so that every tile of a genome, on whatever GPU, sees the same reads,
// and a tile can be downloaded and handed to the oracle as SAM text (tests/test_gpu_synth_parity.py).
// Rates follow the reference's own simulator (simulation/SingleIndividualSimulator.java:57-58: SNV 1e-3 per base, half of
// them heterozygous); reads are one match block each (the indel path is host work and is exercised by the host-simulated
// workloads), both strands, substitution errors, the quality alphabet of ngsepcore_b200/simulate.py.
#include "pileup_kernels.cuh"

namespace ngsep_b200 {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {   // splitmix64 finaliser
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
constexpr uint64_t kSaltRef = 0x52454600ull, kSaltSnv = 0x534e5600ull, kSaltHap = 0x48415000ull, kSaltErr = 0x45525200ull, kSaltPos = 0x504f5300ull;

// ---- reference: ASCII bases, GC content gc ----------------------------------------------------------------------------------
__global__ void k_synth_ref(char* ascii, int64_t len, uint64_t seed, uint32_t gc_thr /* P(G or C) * 2^32 */) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    const uint64_t h = mix64(seed ^ kSaltRef ^ ((uint64_t)i * 0x9e3779b97f4a7c15ull));
    const bool gc = (uint32_t)h < gc_thr;
    const bool hi = (h >> 40) & 1u;
    ascii[i] = gc ? (hi ? 'G' : 'C') : (hi ? 'T' : 'A');
}
void launch_synth_ref(char* ascii, int64_t len, uint64_t seed, double gc, cudaStream_t s) {
    const double t = gc * 4294967296.0;
    const uint32_t thr = t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
    k_synth_ref<<<(unsigned)((len + 255) / 256), 256, 0, s>>>(ascii, len, seed, thr);
}

// packed reference -> ASCII (for parity checks: the oracle reads a FASTA)
__global__ void k_ref_unpack(const uint32_t* __restrict__ ref2, const uint32_t* __restrict__ ref_n, const uint32_t* __restrict__ ref_l, int64_t len, char* ascii) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    char c = "ACGT"[(ref2[i >> 4] >> ((i & 15) * 2)) & 3u];
    if ((ref_n[i >> 5] >> (i & 31)) & 1u) c = 'N';
    if ((ref_l[i >> 5] >> (i & 31)) & 1u) c = (char)(c + 32);
    ascii[i] = c;
}
void launch_ref_unpack(const uint32_t* ref2, const uint32_t* ref_n, const uint32_t* ref_l, int64_t len, char* ascii, cudaStream_t s) {
    k_ref_unpack<<<(unsigned)((len + 255) / 256), 256, 0, s>>>(ref2, ref_n, ref_l, len, ascii);
}

// ---- reads --------------------------------------------------------------------------------------------------------------------
// Read g of a sequence (g = 0, 1, ...) starts at 1 + floor((g + u_g) * step), u_g in [0, 1), step = read_len / depth: starts
// are non-decreasing in g, so a tile is a contiguous range of g and needs no sort.
__host__ __device__ __forceinline__ int64_t synth_start(uint64_t seed, int64_t g, double step) {
    const double u = (double)(mix64(seed ^ kSaltPos ^ (uint64_t)g) >> 11) * (1.0 / 9007199254740992.0);
    return 1 + (int64_t)(((double)g + u) * step);
}

void synth_tile_shape(const SynthSpec& sp, int64_t seq_len, int32_t tile_first, int32_t tile_last, int64_t* g_lo, int64_t* n_reads) {
    const double step = sp.step;
    const int64_t G = (int64_t)((double)(seq_len - sp.read_len) / step);      // reads g < G end inside the sequence
    int64_t lo = (int64_t)((double)(tile_first - sp.read_len - 1) / step) - 1;  // generous: a few reads that do not reach the tile are harmless
    int64_t hi = (int64_t)((double)tile_last / step) + 2;
    if (lo < 0) lo = 0;
    if (hi > G) hi = G;
    if (hi < lo) hi = lo;
    *g_lo = lo;
    *n_reads = hi - lo;
}

__global__ void k_synth_tile(SynthSpec sp, int64_t g_lo, int n_reads, const uint32_t* __restrict__ ref2, ngsep_read_hdr* hdr, uint32_t* payload, uint8_t* qual) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_reads) return;
    const int L = sp.read_len, nb = (L + 15) >> 4, nw = (L + 31) >> 5;
    const int W = 1 + nb + nw + 1, Q = (((L * 2 + 7) >> 3) + 3) >> 2;   // payload / quality words per read
    const int64_t g = g_lo + i;
    const int64_t start = synth_start(sp.seed, g, sp.step);
    const uint64_t hr = mix64(sp.seed ^ kSaltHap ^ (uint64_t)g);
    const uint32_t hap = (uint32_t)hr & 1u, neg = (uint32_t)(hr >> 1) & 1u;
    ngsep_read_hdr h;
    h.first = (int32_t)start;
    h.meta = (uint32_t)L | (1u << 16) | (neg ? NGSEP_M_NEG : 0u) | NGSEP_M_SIMPLE;
    h.ign_start = 0; h.ign_end = 0;
    h.off = (uint32_t)i * (uint32_t)W;
    hdr[i] = h;
    uint32_t* p = payload + (size_t)i * (size_t)W;
    uint32_t* q = (uint32_t*)qual + (size_t)i * (size_t)Q;
    p[0] = ((uint32_t)L << 3) | 3u;
    uint32_t bw = 0u, lw = 0u, qw = 0u;
    for (int k = 0; k < L; k++) {
        const int64_t idx = start - 1 + k;                                   // 0-based reference index
        uint32_t b = (ref2[idx >> 4] >> ((idx & 15) * 2)) & 3u;
        const uint64_t hs = mix64(sp.seed ^ kSaltSnv ^ ((uint64_t)idx * 0x9e3779b97f4a7c15ull));
        if ((uint32_t)hs < sp.snv_thr) {                                     // a planted SNV site: homozygous, or carried by one haplotype
            const bool hom = (hs >> 40) & 1u;
            if (hom || (((hs >> 41) & 1u) == hap)) b = (b + 1u + (uint32_t)((hs >> 42) % 3u)) & 3u;
        }
        const uint64_t he = mix64(sp.seed ^ kSaltErr ^ ((uint64_t)g * 65536ull + (uint64_t)k));
        if ((uint32_t)he < sp.err_thr) b = (b + 1u + (uint32_t)((he >> 32) % 3u)) & 3u;   // substitution error
        // qualities {40: 70 %, 30: 20 %, 20: 7 %, 10: 2 %, 2: 1 %} = classes q >= 30, 20, 10, low: codes 3, 2, 1, 0
        const uint32_t r = (uint32_t)(he >> 40) & 0xffffu;
        const uint32_t code = r < 655u ? 0u : (r < 1966u ? 1u : (r < 6554u ? 2u : 3u));
        bw |= b << ((k & 15) * 2);
        lw |= (code == 0u ? 1u : 0u) << (k & 31);
        qw |= code << ((k & 15) * 2);
        if ((k & 15) == 15 || k == L - 1) { p[1 + (k >> 4)] = bw; q[k >> 4] = qw; bw = 0u; qw = 0u; }
        if ((k & 31) == 31 || k == L - 1) { p[1 + nb + (k >> 5)] = lw; lw = 0u; }
    }
    for (int k = ((L + 15) >> 4); k < Q; k++) q[k] = 0u;
    p[1 + nb + nw] = (uint32_t)i * (uint32_t)Q;                               // qoff
}

void launch_synth_tile(const SynthSpec& sp, int64_t g_lo, int n_reads, const uint32_t* ref2, ngsep_read_hdr* hdr, uint32_t* payload, uint8_t* qual, cudaStream_t s) {
    if (n_reads > 0) k_synth_tile<<<(n_reads + 127) / 128, 128, 0, s>>>(sp, g_lo, n_reads, ref2, hdr, payload, qual);
}

}  // namespace ngsep_b200
