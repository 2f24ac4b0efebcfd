// CoverageStatisticsCalculator on the device (a PileupListener reducer of SURVEY.md 8f-4): the read-depth histograms of a region
// tile.  Reference: discovery/CoverageStatisticsCalculator.java:186-199 (processPileup: PileupRecord.getNumAlignments /
// getNumUniqueAlns per position, PR:154-167) and :250-256 (the printed table).  A position's depth is the number of admitted
// reads with first <= p <= last, so the walk reduces to: +1 / -1 at the ends of every read (atomics into an L2-resident
// difference array), one prefix sum, one histogram.
#include <cub/device/device_scan.cuh>

#include "pileup_kernels.cuh"

namespace ngsep_b200 {

// diff[2 * i] / diff[2 * i + 1]: all / unique alignments at tile position i (interleaved: one scan serves both)
__global__ void k_cov_deltas(const int32_t* __restrict__ first, const int32_t* __restrict__ last, const uint8_t* __restrict__ unique, int64_t n, int32_t tile_first,
                             int32_t tile_last, int2* diff) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int a = max(first[i], tile_first), b = min(last[i], tile_last);
    if (a > b) return;
    const int u = unique[i] ? 1 : 0;
    int* lo = (int*)(diff + (a - tile_first));
    int* hi = (int*)(diff + (b - tile_first + 1));
    atomicAdd(lo, 1); atomicAdd(hi, -1);
    if (u) { atomicAdd(lo + 1, 1); atomicAdd(hi + 1, -1); }
}

struct Int2Sum { __host__ __device__ __forceinline__ int2 operator()(const int2& a, const int2& b) const { return make_int2(a.x + b.x, a.y + b.y); } };

__global__ void k_cov_hist(const int2* __restrict__ depth, int64_t n_pos, int max_cov, unsigned long long* hist_all, unsigned long long* hist_unique) {
    extern __shared__ unsigned int s_h[];   // [2][max_cov + 1]
    for (int k = threadIdx.x; k < 2 * (max_cov + 1); k += blockDim.x) s_h[k] = 0u;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pos; i += (int64_t)gridDim.x * blockDim.x) {
        const int2 d = depth[i];
        atomicAdd(&s_h[min(d.x, max_cov)], 1u);
        atomicAdd(&s_h[max_cov + 1 + min(d.y, max_cov)], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k <= max_cov; k += blockDim.x) {
        if (s_h[k]) atomicAdd(&hist_all[k], (unsigned long long)s_h[k]);
        if (s_h[max_cov + 1 + k]) atomicAdd(&hist_unique[k], (unsigned long long)s_h[max_cov + 1 + k]);
    }
}

size_t coverage_scan_temp_bytes(int64_t n_pos) {
    size_t bytes = 0;
    cub::DeviceScan::InclusiveScan(nullptr, bytes, (const int2*)nullptr, (int2*)nullptr, Int2Sum(), (int)n_pos);
    return bytes;
}

cudaError_t launch_coverage(const int32_t* first, const int32_t* last, const uint8_t* unique, int64_t n, int32_t tile_first, int32_t tile_last, int max_cov, int2* diff,
                            void* tmp, size_t tmp_bytes, unsigned long long* hist_all, unsigned long long* hist_unique, int sm_count, cudaStream_t s) {
    const int64_t n_pos = (int64_t)tile_last - tile_first + 1;
    cudaError_t e = cudaMemsetAsync(diff, 0, (size_t)(n_pos + 1) * sizeof(int2), s);
    if (e != cudaSuccess) return e;
    if (n > 0) k_cov_deltas<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(first, last, unique, n, tile_first, tile_last, diff);
    e = cub::DeviceScan::InclusiveScan(tmp, tmp_bytes, (const int2*)diff, diff, Int2Sum(), (int)n_pos, s);
    if (e != cudaSuccess) return e;
    k_cov_hist<<<sm_count * 4, 256, 2 * (size_t)(max_cov + 1) * sizeof(unsigned int), s>>>(diff, n_pos, max_cov, hist_all, hist_unique);
    return cudaGetLastError();
}

}  // namespace ngsep_b200
