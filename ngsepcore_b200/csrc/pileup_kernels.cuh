// Device side of libngsep_pileup.so: sm_100a kernels for the pileup walk + SNVQ genotype posterior.
//
// Design: reads are coordinate-sorted, so a window of T reference positions is covered by a
// CONTIGUOUS range of read records.  One CTA owns one window at a time:
//   load    one cp.async.bulk (TMA 1-D bulk copy, mbarrier completion) for the 16-byte headers and one for the
//           packed payload (CIGAR | 2-bit bases | qualities) of the window's records -> shared memory;
//   screen  thread-per-record, bit-parallel: XOR 16 packed read bases against 16 packed reference bases per
//           instruction; only positions where some read disagrees with the reference can yield a non-hom-ref
//           call (VD:100-222 starts from (ref,ref) and needs >0.01 more posterior to move), so only those become
//           candidates (exact superset, no heuristics);
//   exact   thread-per-candidate: walks the covering records IN PILEUP ORDER and restates
//           PileupRecord.getAlleleCalls(1) / CountsHelper.updateCounts / getPosteriorProbabilities /
//           VariantDiscoverySNVQAlgorithm.discoverSNV, summing the 10 distinct log-likelihoods in fp64 in the
//           reference's own order (bit-identical logConditionalProbs given the same tables).
// HBM traffic = each record once (plus the window halo, served by L2); no atomics on the data path except the
// rare candidate-bit / call-slot claims.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ngsep_pileup.h"

namespace ngsep_b200 {

constexpr int kThreads = 256;          // threads per CTA
constexpr int kMaxChunkReads = 2048;   // records staged per chunk (headers: 32 KB + sentinel)

struct WinRange { int lo, hi; uint32_t off_lo, off_hi; };  // record range [lo,hi) of a window and its payload span

struct TileDev {
    const ngsep_read_hdr* hdr;   // [n_reads + 1], sentinel at n_reads carries off = n_words, first = INT_MAX
    const uint32_t* payload;
    const uint16_t* sample;
    int n_reads, max_span, max_back;
    int tile_first, tile_last, seq_idx;
    const uint32_t* ref2;        // 2 bits/base, base idx i (0-based) at bits 2*(i%16) of word i/16
    const uint32_t* ref_n;       // 1 bit/base: not ACGT (after upper-casing)
    const uint32_t* ref_l;       // 1 bit/base: lower case in the FASTA
    int64_t ref_len;
    int T;                       // window length in positions (power of two, >= 64)
    int win0, n_windows;         // first absolute window index of the tile, number of windows
    const WinRange* win;         // [n_windows]
    ngsep_snv_call* calls; unsigned int* n_calls; unsigned int cap_calls;
    int32_t* counts_dbg;         // [tile_len][5] or null
    unsigned long long* counters;// [0] covered positions, [1] candidates, [2] aligned bases
    unsigned int pay_cap_words;  // payload words that fit one staged chunk
};

struct DevParams {
    int max_base_qs, min_quality;
    unsigned int flags;
    int fast_homref;             // 1: columns without a counted non-reference base are provably hom-ref
    double tables[NGSEP_N_TABLES];
};

size_t pileup_smem_bytes(int T, unsigned int pay_cap_words);
void launch_window_index(const TileDev& d, cudaStream_t s);
void launch_pileup(const TileDev& d, const DevParams* dprm, int grid, size_t smem, cudaStream_t s);
void launch_ref_pack(const char* ascii, int64_t len, uint32_t* ref2, uint32_t* ref_n, uint32_t* ref_l, cudaStream_t s);

}  // namespace ngsep_b200
