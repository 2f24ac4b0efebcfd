// Device side of libngsep_pileup.so: sm_100a kernels for the pileup walk + SNVQ genotype posterior.
//
// Design.  Reads are coordinate-sorted, so a WINDOW of T reference positions is covered by a
// CONTIGUOUS range of read records.  One persistent CTA owns one window at a time; ONE launch walks the windows of
// every staged tile (all sequences of a genome):
//   load    cp.async.bulk (TMA 1-D bulk copies, one mbarrier per stage): the window's 16-byte record headers, its
//           packed payload (CIGAR | 2-bit bases | qualities) and the packed reference window -> shared memory;
//   screen  thread-per-record, bit-parallel: 16 packed read bases XOR 16 packed reference bases per instruction.
//           Only a position where some read base differs from the reference can give a non-hom-ref call
//           (VD:100-222 starts from (ref,ref) and needs > 0.01 more posterior to move), so only those positions become
//           candidates (an exact superset; no heuristic).  The same pass derives the covered-position count from a
//           prefix-max scan over the sorted records (no atomics on the per-read path);
//   exact   warp-per-candidate: the 32 lanes resolve the covering records IN PILEUP ORDER (restating
//           PileupRecord.getAlleleCalls(1) / CountsHelper.updateCounts), ballots give the exact allele / strand counts,
//           a proven bound drops columns whose posterior of (ref,ref) exceeds 1/2, and the survivors get the 10
//           distinct log-likelihoods summed in fp64 in the reference's own order (bit-identical logConditionalProbs
//           given the same tables) followed by CH.getPosteriorProbabilities / VD.discoverSNV, 16 events on 16 lanes.
// HBM traffic = each record once (window halos are L2 hits); the only atomics on the data path are the rare
// candidate-bit and call-slot claims.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ngsep_pileup.h"

namespace ngsep_b200 {

constexpr int kThreads = 256;          // largest CTA (the window kernel is built for 128 and 256 threads)
constexpr int kRefLeft = 64;           // reference bases staged to the left of a window (read words straddle its start)
constexpr int kMaxWindow = 8192;       // largest window length (positions); windows are multiples of 128

// Record range [lo,hi) of one window, its payload span (words) and where the window lives: 0-based index of its first
// position and the index range [g_lo, g_hi) that lies inside the tile and the sequence.  32 bytes: bulk-copied with the window.
struct WinRange {
    int lo, hi; uint32_t off_lo, off_hi; int tile, wbeg;
    uint32_t g_pack;     // (g_lo - wbeg) | (g_hi - wbeg) << 16
    uint32_t span_pack;  // the tile's max_span | max_back << 16 (65535 = larger than that: columns are handed over without a record range)
};

struct TileDev {
    const ngsep_read_hdr* hdr;   // [n_reads + 1], sentinel at n_reads carries off = n_words, first = INT_MAX
    const uint32_t* payload;
    const uint16_t* group;       // [n_reads] read-group id per record (multi-sample contexts), else null
    const uint32_t* gidx;        // multi-sample contexts: record indexes sorted by read group (stable), see launch_group_index
    const uint32_t* goff;        //                        [n_groups + 2] list offsets into gidx; list n_groups = records outside every sample
    const uint32_t* gblk;        //                        [n_groups + 1][n_gblk + 1] list position of the first record index >= b * 256
    int n_gblk;
    // multi-sample contexts: what a (record, column) evaluation needs of a SIMPLE record in two sectors instead of four (k_build_evidence):
    // mhdr = the record headers with `off` replaced by the record's qoff; ev[qoff * ev_mul + k] = the evidence of read base k
    // (0: a call that is not counted; else 0x80 | quality << 2 | base, the quality rules of CH:91, 214-219 already applied)
    const ngsep_read_hdr* mhdr; const uint8_t* ev; uint32_t ev_mul;
    const uint8_t* qual;         // phred+33 bytes (or qual_bits-bit codes into qual_dict), addressed through the records' qoff (only the genotype kernels read them)
    uint32_t qual_bits;          // 8 = bytes, 2 / 4 = codes
    uint32_t qual_dict[4];       // code -> phred+33 byte, four codes per word
    int n_reads, max_span, max_back;
    int tile_first, tile_last, seq_idx;
    const uint32_t* ref2;        // 2 bits/base, base idx i (0-based) at bits 2*(i%16) of word i/16; 64 B of zero padding in front
    const uint32_t* ref_n;       // 1 bit/base: not ACGT (after upper-casing)
    const uint32_t* ref_l;       // 1 bit/base: lower case in the FASTA
    int64_t ref_len;
    int T;                       // window length in positions (multiple of 128, <= the launch's layout window)
    int win0, n_windows;         // first absolute window index of the tile, number of windows
    ngsep_snv_call* calls; unsigned int* n_calls; unsigned int cap_calls;
    ngsep_pop_site* pop_sites; ngsep_sample_call* pop_calls;   // multi-sample contexts: cap_calls sites x n_samples sample calls
    int32_t* counts_dbg;         // [tile_len][5] or null
    unsigned long long* counters;// [0] covered positions, [1] candidates
    const unsigned int* report;  // staging check of the tile (k_check_headers): [0] error bits
    // -knownVariants (ngsep_tile_sites): the SNV sites genotyped on this tile instead of the discovery
    const int32_t* site_pos; const uint16_t* site_alleles; int n_sites;
};

// tab2[(q * 4 + b) * kTabLanes + l]: what a counted base b of quality q adds to the log-likelihood lane l of the genotype kernels owns
// (lanes 0-3 the homozygous L[l][l], lanes 4-9 the pairs (0,1) (0,2) (0,3) (1,2) (1,3) (2,3); CH:231-249 with f == g == 250: the matrix
// is symmetric); row kTabZeroRow is all zeros (a call that is not counted adds nothing).  Filled once per context.
constexpr int kTabLanes = 10, kTabZeroRow = 31 * 4, kTab2Size = (kTabZeroRow + 1) * kTabLanes;

struct DevParams {
    int max_base_qs, min_quality;
    unsigned int flags;
    int fast_homref;             // 1: the proven hom-ref skips below are valid for these tables / priors
    int pool_ploidy;             // >= 3: pools / polyploids (SSVPL:240-254): no hom-ref proof, k_pool_columns instead of k_genotype_columns; else 0
    int k_min[16];               // k_min[m]: with m counted non-reference bases, k >= k_min[m] counted reference bases
                                 // prove P(ref,ref) > 1/2; INT_MAX when no k suffices
    double tables[NGSEP_N_TABLES];
    double tab2[kTab2Size];
    // sample axis (MultisampleVariantsDetector); n_samples == 1: everything below is unused
    int n_samples, n_groups, max_groups_per_sample;
    int column_run;                      // k_genotype_columns_multi: neighbouring list entries one CTA takes in a row
    int multi_ordered;                   // 1: the window kernel lists a window's candidates in position order (sample axis)
    int sample_min_quality;              // minQuality of the fresh per-sample listener: always 40 (MSVD:676-678, SSVPL:50)
    double min_allele_depth_freq;        // MSVD:58
    const int* sample_group_off;         // [n_samples + 1]
    const uint16_t* sample_groups;       // read groups of each sample, in the order PR.getAlleleCalls(span, Set) visits them
    ngsep_sample_call* scratch;          // [multi grid][n_samples] per-CTA work area of k_genotype_columns_multi
};

// Open columns = the positions the window kernel could not close (screen + hom-ref proof), handed to k_genotype_columns:
// (tile, 1-based position, index in the tile of the first record that can cover it, number of such records | 0xffff = search).
struct ColumnQueue {
    uint4* open;
    unsigned int* ctr;                       // [2] open columns listed, [3] entries reserved in `ent` ([0], [1] unused)
    unsigned int cap_open;
    uint2* ent;                              // per-record entries of the open columns the window kernel resolved (see emit_open_column)
    unsigned int cap_ent;
};

struct LaunchCfg {
    int T_lay;                   // window length the shared-memory layout is sized for
    unsigned int pay_cap_words;  // payload words one stage holds
    int n_stages;                // 1 or 2 staged windows per CTA
    int min_blocks;              // register budget: resident CTAs per SM the kernel variant is compiled for (256 threads: 3, 4, 5; 128: 8, 10, 12)
    int threads;                 // CTA size: 128 or 256
    int l2_prefetch;             // 1: the window after the next one is prefetched into L2
};

size_t pileup_smem_bytes(const LaunchCfg& cfg);
void launch_window_index(const TileDev& d, WinRange* out, int tile_idx, cudaStream_t s);
int pileup_max_ctas_per_sm(size_t smem, int threads, int min_blocks);   // also raises the kernel's dynamic shared-memory limit
size_t expand_scan_temp_bytes(int n);
cudaError_t launch_expand_compact(ngsep_read_hdr* hdr, int n, const uint32_t* wire, unsigned long long n_wire, uint32_t* payload, unsigned long long n_words_full,
                                  const uint8_t* qual, unsigned long long n_qual, uint32_t qual_bits, const uint32_t* qual_dict4, uint32_t* sizes, uint32_t* offs,
                                  void* tmp, size_t tmp_bytes, unsigned int* report, cudaStream_t s);
void launch_check_headers(ngsep_read_hdr* hdr, const uint32_t* payload, int n, unsigned long long n_words, unsigned long long n_qual, unsigned int qual_bits,
                          unsigned int pay_cap_words, unsigned int* report, cudaStream_t s);
void launch_pileup(const TileDev* d_tiles, const WinRange* d_win, int n_windows, const DevParams* dprm, const ColumnQueue& cq, const LaunchCfg& cfg,
                   int grid, size_t smem, cudaStream_t s);
void launch_genotype(const TileDev* d_tiles, const ColumnQueue& cq, const DevParams* dprm, int grid, cudaStream_t s);
void launch_pool_sites(const TileDev* d_tile, const DevParams* dprm, int n_sites, int max_grid, cudaStream_t s);
void launch_pool_columns(const TileDev* d_tiles, const ColumnQueue& cq, const DevParams* dprm, int grid, cudaStream_t s);
void launch_genotype_sites(const TileDev* d_tile, const DevParams* dprm, int n_sites, int max_grid, cudaStream_t s);
// indel / STR allele tallies (k_indel_tallies): device view of ngsep_indel_events
struct IndelEventsDev {
    int n_events, max_qs;
    const int* allele_first; const int* call_first; const int* allele_off; const int* call_off;
    const char* allele_chars; const char* call_chars; const unsigned char* call_quals; const unsigned char* call_neg;
    double log_err;
};
void launch_indel_tallies(const IndelEventsDev& ev, const DevParams* dprm, int* total, int* counts, int* strand, double* diag, int* call_index, double* call_fit,
                          double* lc_scratch, cudaStream_t s);
void launch_allele_counts(const TileDev* d_tile, int n_pos, int min_rd, int min_bq, uint2* out, int max_grid, cudaStream_t s);
void launch_genotype_multi(const TileDev* d_tiles, const ColumnQueue& cq, const DevParams* dprm, int n_samples, int grid, cudaStream_t s);
int genotype_multi_ctas_per_sm(int n_samples);
size_t group_index_temp_bytes(int n);
cudaError_t launch_group_index(const uint16_t* group, const uint8_t* in_sample /* [n_groups]: the read group belongs to a sample */, int n, int n_groups, uint16_t* keys_tmp, uint32_t* vals_tmp, void* cub_tmp, size_t cub_bytes,
                               uint32_t* gidx, uint32_t* goff, uint32_t* gblk, cudaStream_t s);
void launch_build_evidence(const TileDev& d, ngsep_read_hdr* mhdr, uint8_t* ev, int max_qs, cudaStream_t s);
int group_blocks(int n);   // blocks of record indexes in the per-group block table
void launch_ref_pack(const char* ascii, int64_t len, uint32_t* ref2, uint32_t* ref_n, uint32_t* ref_l, cudaStream_t s);

// read-depth histograms of a tile (coverage_kernels.cu)
size_t coverage_scan_temp_bytes(int64_t n_pos);
cudaError_t launch_coverage(const int32_t* first, const int32_t* last, const uint8_t* unique, int64_t n, int32_t tile_first, int32_t tile_last, int max_cov, int2* diff,
                            void* tmp, size_t tmp_bytes, unsigned long long* hist_all, unsigned long long* hist_unique, int sm_count, cudaStream_t s);

// synthetic inputs generated on the device (synth_kernels.cu)
struct SynthSpec { uint64_t seed; double step; int read_len; uint32_t snv_thr, err_thr; };   // step = read_len / depth; thresholds = rate * 2^32
void launch_synth_ref(char* ascii, int64_t len, uint64_t seed, double gc, cudaStream_t s);
void launch_ref_unpack(const uint32_t* ref2, const uint32_t* ref_n, const uint32_t* ref_l, int64_t len, char* ascii, cudaStream_t s);
void synth_tile_shape(const SynthSpec& sp, int64_t seq_len, int32_t tile_first, int32_t tile_last, int64_t* g_lo, int64_t* n_reads);
void launch_synth_tile(const SynthSpec& sp, int64_t g_lo, int n_reads, const uint32_t* ref2, ngsep_read_hdr* hdr, uint32_t* payload, uint8_t* qual, cudaStream_t s);

}  // namespace ngsep_b200
