/*
 * ngsep_pileup.h — C ABI of the B200-native pileup / SNVQ hot path (libngsep_pileup.so).
 *
 * This is the drop-in boundary for ONE path of NGSEP 4.3.2 (acastem15/NGSEPcore): the alignment->pileup walk,
 * per-position allele / base-quality counting and the SNVQ genotype posterior.  The reference has no FFI of its
 * own (it is 100 % Java); each entry point below names the Java interface it stands in for (paths relative to
 * /root/reference/src/ngsep/).  The reference-side JNI / ctypes binding is shown in INTEGRATION.md.
 *
 * Conventions: every function returns 0 on success and a negative NGSEP_E* code on failure; ngsep_last_error()
 * gives a context-owned UTF-8 message.  No exception, exit() or CPU fallback crosses this boundary: without a
 * usable CUDA device ngsep_ctx_create fails with NGSEP_ENODEVICE.  A context is bound to one GPU and one host
 * thread.  The caller owns every host buffer it passes; the library owns all device memory.
 */
#ifndef NGSEP_PILEUP_H
#define NGSEP_PILEUP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGSEP_OK 0
#define NGSEP_EINVAL (-1)
#define NGSEP_ENODEVICE (-2)
#define NGSEP_ECUDA (-3)
#define NGSEP_ENOMEM (-4)
#define NGSEP_ESTATE (-5)
#define NGSEP_EOVERFLOW (-6)

/* ngsep_params.flags */
#define NGSEP_F_STRAND_BIAS 0x1u      /* -csb            discovery/SingleSampleVariantsDetector.java:480 (calcStrandBias)   */
#define NGSEP_F_IGNORE_LOWER_REF 0x2u /* -ignoreLowerCaseRef  discovery/SingleSampleVariantPileupListener.java:198          */
#define NGSEP_F_QUALITY_IN_PLACE 0x10u /* single-sample tiles whose quality buffer is page-locked are not copied: the genotype kernel
                                          reads the few quality bytes it needs in place over PCIe (see ngsep_host_register).  Pays on a
                                          host that feeds ONE GPU; with several GPUs reading host memory at once the bulk copy is faster */
#define NGSEP_F_EMIT_ALL 0x8u         /* test hook: also return A,C,G,T,DP for every tile position (counts_dbg)             */

/* Options of main/CommandsDescriptor.xml:565-703 that change what the device computes. */
typedef struct ngsep_params {
    int32_t max_base_qs;  /* -maxBaseQS, default 30      discovery/CountsHelper.java:47,217-219                  */
    int32_t min_quality;  /* -minQuality, default 40     discovery/SingleSampleVariantPileupListener.java:50,223 */
    double het_rate;      /* -h, 1e-3 diploid / 1e-6 haploid   discovery/CountsHelper.java:42-43                 */
    int32_t ploidy;       /* -ploidy, default 2; >= 3: the pool / polyploid algorithm (SSVPL:240-254): single-sample contexts return
                             NGSEP_CALL_POOL columns for the host's genotypeVariantPool                                  */
    uint32_t flags;       /* NGSEP_F_*                                                                          */
    int32_t n_samples;    /* 1 = SingleSampleVariantsDetector; >1 = MultisampleVariantsDetector sample axis      */
    int32_t reserved;
} ngsep_params;

/*
 * Log-probability tables, computed by the HOST (the JVM's Math.log10 in the reference deployment) so that libm
 * differences cannot leak into the device arithmetic.  Layout: tables[3*q + k], q = 0..30:
 *   k=0  a_q = logProbCacheGT[250][q][0]   log10(1-e_q)                      CountsHelper.java:175
 *   k=1  b_q = logProbCacheError[q][4]     -0.1q - log10(3)                  CountsHelper.java:149-153
 *   k=2  c_q = logProbCacheGT[250][q][4]   log10(.5(1-e_q) + .5 e_q/3)       CountsHelper.java:177-178
 * followed by tables[93] = log10((1-h)/4) and tables[94] = log10(h/12)       CountsHelper.java:414-415
 */
#define NGSEP_N_TABLES 95

/* One read record (or one VERSION of a read, see valid window below), 16 bytes, admission order
 * (discovery/AlignmentsPileupGenerator.java:407-433: start ascending, primaries before secondaries, file order). */
typedef struct ngsep_read_hdr {
    int32_t first;       /* 1-based start used for ordering (ReadAlignment.first at admission); non-decreasing */
    uint32_t meta;       /* bits 0-15 read length (incl. soft clips), 16-23 number of CIGAR words, NGSEP_M_* flags */
    int16_t ign_start;   /* ReadAlignment.basesToIgnoreStart after the strand mapping of RA:613-644             */
    int16_t ign_end;     /* ReadAlignment.basesToIgnoreEnd                                                      */
    uint32_t off;        /* offset of this record's payload in 32-bit words                                     */
} ngsep_read_hdr;

#define NGSEP_M_NEG 0x01000000u    /* negative strand (SAM 0x10)                                               */
#define NGSEP_M_NMASK 0x02000000u  /* payload carries a 1-bit/base "not ACGT" mask                             */
#define NGSEP_M_WINDOW 0x04000000u /* payload starts with {valid_from, valid_to, first}: the record describes the
                                      read only for reference positions valid_from <= p < valid_to, with its own
                                      start `first` (versions written by the indel-realigner pre-pass)          */
#define NGSEP_M_SIMPLE 0x10000000u /* packer's promise: ONE CIGAR word, a match block as long as the read (len*8+3), no
                                      N mask, no valid window.  The device takes its fast path on it; a record without
                                      the flag goes through the general path (always correct).  ngsep_tile_submit
                                      verifies the promise against the payload.                                  */

/* Payload of one record, 32-bit words, in this order:
 *   [valid_from, valid_to, first]      only with NGSEP_M_WINDOW
 *   cigar[n_cigar]                     len*8+op, op codes H0 D1 I2 M3 P4 N5 S6 X7 (ReadAlignment.java:60-69,1180-1198)
 *   bases2[ceil(L/16)]                 2-bit A0 C1 G2 T3, base k in bits 2*(k%16) of word k/16
 *   nmask[ceil(L/32)]                  only with NGSEP_M_NMASK
 *   lowq[ceil(L/32)]                   bit k set when quality char k - 33 <= 3: the base is never counted (CountsHelper.java:214)
 *   qoff                               the read's qualities start at byte 4*qoff of ngsep_tile.qual
 * The phred+33 quality bytes themselves (ReadAlignment.qualityScores, RA:581-595) live in the separate array
 * ngsep_tile.qual, L bytes per read padded to a multiple of 4 (versions of one read share them): the window kernel never
 * touches them — they are read only for the positions that reach the genotype kernel.
 */

/* A region tile: every admitted read overlapping [tile_first, tile_last] on one sequence (halo included); calls
 * are produced only for positions inside the tile.  Stands in for the per-position loop
 * AlignmentsPileupGenerator.processPileups/processCurrentPosition (APG:453-504) over that region. */
typedef struct ngsep_tile {
    int32_t seq_idx;
    int32_t tile_first, tile_last; /* 1-based inclusive */
    int32_t n_reads;
    int32_t max_span;              /* max over records of (last reference position of the record - hdr.first + 1) */
    int32_t max_back;              /* max over records of (hdr.first - record's own first), >= 0                   */
    uint64_t n_words;              /* payload length in words                                                      */
    const ngsep_read_hdr* hdr;     /* [n_reads]                                                                    */
    const uint32_t* payload;       /* [n_words]                                                                    */
    const uint16_t* group;         /* [n_reads] read-group id of each record (see ngsep_samples_set; 0xFFFF = no read group
                                      that belongs to a sample); NULL when n_samples == 1                          */
    const uint8_t* qual;           /* [n_qual] phred+33 quality bytes, addressed through each record's qoff        */
    uint64_t n_qual;               /* bytes in qual (a multiple of 4)                                              */
    uint32_t qual_bits;            /* 0 or 8: qual holds the phred+33 bytes themselves.  2 or 4: qual holds CODES of that many
                                      bits — code k of a read at bits [k*qual_bits, (k+1)*qual_bits) of the read's run (little
                                      endian, runs padded to 4 bytes, qoff = start of the run in 4-byte units) — and qual_dict
                                      maps a code to a phred+33 byte.  The path only distinguishes min(30, q) with every q <= 3
                                      alike (CountsHelper.java:91,214-219), so a tile whose qualities fall into at most 4 / 16 such
                                      classes (binned Illumina qualities do) travels at 2 / 4 bits per base with identical results */
    uint8_t qual_dict[16];
    uint32_t layout;               /* NGSEP_LAYOUT_*: how `payload` is laid out on the caller's side                              */
    uint64_t n_wire_words;         /* NGSEP_LAYOUT_COMPACT: words in `payload` (n_words stays the length of the full layout)     */
} ngsep_tile;

#define NGSEP_LAYOUT_FULL 0u       /* payload as described above: n_words words                                                */
#define NGSEP_LAYOUT_COMPACT 1u    /* what need not cross the bus does not: the payload of a record flagged NGSEP_M_SIMPLE is
                                      [bases2, qoff] only — its CIGAR word (len*8+3) and its low-quality mask (a function of the
                                      qualities) are rebuilt on the device —, every other record's payload is unchanged; hdr.off are
                                      offsets into this stream (n_wire_words words).  The device expands the stream to the full layout at
                                      staging (one scan + one pass), so everything behind the copy is the same.  20 of the 96 bytes of a
                                      100-bp read with 2-bit quality codes stay on the host.                                      */

/* One SNV-path decision for one position.  Field semantics follow VariantDiscoverySNVQAlgorithm.discoverSNV
 * (VD:100-222) and the filter of SingleSampleVariantPileupListener.discoverVariant (SSVPL:223). */
typedef struct ngsep_snv_call {
    int32_t seq_idx;
    int32_t pos;
    int32_t count[4];     /* CountsHelper.counts, A C G T                          CH:223      */
    int32_t depth;        /* CountsHelper.totalCount                               CH:210      */
    int32_t strand[8];    /* countsStrand[b][0 = negative, 1 = positive]           CH:226-227  */
    double logcond[16];   /* logConditionalProbs[i][j], summed in pileup order     CH:231-249  */
    double post_ref;      /* posterior of the homozygous-reference genotype        VD:120      */
    double post_max;      /* posterior of the called genotype                      VD:122-123  */
    int16_t gq;           /* genotype quality                                      VD:124      */
    int16_t qual;         /* variant quality = phred(post_ref)                     VD:212      */
    int8_t a0, a1;        /* allele indexes (A0 C1 G2 T3) of the called genotype, a0 <= a1     */
    int8_t ref;           /* reference allele index                                            */
    int8_t kind;          /* NGSEP_CALL_*                                                      */
    int32_t sample;       /* sample index (0 in single-sample runs)                            */
} ngsep_snv_call;

#define NGSEP_CALL_HET 1      /* 0/1   CalledSNV heterozygous     VD:181-188 */
#define NGSEP_CALL_HOMALT 2   /* 1/1   CalledSNV homozygous alt   VD:189-193 */
#define NGSEP_CALL_TRIALLELIC 3 /* 1/2 MULTISNV                   VD:128-177; a0 = alt, a1 = third allele as ordered by VD:133-139 */
#define NGSEP_CALL_POOL 5      /* ploidy >= 3 (pools / polyploids, SSVPL:240-254): a column where an alternative base has the support
                                  createSNVVariantPool asks for (SSVPL:306-319), for the host's genotypeVariantPool (SSVPL:402-503): count /
                                  depth / strand / ref as always; sample = n, the number of counted bases; their evidence bytes
                                  (quality << 2 | base, pileup order) fill the logcond field (the first 128) and, for a deeper column, the
                                  ceil((n - 128) / 224) call slots behind this one, which hold nothing else.  A tile with known sites
                                  (ngsep_tile_sites) returns one such column per site, whatever it holds, post_max = the site's index  */
#define NGSEP_CALL_SITE 4      /* a known site (ngsep_tile_sites): VD.genotypeSNV, VD:21-97.  a0 <= a1 are indexes into the SITE's allele list
                                  (a0 == a1: homozygous, 0/0 included), -1,-1 = undecided (gq 0: no call, gq == 0, or gq < minQuality, SSVPL:388);
                                  count / depth / logcond as always; sample = index of the site in the list given to ngsep_tile_sites;
                                  qual, post_max unused */

/* ---- MultisampleVariantsDetector (ngsep_params.n_samples > 1) -------------------------------------------------------
 * One emitted population SNV site: MultisampleVariantsDetector.onPileup for a span-1 position
 * (discovery/MultisampleVariantsDetector.java:521-537): pooled allele discovery (discoverPopulationSNV, MSVD:585-597 +
 * SingleSampleVariantPileupListener.createSNVVariantPool, SSVPL:297-332), per-sample VD.genotypeSNV (VD:21-97) through
 * SSVPL.genotypeVariantSample (SSVPL:361-391, sample-level minQuality 40), QUAL = best GQ of the decided,
 * non-homozygous-reference samples (MSVD:685-691), kept when QUAL != 0 and QUAL >= minQuality (MSVD:533). */
typedef struct ngsep_pop_site {
    int32_t seq_idx;
    int32_t pos;
    int32_t count[4];     /* pooled CountsHelper.counts over ALL reads of the column (any read group)   MSVD:586-587 */
    int32_t depth;        /* pooled totalCount                                                                        */
    int16_t qs;           /* variant quality                                                           MSVD:685-691 */
    int8_t ref;           /* reference allele index (A0 C1 G2 T3)                                                     */
    int8_t n_alleles;     /* alleles of the final variant, reference first                             MSVD:590-595 */
    int8_t allele[4];     /* their base indexes; allele[0] == ref                                                     */
    int8_t n_alleles0;    /* alleles createSNVVariantPool found before the refinement loop (a variant that was reduced
                             to >2 alleles loses its MULTISNV type, SSVPL:346-359)                                    */
    int8_t pad[7];
} ngsep_pop_site;

/* One sample's genotype at a population site; ngsep_tile_fetch_pop returns n_samples of them per site, in sample order. */
typedef struct ngsep_sample_call {
    int32_t count[4];     /* the sample's CountsHelper.counts                                          SSVPL:365-367 */
    int32_t depth;        /* the sample's totalCount                                                                  */
    int16_t gq;           /* genotype quality, 0 when undecided                                        SSVPL:388     */
    int8_t g0, g1;        /* called alleles as indexes into ngsep_pop_site.allele (g0 == g1: homozygous); -1,-1 undecided */
    double logcond[10];   /* logConditionalProbs of the sample, summed in ITS pileup order: [i][i] for i = 0..3, then
                             [0][1] [0][2] [0][3] [1][2] [1][3] [2][3] (the matrix is symmetric, CH:231-249)         */
} ngsep_sample_call;

/* Counters of the last completed tile. */
typedef struct ngsep_tile_stats {
    uint64_t covered_positions;  /* positions with >= 1 covering admitted read ("genotyped positions")   */
    uint64_t candidate_positions;/* positions that went through the exact SNVQ evaluation                */
    uint64_t calls;              /* records written                                                      */
    uint64_t aligned_bases;      /* sum of read lengths of the tile's records                            */
    float kernel_ms;             /* device time of the launch that processed the tile (CUDA events on the context stream;
                                    a batched run reports the whole batch's launch for every tile in it)  */
    float pileup_ms;             /* of which: the window kernel (k_pileup_windows)                        */
    uint32_t kernel_launches;    /* kernels launched for the tile                                        */
    uint32_t open_columns;       /* positions the window kernel left to k_genotype_columns (whole launch) */
    uint64_t h2d_bytes;          /* bytes of the tile copied host -> device at staging                     */
    uint32_t quality_in_place;   /* 1: the quality bytes stayed in the caller's page-locked buffer and were read on demand */
    uint32_t d2h_bytes;          /* bytes the tile's fetch brought back device -> host (counters, staging report, call records; a pipelined tile's
                                    first call slots travel behind its kernels whether they are filled or not)  */
} ngsep_tile_stats;

/* Context = one GPU + its streams (kernels, bulk copies, genotype kernels of pipelined tiles) + the device-resident reference.  Replaces the construction of
 * AlignmentsPileupGenerator + CountsHelper's static caches (APG:38-69, CH:135-187). */
int ngsep_ctx_create(int device, const ngsep_params* params, const double* tables, size_t n_tables, void** ctx);
int ngsep_ctx_destroy(void* ctx);
const char* ngsep_last_error(void* ctx);

/* Uploads one reference sequence (ASCII, case kept, any length); packed on the device to 2 bits/base + masks.
 * Replaces ReferenceGenome.getReference for the path (genome/ReferenceGenome.java:222-245). */
int ngsep_ref_upload(void* ctx, int32_t seq_idx, const char* bases, int64_t length);

/* Asynchronous: copies the tile's host buffers to the device (page-locked buffers: on a copy stream that overlaps the kernels
 * of the previous tile) and launches the kernels.  Up to 4 tiles may be in flight; they are fetched in submission order, so
 * the host can pack / copy the next tile while the device works on the previous one (NGSEP_ESTATE when a fifth is submitted
 * before a fetch).  A tile's host buffers must stay valid until the ngsep_tile_fetch that returns it. */
int ngsep_tile_submit(void* ctx, const ngsep_tile* tile);

/* -knownVariants (discovery/SingleSampleVariantsDetector.java:897-905): declares the known SNV sites of the NEXT ngsep_tile_submit on
 * this context.  That tile is then GENOTYPED at these positions instead of being searched for variants — what
 * SingleSampleVariantPileupListener.onPileup does with input variants (SSVPL:162-176): genotypeVariantSample (SSVPL:361-391) with
 * VD.genotypeSNV (VD:21-97) at every listed site, homozygous-reference and undecided results included.  pos[i]: 1-based position inside
 * the tile; alleles[i]: the site's alleles as base indexes (A0 C1 G2 T3), reference first, 2 bits each from bit 0, and their number
 * (2..4) in bits 8-10 — two alleles = a variants/SNV (CalledSNV rules, VD:33-60), more = the sub-matrix arg-max of VD:61-95.
 * ngsep_tile_fetch returns exactly n_sites records of kind NGSEP_CALL_SITE, record i for site i (a site no read gives a call at has
 * depth 0).  The arrays are copied before the call returns.  Single-sample contexts, submit / fetch tiles only. */
int ngsep_tile_sites(void* ctx, int32_t n_sites, const int32_t* pos, const uint16_t* alleles);

/* Waits for the OLDEST submitted tile and copies its calls (unordered) into out[0..cap).  counts_dbg, when not NULL and
 * NGSEP_F_EMIT_ALL is set, receives [tile_last-tile_first+1][5] = A,C,G,T,DP.  NGSEP_EOVERFLOW with *n_out > cap: the
 * tile stays at the head of the queue, call again with room for *n_out calls. */
int ngsep_tile_fetch(void* ctx, ngsep_snv_call* out, size_t cap, size_t* n_out, int32_t* counts_dbg, ngsep_tile_stats* stats);

/* Multi-sample contexts (n_samples > 1).  ngsep_samples_set declares, once per context and before the first tile, which
 * read groups form each sample and in which order PileupRecord.getAlleleCalls(span, Set<String>) (PR:104-111) visits them
 * (the iteration order of Sample.readGroups): sample s owns sample_groups[sample_group_off[s] .. sample_group_off[s+1]).
 * ngsep_tile.group carries each record's read-group id (0 <= id < n_groups, or 0xFFFF).  ngsep_tile_submit is unchanged;
 * results come back through ngsep_tile_fetch_pop: sites[i] with its n_samples sample calls at calls[i * n_samples].
 * min_allele_depth_freq = MultisampleVariantsDetector.minAlleleDepthFrequency (MSVD:58, default 0). */
int ngsep_samples_set(void* ctx, int32_t n_groups, const int32_t* sample_group_off, const uint16_t* sample_groups, double min_allele_depth_freq);
int ngsep_tile_fetch_pop(void* ctx, ngsep_pop_site* sites, size_t cap_sites, size_t* n_sites, ngsep_sample_call* calls, int32_t* counts_dbg,
                         ngsep_tile_stats* stats);
int ngsep_resident_fetch_pop(void* ctx, int slot, ngsep_pop_site* sites, size_t cap_sites, size_t* n_sites, ngsep_sample_call* calls,
                             int32_t* counts_dbg, ngsep_tile_stats* stats);

/* Device-resident tiles: stage any number of tiles in HBM once (slots 1..n; slot 0 is the submit/fetch slot), then
 * run the kernels over them repeatedly without host<->device copies.  ngsep_resident_run(slot = -1) walks the windows
 * of EVERY resident tile in ONE kernel launch on the context stream (all sequences of a genome = one launch); results
 * are read per slot with ngsep_resident_fetch, or summed over the batch with ngsep_resident_stats (kernel_ms = device
 * time of that one launch, CUDA events on the context stream). */
int ngsep_resident_add(void* ctx, const ngsep_tile* tile, int* slot);
int ngsep_resident_clear(void* ctx);
int ngsep_resident_run(void* ctx, int slot);
int ngsep_resident_fetch(void* ctx, int slot, ngsep_snv_call* out, size_t cap, size_t* n_out, int32_t* counts_dbg, ngsep_tile_stats* stats);
int ngsep_resident_stats(void* ctx, ngsep_tile_stats* stats);

/* CoverageStatisticsCalculator (discovery/CoverageStatisticsCalculator.java:186-199, 250-256), a PileupListener that only needs
 * PileupRecord.getNumAlignments / getNumUniqueAlns: the read-depth histograms of the region [tile_first, tile_last] given the
 * admitted reads that overlap it (first / last = ReadAlignment.getFirst / getLast, unique = ReadAlignment.isUnique).
 * hist_all[d] / hist_unique[d] = positions of the tile covered by exactly d (unique) alignments for d < max_coverage,
 * [max_coverage] = positions with more.  Synchronous; the histograms of the tiles of a run are summed by the caller. */
int ngsep_coverage_histogram(void* ctx, int32_t tile_first, int32_t tile_last, int64_t n, const int32_t* first, const int32_t* last, const uint8_t* unique,
                             int32_t max_coverage, uint64_t* hist_all, uint64_t* hist_unique);

/* Indel / STR allele tallies of the events with a reference span > 1: the per-call part of
 * CountsHelper.updateCountsIndel (discovery/CountsHelper.java:253-304) over the calls of PileupRecord.getAlleleCalls(span) — calculateLogCond
 * (CH:384-396) of every call against every allele of its length, the allele the call counts for (an exact match, else the best fit,
 * CH:262-283), counts and strand counts per allele (CH:284-287), the diagonal of logConditionalProbs (CH:289) — one event per device thread,
 * calls in pileup order, every sum a sum of host-supplied table values in the reference's order (bit-identical).  The host keeps what decides
 * WHICH alleles and calls an event has (the realigner, IRPL:84-526; AlleleCallClustersBuilder, ACCB:72-141) and what hangs on libm: the
 * heterozygous mixture terms log10(f 10^x + (1-f) 10^err) of the off-diagonal entries (CH:290-300), the posteriors (CH:410-495), the record
 * (VD:265-361) — from call_index / call_fit, one addition per call and entry.
 * Event k owns alleles allele_first[k] .. allele_first[k+1]) — the reference allele first — and calls call_first[k] .. call_first[k+1]) in
 * pileup order; allele a = allele_chars[allele_off[a] .. allele_off[a+1]), call c = call_chars / call_quals (phred+33) [call_off[c] ..
 * call_off[c+1]).  log_err = log10(0.0001) (CH:52).  Out: total[k] calls of event k (CH:254); counts / strand (negative, positive) / diag per
 * allele; call_index[c] = the allele call c counts for (-1: none), call_fit[c] = its log-likelihood under that allele.  Synchronous. */
typedef struct ngsep_indel_events {
    int32_t n_events, n_alleles, n_calls, reserved;
    const int32_t* allele_first;   /* [n_events + 1] */
    const int32_t* call_first;     /* [n_events + 1] */
    const int32_t* allele_off;     /* [n_alleles + 1] */
    const int32_t* call_off;       /* [n_calls + 1]   */
    const char* allele_chars;
    const char* call_chars;
    const uint8_t* call_quals;
    const uint8_t* call_neg;       /* [n_calls] 1 = negative strand */
    double log_err;
} ngsep_indel_events;
int ngsep_indel_tallies(void* ctx, const ngsep_indel_events* events, int32_t* total, int32_t* counts, int32_t* strand, double* diag, int32_t* call_index,
                        double* call_fit);

/* RelativeAlleleCountsCalculator (discovery/RelativeAlleleCountsCalculator.java:246-290), a PileupListener over PileupRecord.getAlleleCalls(1):
 * for every position of the tile, out[2 i] = (largest allele count) | (number of alleles seen) << 28 and out[2 i + 1] = the second largest
 * count, counting the calls whose base quality is >= min_base_qs; both words are 0 at a position with fewer than min_rd calls or without a
 * counted call (the listener records nothing there).  Alleles are A, C, G, T and "any other base".  Synchronous; no tile may be in flight.
 * With coded qualities (ngsep_tile.qual_bits 2 / 4) min_base_qs must lie in 4..30 — the codes keep min(30, q) and every q <= 3 alike. */
int ngsep_tile_allele_counts(void* ctx, const ngsep_tile* tile, int32_t min_rd, int32_t min_base_qs, uint32_t* out);

/* Synthetic inputs generated ON THE DEVICE — BASELINE configs[2] and [4] (64 Mb at 60x, 3.1 Gb at 30x, the 1000x panel): the
 * reference sequence and the packed tile of any region are pure functions of (seed, sequence, position / read index), so no
 * input crosses PCIe and every tile of a genome sees the same reads on whatever GPU.  Rates as the reference's simulator
 * (simulation/SingleIndividualSimulator.java:57-58); reads are one match block each.  ngsep_synth_tile stages the tile as a
 * resident slot (ngsep_resident_run / _fetch as usual); ngsep_resident_download / ngsep_ref_download bring a tile and its
 * reference back so that a slice can be handed to the CPU oracle as SAM text. */
typedef struct ngsep_synth_params {
    uint64_t seed;
    double depth;       /* read_len * reads / positions                          */
    int32_t read_len;   /* 16..1000                                              */
    int32_t reserved;
    double snv_rate;    /* planted SNV sites per base, half of them heterozygous */
    double error_rate;  /* substitution errors per read base                     */
} ngsep_synth_params;
int ngsep_synth_reference(void* ctx, int32_t seq_idx, int64_t length, uint64_t seed, double gc);
int ngsep_synth_tile(void* ctx, int32_t seq_idx, int32_t tile_first, int32_t tile_last, const ngsep_synth_params* params, int* slot);
/* shape: extents, counts and quality coding of the resident tile; hdr / payload / qual (each may be NULL) receive shape->n_reads
 * headers, shape->n_words words, shape->n_qual bytes. */
int ngsep_resident_download(void* ctx, int slot, ngsep_tile* shape, ngsep_read_hdr* hdr, uint32_t* payload, uint8_t* qual);
int ngsep_ref_download(void* ctx, int32_t seq_idx, char* bases, int64_t length);

/* CUDA-event stopwatch on the context stream (ngsep_timer_stop also waits for the genotype kernels of pipelined tiles). */
int ngsep_timer_start(void* ctx);
int ngsep_timer_stop(void* ctx, float* elapsed_ms);   /* synchronises the stream */

/* Page-locks / releases a caller-owned host buffer (what a JVM host does for its direct ByteBuffers) so that
 * ngsep_tile_submit copies asynchronously at full PCIe rate.  With NGSEP_F_QUALITY_IN_PLACE in ngsep_params.flags and a
 * page-locked ngsep_tile.qual, a single-sample ngsep_tile_submit does not copy the qualities at all: they are only needed
 * at the few positions that reach the genotype kernel, which then reads them in place (ngsep_tile_stats.quality_in_place;
 * NGSEP_LAZY_QUAL=0/1 in the environment overrides the flag).  The buffers must stay valid and unchanged until the tile
 * has been fetched — as before. */
int ngsep_host_register(void* ptr, size_t bytes);
int ngsep_host_unregister(void* ptr);

/* Library/ABI version (major*1000+minor) and the compute capability the kernels were built for. */
int ngsep_abi_version(void);   /* 2004: + ngsep_tile.layout / n_wire_words (compact wire layout), ngsep_indel_tallies, ngsep_tile_allele_counts; 2003: + ngsep_tile_sites (known sites); 2002: + ngsep_tile.qual_bits / qual_dict; 2001: + sample axis (ngsep_samples_set, ngsep_tile_fetch_pop); 2000: quality bytes in their own array */
int ngsep_built_sm(void);

#ifdef __cplusplus
}
#endif
#endif /* NGSEP_PILEUP_H */
