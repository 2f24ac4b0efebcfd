/*
 * ngsep_host.h — C entry points of libngsep_host.so, the host pipeline that sits ABOVE the device C ABI
 * (ngsep_pileup.h).  It mirrors, for this path only, the reference's detector surface
 *   src/ngsep/discovery/SingleSampleVariantsDetector.java   (setters :141-560, run() :589-656, findSNVs :896-931)
 *   src/ngsep/discovery/AlignmentsPileupGenerator.java      (APG:71-193, processAlignments APG:334-361)
 * over decoded alignment records (what htsjdk hands to ReadAlignmentFileReader.loadAlignment, RAFR:219-272).
 * In a JVM deployment the Java host does this work itself and binds ngsep_pileup.h directly;
 * this library exists because no JVM is available in the build image.
 */
#ifndef NGSEP_HOST_H
#define NGSEP_HOST_H

#include "ngsep_pileup.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Decoded, coordinate-sorted alignment records, struct-of-arrays. */
typedef struct ngsep_host_batch {
    int64_t n;
    const int32_t* seq_idx;     /* index into the genome set with ngsep_ssvd_set_genome                 */
    const int32_t* first;       /* SAMRecord.getAlignmentStart, 1-based                                  */
    const uint16_t* flags;      /* SAM flags                                                             */
    const uint8_t* mapq;
    const int16_t* rg;          /* read-group index, -1 = no RG tag (read group "" in the reference)     */
    const int16_t* nh;          /* NH tag, -1 = absent (RAFR:284-291)                                    */
    const int64_t* cigar_off;   /* [n+1]                                                                 */
    const uint32_t* cigar;      /* len*8+op, op codes H0 D1 I2 M3 P4 N5 S6 X7 ('=' is 3), RA:60-69       */
    const int64_t* base_off;    /* [n+1]                                                                 */
    const uint8_t* bases;       /* SEQ, ASCII                                                            */
    const uint8_t* quals;       /* QUAL, phred+33; a missing QUAL is passed as '+' bytes (RA:1025)       */
    const uint64_t* name_hash;  /* optional hash of QNAME for the duplicate-record rule (RAFR:292-306)   */
} ngsep_host_batch;

/* Option ids of main/CommandsDescriptor.xml:565-703 that affect this path, plus device plumbing. */
typedef struct ngsep_detector_options {
    const char* sample_id;          /* -sampleId, NULL = "Sample"   */
    int32_t ploidy;                 /* -ploidy, 2                   */
    int32_t print_sample_ploidy;    /* -psp                         */
    int32_t min_mq;                 /* -minMQ, 20                   */
    const char* query_seq;          /* -querySeq, NULL = none       */
    int32_t query_first, query_last;/* -first / -last               */
    int32_t ignore_lower_case_ref;  /* -ignoreLowerCaseRef          */
    int32_t max_alns_per_start_pos; /* -maxAlnsPerStartPos, 5       */
    int32_t process_non_unique_primary; /* -p                       */
    int32_t process_secondary;      /* -s                           */
    int32_t ignore5, ignore3;       /* -ignore5 / -ignore3          */
    double het_rate;                /* -h, 0.001                    */
    int32_t het_rate_set;           /* 1 when -h was given (SSVD:591-593 haploid default otherwise) */
    int32_t max_base_qs;            /* -maxBaseQS, 30               */
    int32_t min_quality;            /* -minQuality, 40              */
    int32_t call_embedded_snvs;     /* -embeddedSNVs                */
    int32_t calc_strand_bias;       /* -csb                         */
    int32_t device;                 /* CUDA device index            */
    int32_t tile_len;               /* region tiles of this many positions; 0 = automatic (2 Mb: shorter sequences are one tile) */
    int32_t shard, n_shards;        /* region sharding across ranks: tile t belongs to rank t % n_shards */
    int32_t host_only;              /* test hook: admission + event pre-pass only; output holds just the host-made indel/STR records */
    int32_t synchronous;            /* test hook: one tile at a time from pageable buffers, host stages on the calling thread (default: sequences planned
                                       in parallel, tiles pipelined through page-locked ring buffers) */
    int32_t n_devices;              /* GPUs device .. device + n_devices - 1: region tiles are dealt to them in turn (0 / 1 = one GPU) */
    const char* known_strs;         /* -knownSTRs: file with tandem-repeat loci (sequence, first, last in its first three columns), NULL = none;
                                       SingleSampleVariantsDetector only (SSVD:906-913) */
    const char* known_variants;     /* -knownVariants: VCF with the variants to genotype instead of discovering variants (SSVD:897-905), NULL = none;
                                       SingleSampleVariantsDetector only.  Wins over known_strs, as in the reference */
} ngsep_detector_options;

typedef struct ngsep_run_stats {
    uint64_t records, admitted, versions, events, indel_calls, snv_calls;
    uint64_t covered_positions, candidate_positions, aligned_bases, tiles, kernel_launches;
    double kernel_ms, prepass_s, pack_s, device_s, total_s;
} ngsep_run_stats;

void ngsep_detector_options_default(ngsep_detector_options* o);
int ngsep_ssvd_create(const ngsep_detector_options* o, void** det);
/* MultisampleVariantsDetector (src/ngsep/discovery/MultisampleVariantsDetector.java: run() :421-462): the detector handle is
 * used with the same ngsep_ssvd_set_genome / _run / _vcf / _stats / _destroy calls.  rg_ids / rg_samples = the @RG table of the
 * input(s) (ID and SM, SM NULL or "" when absent: the sample is then the read-group id, RAFR:188); ngsep_host_batch.rg indexes
 * it.  Samples and their column order follow MSVD.loadSamplesFromAlignmentHeaders (MSVD:492-516). */
int ngsep_msvd_create(const ngsep_detector_options* o, int n_read_groups, const char* const* rg_ids, const char* const* rg_samples,
                      double min_allele_depth_frequency, void** det);
int ngsep_ssvd_destroy(void* det);
const char* ngsep_ssvd_error(void* det);
/* setGenome (SSVD:161-163): sequence names, lengths and bases exactly as in the FASTA (case kept). */
int ngsep_ssvd_set_genome(void* det, int n, const char* const* names, const int64_t* lengths, const char* const* bases);
/* findSNVs over the batch; the VCF text (header when with_header != 0) is kept in the handle. */
int ngsep_ssvd_run(void* det, const ngsep_host_batch* batch, int with_header);
const char* ngsep_ssvd_vcf(void* det, size_t* len);
int ngsep_ssvd_stats(void* det, ngsep_run_stats* stats);

/* AlignmentsPileupGenerator.processFiles (APG:212-289) for callers that hold one batch per input file: the order in which the
 * records of the inputs are consumed — each input reduced to what its own reader hands over (record filter + duplicate-record
 * rule), then a k-way merge by (sequence, first, last), ties to the lowest input.  *input_idx / *record_idx: n_out pairs,
 * malloc'ed (ngsep_host_free).  Gather the records in that order into ONE batch (without name hashes: the duplicate rule has
 * been applied) and hand it to the multi-sample detector. */
int ngsep_host_merge_order(int n_inputs, const ngsep_host_batch* inputs, int min_mq, int process_non_unique_primary, int process_secondary,
                           int32_t** input_idx, int64_t** record_idx, int64_t* n_out);

/* Admission + packing of one sequence without the event pre-pass (device-resident benchmarking, packer tests). */
int ngsep_host_pack(const ngsep_host_batch* b, int seq_idx, int min_mq, int max_alns_per_start, int ignore5, int ignore3,
                    ngsep_read_hdr** hdr, uint32_t** payload, uint8_t** qual, int64_t* n_reads, uint64_t* n_words, uint64_t* n_qual, int* max_span,
                    uint64_t* aligned_bases);
/* The tile the detector itself submits for a whole sequence: admission + event pre-pass (read versions) + packing. */
int ngsep_ssvd_pack(void* det, const ngsep_host_batch* b, int seq_idx, ngsep_read_hdr** hdr, uint32_t** payload, uint8_t** qual, int64_t* n_reads,
                    uint64_t* n_words, uint64_t* n_qual, int* max_span, int* max_back, int* tile_first, int* tile_last, uint64_t* aligned_bases);
void ngsep_host_free(void* p);
/* CoverageStats module (discovery/CoverageStatisticsCalculator.java:119-143, 250-256; option ids -i -o -minMQ of main/CommandsDescriptor.xml:458-478):
 * admission with secondary alignments and 100 alignments per start position on the host, read-depth histograms per region tile on the
 * device (ngsep_coverage_histogram), the printed table (depth, positions, positions counting unique alignments only; "More") as text. */
int ngsep_coverage_stats(int n_seqs, const char* const* names, const int64_t* lengths, const ngsep_host_batch* batch, int min_mq, int device, char** text);
/* RelativeAlleleCountsCalculator module (discovery/RelativeAlleleCountsCalculator.java:209-243, 246-290; option ids -i -o -m -M -q -s of
 * main/CommandsDescriptor.xml:805-847): admission with maxRD alignments per start position on the host, the
 * allele counts of every position per region tile on the device (ngsep_tile_allele_counts), the three distributions of the report as text. */
int ngsep_relative_allele_counts(int n_seqs, const char* const* names, const int64_t* lengths, const ngsep_host_batch* batch, int min_rd, int max_rd,
                                 int min_base_qs, int secondary, int device, char** text);
/* ... with the region files of -frs (regions to filter) and -srs (regions to select), NULL = none (RelativeAlleleCountsCalculator.java:131-161,
 * 247-254; genome/io/SimpleGenomicRegionFileHandler.java:57-80: sequence, first, last in the first three blank- or tab-separated fields). */
int ngsep_relative_allele_counts_regions(int n_seqs, const char* const* names, const int64_t* lengths, const ngsep_host_batch* batch, int min_rd, int max_rd,
                                         int min_base_qs, int secondary, const char* regions_to_filter, const char* regions_to_select, int device, char** text);
/* How the quality array of the tile the calling thread packed last (ngsep_host_pack / ngsep_ssvd_pack) is coded: the values for
 * ngsep_tile.qual_bits and ngsep_tile.qual_dict. */
/* ... and how its payload is laid out: the values for ngsep_tile.layout and ngsep_tile.n_words (the detector packs NGSEP_LAYOUT_COMPACT tiles —
 * ngsep_ssvd_pack returns the compact stream, its length is the n_words it reports = ngsep_tile.n_wire_words —, ngsep_host_pack the full layout). */
int ngsep_host_pack_layout(uint32_t* layout, uint64_t* n_words_full);
int ngsep_host_pack_quality(uint32_t* qual_bits, uint8_t* qual_dict16);

#ifdef __cplusplus
}
#endif
#endif /* NGSEP_HOST_H */
