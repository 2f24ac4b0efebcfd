// Host pipeline above the C ABI: event pre-pass (realigner + host indel calls), tile packing, device calls,
// ordered merge and VCF text.  Mirrors the public surface of the reference's detector
// (src/ngsep/discovery/SingleSampleVariantsDetector.java: setters of :141-560, run() :589-656) and of
// AlignmentsPileupGenerator (APG:71-193) for this path only.
#pragma once
#include <array>
#include <cstdint>
#include <deque>
#include <functional>
#include <map>
#include <set>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../../include/ngsep_pileup.h"
#include "model.hpp"

namespace ngsep_b200 {
namespace host {

struct RunStats {
    uint64_t records = 0, admitted = 0, versions = 0, events = 0, indel_calls = 0, snv_calls = 0;
    uint64_t indel_calls_late = 0;   // (per sequence) indel records made after the device's tallies
    uint64_t covered_positions = 0, candidate_positions = 0, aligned_bases = 0, tiles = 0, kernel_launches = 0;
    double kernel_ms = 0, prepass_s = 0, pack_s = 0, device_s = 0, total_s = 0;   // prepass_s: wall time of the (parallel) host planning stage
    void add_host(const RunStats& o) { admitted += o.admitted; versions += o.versions; events += o.events; indel_calls += o.indel_calls; }
};

struct EventOutcome { bool fired = false; int span = 1; bool str = false, new_str = false; };

// An input tandem-repeat locus of the realigner (-knownSTRs): first .. last after merging and padding (SSVD:843-894)
struct StrLocus {
    int first = 0, last = 0;
    bool str = true;       // the input variant is a tandem repeat (PileupRecord.setSTR, IRPL:91)
    bool starter = true;   // the realigner sees the variant START at `first` (IRPL:90); false: only the tail of an input variant that an
                           // earlier, overlapping one hides (IRPL:141-157 returns the earlier variant while it lasts): every position embedded
};
// per sequence of the genome, sorted by position (strs.cpp)
bool load_known_strs(const std::string& path, const Genome& genome, std::vector<std::vector<StrLocus>>& out, std::string& err);

// -knownVariants (SSVD:897-905): one variant of the input VCF as vcf/VCFFileReader.loadVariants keeps it (VCFFileReader.java:192-255,
// 261-306, 585-600): a variants/SNV (two alleles, both one base of ACGT) or a GenomicVariantImpl
struct KnownVariant {
    int first = 0, last = 0;
    std::vector<std::string> alleles;      // reference first, as written in the file
    int type = 0;                          // GenomicVariant.TYPE_* (0 undetermined, 1 SNV, 2 MULTISNV, 3 EMBEDDED, 4 INDEL, 5 STR)
    int qs = 0;                            // QUAL column, rounded
    std::string id;                        // ID column, "" = none
    bool snv_class = false;                // instance of variants/SNV
    bool is_snv = false;                   // GenomicVariant.isSNV(): genotyped on the device as a known site
    uint16_t site_code = 0;                // the site's alleles as ngsep_tile_sites takes them
};
// per sequence of the genome, in the order of GenomicRegionSortedCollection.getSequenceRegions(name).asList(): first, then last, stable
bool load_known_variants(const std::string& path, const Genome& genome, std::vector<std::vector<KnownVariant>>& out, std::string& err);
// what IndelRealignerPileupListener.intersectWithVariants (IRPL:141-157) returns along a sequence, as disjoint loci sorted by position
std::vector<StrLocus> known_variant_loci(const std::vector<KnownVariant>& vars);

struct Realigner {
    const Genome* genome = nullptr;
    int seq = 0;
    // IRPL:84-216.  locus: the input STR that starts at pos (the event is then fixed to the locus: every indel inside it moves to its
    // start, IRPL:90-95, 169, 205-206), or null (an event is found from the reads' own indels)
    EventOutcome on_event(int pos, std::vector<HostRead*>& pileup, const StrLocus* locus = nullptr);
    void process_ends(std::vector<HostRead*>& pileup, int ev_first, int ev_last);  // IRPL:400-526
};

// A called indel / STR (host side), already reduced to what the VCF record needs.
struct IndelRecord {
    int pos = 0, last = 0;
    std::vector<std::string> alleles;      // reference first
    bool str = false;
    int qual = 0, gq = 0, depth = 0;
    std::vector<int> called;               // allele indexes of the genotype (1 or 2 entries)
    std::vector<int> counts;               // per allele
    std::vector<std::vector<double>> logc; // per allele pair
    int strand_bias = -1;
    bool pool = false;                     // a pool / polyploid call (SSVPL:281-295): every allele of the variant is listed, the variant has no type and no quality
};

// One event with a reference span > 1 as the variant listener sees it (SSVPL:257-273): the calls of PR.getAlleleCalls(span) in pileup
// order and the alleles AlleleCallClustersBuilder proposes for them (reference first)
struct SpanCall { std::string allele, quals; bool neg = false; };
struct SpanEventInput { int pos = 0; std::string ref_allele; bool is_str = false, input_str = false; std::vector<SpanCall> calls; std::vector<std::string> alleles; };
// CountsHelper.calculateCountsIndel + posteriors + the arg-max genotype of one event: counts, strand counts, logConditionalProbs,
// P(ref, ref), the genotype (m0 <= m1) and its probability
struct IndelTally { int total = 0; std::vector<int> counts; std::vector<std::array<int, 2>> strand; std::vector<std::vector<double>> L; double p00 = 0, max_p = 0; int m0 = 0, m1 = 0; };

// The per-call part of CountsHelper.updateCountsIndel of one event (the device's k_indel_tallies): per allele counts, strand counts and the
// diagonal of logConditionalProbs; per call the allele it counts for (-1: none) and its log-likelihood under that allele
struct CallTallies { int total = 0; std::vector<int> counts; std::vector<std::array<int, 2>> strand; std::vector<double> diag; std::vector<int> call_index; std::vector<double> call_fit; };

struct IndelCallerOptions { int max_base_qs = 30; double het_rate = 0.001; int min_quality = 40; bool csb = false; int ploidy = 2; };

// SingleSampleVariantPileupListener.discoverVariantWithSpan (SSVPL:257-273) up to the SNV fall-back:
// AlleleCallClustersBuilder + CountsHelper.calculateCountsIndel + VariantDiscoverySNVQAlgorithm.callIndel.
// input_str: the span is an input tandem-repeat locus (PileupRecord.isInputSTR): a call without a length change is kept (VD:307)
bool call_indel_at(int pos, const std::string& ref_allele, bool is_str, std::vector<HostRead*>& pileup, const IndelCallerOptions& opt, IndelRecord& out, bool input_str = false);

// The three steps of call_indel_at: gather (PR.getAlleleCalls(span) + AlleleCallClustersBuilder), tally (the device's k_indel_tallies in
// single-sample runs; tally_span_event_host for the sample axis and the host-only test hook), finish (VD.callIndel's record + SSVPL:263)
void gather_span_event(int pos, const std::string& ref_allele, bool is_str, bool input_str, std::vector<HostRead*>& pileup, int max_base_qs, SpanEventInput& out);
void tally_span_event_host(const SpanEventInput& in, const IndelCallerOptions& opt, IndelTally& t);
void tally_calls_host(const SpanEventInput& in, const IndelCallerOptions& opt, CallTallies& ct);
void finish_tally(const CallTallies& ct, size_t n_alleles, double het_rate, IndelTally& t);   // mixture terms, posteriors, arg-max genotype
double indel_log_error();   // log10(0.0001), CH:52
bool finish_indel_call(const SpanEventInput& in, const IndelTally& t, const IndelCallerOptions& opt, IndelRecord& out);

// Pools / polyploids (ploidy >= 3, pool.cpp): the genotype genotypeVariantPool gives (SSVPL:402-503)
struct PoolCall { std::vector<int> called; int gq = 0, depth = 0; bool undecided = false; std::vector<int> counts; std::vector<std::vector<double>> logc; };
// discoverSNV's pool branch (SSVPL:240-254) on a column of kind NGSEP_CALL_POOL and its evidence bytes; alleles: base indexes, reference first
bool discover_pool_snv(const ngsep_snv_call& col, const uint8_t* ev, int ploidy, double het_rate, int min_quality, std::vector<int>& alleles, PoolCall& out);
// discoverIndel's pool branch (SSVPL:281-295) + the filter of SSVPL:263 from the event's allele tallies
bool finish_indel_pool(const SpanEventInput& in, const CallTallies& ct, const IndelCallerOptions& opt, int ploidy, IndelRecord& out);
std::string pool_snv_vcf_line(int pos, const std::vector<int>& alleles, const PoolCall& c, const std::string& chrom, int ploidy, bool embedded);

// Phred / rounding helpers shared by the host writer (math/PhredScoreHelper.java:31-40)
long jround(double x);
int phred(double p);
double fisher_pvalue(int a, int b, int c, int d);   // math/FisherExactTest.java:65-136

// ---- sample axis: MultisampleVariantsDetector ------------------------------------------------------------------------
// One sample = the read groups whose @RG SM names it (MSVD:492-516); rgs holds batch read-group indexes in the order
// PileupRecord.getAlleleCalls(span, Set<String>) visits them (PR:104-111: the iteration order of Sample.readGroups, a HashSet).
struct SampleDef { std::string id; int ploidy = 2; std::vector<int> rgs; };
// MSVD.loadSamplesFromAlignmentHeaders over the @RG table of the input (ids, SM values; SM "" = absent, RAFR:188).
std::vector<SampleDef> samples_from_read_groups(const std::vector<std::string>& rg_ids, const std::vector<std::string>& rg_samples, int ploidy);

// One sample's genotype of a population indel / STR (VD.callIndel with the variant given, VD:265-361, through
// SSVPL.genotypeVariantSample, SSVPL:361-391).
struct SampleIndelCall {
    std::vector<int> called;               // allele indexes, empty = undecided
    int gq = 0, depth = 0;
    bool has_report = false;               // VariantCallReport present (totalDepth > 0, VD:350)
    std::vector<int> counts;               // per variant allele
    std::vector<std::vector<double>> logc; // per allele pair
};
struct PopIndelRecord {
    int pos = 0, last = 0;
    std::vector<std::string> alleles;      // reference first
    bool str = false;
    int qs = 0;
    std::vector<SampleIndelCall> calls;    // one per sample
};
// MSVD.discoverPopulationVariantWithSpan (MSVD:599-634) + genotypeVariant (MSVD:674-693) + the filter of MSVD:533.
// Returns 0: no indel variant — the SNV fall-back applies and the device's decision for the position stands;
//         1: an indel variant that is not reported (nothing is reported at the position);  2: rec holds the record.
int call_population_indel_at(int pos, const std::string& ref_allele, bool is_str, std::vector<HostRead*>& pileup, const std::vector<SampleDef>& samples,
                             const IndelCallerOptions& opt, PopIndelRecord& rec);

// SSVPL.genotypeVariantSample for a known variant that is not an SNV (SSVPL:378-388): PR.getAlleleCalls(reference allele length) +
// CountsHelper.calculateCountsIndel over the variant's alleles + VD.callIndel with the variant given, then minQuality => undecided.
SampleIndelCall genotype_known_indel_at(int pos, const std::vector<std::string>& alleles, std::vector<HostRead*>& pileup, const IndelCallerOptions& opt);
// ... with ploidy >= 3: the pool branch of genotypeVariantSample (SSVPL:366-377), pool.cpp
SampleIndelCall genotype_known_snv_pool(const ngsep_snv_call* col, const uint8_t* ev, const std::vector<int>& alleles, int ploidy, double het_rate, int min_quality);
SampleIndelCall genotype_known_indel_pool_at(int pos, const std::vector<std::string>& alleles, std::vector<HostRead*>& pileup, const IndelCallerOptions& opt);
std::string known_pool_snv_vcf_line(const KnownVariant& v, const ngsep_snv_call* col, const SampleIndelCall& c, const std::string& chrom, int ploidy);
// VCF records of a -knownVariants run (SSVD:933-968): c == nullptr: no record gave a call (an undecided call without data, VD:23-27)
std::string known_snv_vcf_line(const KnownVariant& v, const ngsep_snv_call* c, const std::string& chrom, int ploidy);
std::string known_indel_vcf_line(const KnownVariant& v, const SampleIndelCall& c, const std::string& chrom, int ploidy);

// Grow-only host array for tile buffers.  With `pin` set the allocation is page-locked through ngsep_host_register (what a
// JVM host does for its direct ByteBuffers), so ngsep_tile_submit copies it asynchronously; a registration that fails (no
// device: host-only runs) leaves the memory pageable.  Contents are NOT kept across a growth.
template <typename T>
struct HostBuf {
    T* p = nullptr;
    size_t n = 0, cap = 0;
    bool pin = false, registered = false;
    HostBuf() = default;
    HostBuf(const HostBuf&) = delete;
    HostBuf& operator=(const HostBuf&) = delete;
    HostBuf(HostBuf&& o) noexcept { *this = std::move(o); }
    HostBuf& operator=(HostBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; cap = o.cap; pin = o.pin; registered = o.registered; o.p = nullptr; o.n = o.cap = 0; o.registered = false; }
        return *this;
    }
    ~HostBuf() { release(); }
    void release();
    void reserve(size_t want, bool exact = false);   // (discards the contents when it has to grow; exact: without the growth margin)
    void resize(size_t k) { reserve(k); n = k; }
    void clear() { n = 0; }
    T* data() { return p; }
    const T* data() const { return p; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
    T* begin() { return p; }
    T* end() { return p + n; }
    const T* begin() const { return p; }
    const T* end() const { return p + n; }
};

struct PackedTile {
    HostBuf<ngsep_read_hdr> hdr;
    HostBuf<uint32_t> payload;
    HostBuf<uint16_t> group;     // read-group id per record (0xFFFF = none), filled when pack_reads is given with_groups
    HostBuf<uint8_t> qual;       // phred+33 bytes, one run per read (padded to 4), addressed by the records' qoff
    uint32_t qual_bits = 8;      // how `qual` is coded: 8 = the bytes, 2 / 4 = codes into qual_dict (ngsep_tile.qual_bits)
    uint8_t qual_dict[16] = {0};
    uint32_t layout = NGSEP_LAYOUT_FULL;   // how `payload` is laid out (ngsep_tile.layout); compact: n_words_full = its length once expanded on the device
    uint64_t n_words_full = 0;
    int max_span = 1, max_back = 0;
    uint64_t aligned_bases = 0;
    void set_pinned(bool on) { hdr.pin = payload.pin = group.pin = qual.pin = on; }
    ngsep_tile view(int seq_idx, int first, int last) const;
};
// Packs reads[lo,hi) (all their versions) in pileup order, on up to host_threads() threads.  Returns false (and leaves `err`)
// when a record cannot be represented in the device layout (more than 65535 bases or 255 CIGAR operations — also after the
// realigner rewrote it).
bool pack_reads(const std::vector<HostRead>& reads, size_t lo, size_t hi, PackedTile& out, bool with_groups = false, std::string* err = nullptr, int min_qual_bits = 0,
                bool compact = false);
// How the previous tile's qualities were coded (PackedTile.qual_bits / qual_dict): a caller that packs tile after tile hands it on, and
// the packer skips its pass over the qualities as long as the tiles keep to the same classes
struct QualHint { bool valid = false; uint32_t bits = 8; uint8_t dict[16] = {0}; };
struct SeqPlan;
// ... the same for reads[lo,hi) of a plan (SeqPlan::lite): untouched single-block reads are packed straight from the batch
bool pack_plan_reads(const SeqPlan& plan, size_t lo, size_t hi, PackedTile& out, bool with_groups = false, std::string* err = nullptr, int min_qual_bits = 0,
                     bool compact = false, QualHint* hint = nullptr);
bool wire_compact();   // the detectors pack NGSEP_LAYOUT_COMPACT tiles unless NGSEP_WIRE_COMPACT=0
int host_threads();   // NGSEP_HOST_THREADS, else the hardware concurrency (at most 32)

struct DetectorOptions {   // option ids of main/CommandsDescriptor.xml:565-703 that affect this path
    std::string sample_id = "Sample";
    int ploidy = 2;
    bool print_sample_ploidy = false;
    int min_mq = 20;
    std::string query_seq; bool has_query = false; int query_first = 0, query_last = 1000000000;
    bool ignore_lower = false;
    int max_alns_per_start = 5;
    bool non_unique_primary = false, secondary = false;
    int ignore5 = 0, ignore3 = 0;
    double het_rate = 0.001; bool het_modified = false;
    int max_base_qs = 30;
    int min_quality = 40;
    bool call_embedded = false, csb = false;
    double min_allele_depth_freq = 0;   // MultisampleVariantsDetector.minAlleleDepthFrequency (MSVD:58)
    std::string known_strs;             // -knownSTRs: file with tandem-repeat loci (sequence, first, last); SingleSampleVariantsDetector only
    std::string known_variants;         // -knownVariants: VCF with the variants to genotype (SSVD:897-905); SingleSampleVariantsDetector only
    // device plumbing (not reference options)
    int device = 0;
    int n_devices = 1;         // GPUs device .. device + n_devices - 1 of this host: region tiles are dealt to them in turn, every GPU with its own
                               // context, ring of page-locked tile buffers and tiles in flight (no collective, ordered merge on the host)
    int tile_len = 0;          // region tiles of this many positions; 0 = automatic (kAutoTileLen: sequences up to that length are one tile)
    bool synchronous = false;  // test hook: one tile at a time from pageable buffers, host stages on the calling thread (the pipelined run
                               // must give the same bytes)
    int shard = 0, n_shards = 1;   // region sharding across ranks: this process handles tiles t % n_shards == shard
    bool host_only = false;    // test hook: run admission + event pre-pass only and report the host-made indel/STR records
                               // (no SNV calling happens anywhere in this mode; it is not a fallback)
};

constexpr int kAutoTileLen = 2000000;

// Outcome of the host stages for one sequence (admission + event pre-pass).
struct SeqPlan {
    // The admitted reads in pileup order, 16 bytes each; the reads that are more than one untouched match block — any other CIGAR, or
    // a place in the pileup of a realigner event — have a HostRead in `heavy` (a deque: its elements never move)
    std::vector<LiteRead> lite;
    std::deque<HostRead> heavy;
    const BatchView* batch = nullptr;            // the records behind `lite` (valid for the run that made the plan)
    int64_t rec_base = 0;                        // LiteRead.rec counts from this record of the batch
    int ignore5 = 0, ignore3 = 0;                // RA:613-644 for the reads without a HostRead
    std::vector<IndelRecord> indels;             // host-made indel / STR calls
    std::vector<PopIndelRecord> pop_indels;      // the same for a multi-sample run
    std::set<int> suppressed;                    // positions whose device SNV decision must not be reported
    std::vector<std::pair<int, int>> embedded;   // closed intervals of embedded positions (SSVPL:150-151)
    // single-sample runs: the tallies of the span > 1 events are made on the device (ngsep_indel_tallies), so the listener's sequential part
    // (SSVPL:146-161: what is embedded, which calls stand) runs after the sweep, over the events it fired, in order
    struct FiredEvent { int e = 0; EventOutcome ev; bool ref_ok = false, ref_lower = false; int input = -1; };
    std::vector<FiredEvent> fired;
    std::vector<SpanEventInput> span_inputs;     // FiredEvent.input indexes these (events with a reference span > 1 and a reference allele)
    std::vector<IndelTally> span_tallies;
    std::vector<CallTallies> span_call_tallies;  // (pools: the per-call tallies themselves, finished per allele-frequency hypothesis)
    // -knownVariants: the input variants that got a pileup (SSVPL:162-176), as indexes into the sequence's variant list
    std::vector<int> sites;                                      // SNVs: genotyped on the device (ngsep_tile_sites), ascending position
    std::vector<std::pair<int, SampleIndelCall>> known_indels;   // the others: genotyped by the host at their event
    int win_first = 1, win_last = 0, max_reach = 1, max_back = 0;
    RunStats stats;                              // what the host stages of this sequence add to the run's statistics
    std::string err;
};

// One sequence of a run: its plan (host stages), what came back from the device, and its VCF records.
struct SeqJob {
    int seq = 0;
    int64_t lo = 0, hi = 0;                      // records of the batch
    SeqPlan plan;
    std::vector<ngsep_snv_call> snvs;            // single-sample device results
    struct PoolColumn { ngsep_snv_call col; std::vector<uint8_t> ev; };
    std::vector<PoolColumn> pool_cols;           // ploidy >= 3: the supported columns of the device with their evidence bytes
    std::vector<ngsep_snv_call> site_calls;      // -knownVariants: the device's record of plan.sites[k] (site_have[k] != 0)
    std::vector<char> site_have;
    std::vector<std::vector<uint8_t>> site_ev;   // ... with ploidy >= 3: the evidence bytes of the site's column
    std::vector<ngsep_pop_site> sites;           // multi-sample device results ...
    std::vector<ngsep_sample_call> scalls;       // ... n_samples per site
    uint64_t snv_records = 0;
    int rc = 0;
    std::string text;
};

// Drop-in for SingleSampleVariantsDetector.findSNVs over in-memory alignments.
class GpuSingleSampleVariantsDetector {
public:
    explicit GpuSingleSampleVariantsDetector(const DetectorOptions& o);
    virtual ~GpuSingleSampleVariantsDetector();
    void set_genome(const Genome* g);                 // setGenome, SSVD:161-163
    // Starts creating the device contexts on a thread of their own (CUDA start-up takes most of a second): a caller that still has its
    // inputs to read calls this first; run() waits for it.  Optional.
    void warm_up();
    // Processes a coordinate-sorted batch; appends VCF text (header included when with_header) to out.
    int run(const BatchView& batch, std::string& out, bool with_header = true);
    const RunStats& stats() const { return stats_; }
    const std::string& error() const { return err_; }
    static std::string vcf_header(const std::string& sample_id, int ploidy, bool psp);   // VCFFileHeader.java:45-71,219-246
protected:
    DetectorOptions opt_;
    const Genome* genome_ = nullptr;
    void* ctx_ = nullptr;                        // context of the first device (ctxs_[0])
    std::vector<void*> ctxs_;                    // one context per device
    std::vector<std::vector<StrLocus>> strs_;    // -knownSTRs: loci per sequence (loaded with the genome)
    bool strs_loaded_ = false;
    std::vector<std::vector<KnownVariant>> known_;   // -knownVariants: per sequence
    bool known_loaded_ = false;
    std::thread warm_;                               // warm_up()'s thread
    QualHint qual_hint_;                             // the quality coding of the last tile packed (run_device, pack_sequence)
    int inner_threads_ = 1;                          // host threads one sequence may use for its own stages (set per run)
    bool known_mode() const { return known_loaded_ && n_known_ > 0; }   // SSVPL:147: an empty input list means discovery
    size_t n_known_ = 0;
    void emit_known_text(SeqJob& job);
    // the tallies of every sequence's span > 1 events in ONE device call, then the listener's sequential part per sequence
    int tally_span_events(std::vector<SeqJob>& jobs);
    void resolve_fired_events(SeqPlan& plan);
    bool deferred_tallies() const { return !opt_.host_only && !with_groups(); }
    std::vector<std::vector<bool>> ref_uploaded_;   // per device, per sequence
    RunStats stats_;
    std::string err_;
    static constexpr int kHostRing = 4;          // host tile buffers (page-locked, reused from run to run); up to 3 tiles in flight
    std::vector<std::vector<PackedTile>> rings_;   // per device: kHostRing tile buffers
    std::vector<ngsep_snv_call> call_buf_;       // where ngsep_tile_fetch writes
    virtual int ensure_ctx();
    virtual int create_ctx(int device, void** ctx);   // one device context with this detector's options
    virtual std::string header_text() const { return vcf_header(opt_.sample_id, opt_.ploidy, opt_.print_sample_ploidy); }
    virtual bool with_groups() const { return false; }
    // waits for the oldest tile in flight and appends what it returned to the job it belongs to
    virtual int fetch_tile(void* ctx, SeqJob& job, ngsep_tile_stats& ts);
    // ordered merge (embedded / indel suppression, SSVPL:150-160) + VCF text of one sequence (SSVD:933-968) -> job.text
    virtual void emit_text(SeqJob& job);
    // discoverVariantWithSpan at an event with a reference span > 1 that is not embedded (SSVPL:257-273); may run on any
    // host thread: everything it changes lives in `plan`
    virtual void on_span_event(int e, const std::string& ref_upper, const EventOutcome& ev, std::vector<HostRead*>& pileup, SeqPlan& plan, int& last_indel_end);
    int plan_sequence(const BatchView& b, int seq, int64_t lo, int64_t hi, SeqPlan& plan);
    void tile_records(const SeqPlan& plan, int tf, int tl, size_t& r_lo, size_t& r_hi) const;
    int tile_len() const { return opt_.tile_len > 0 ? opt_.tile_len : kAutoTileLen; }
    // packs and submits the tiles of the sequences next_job() hands out (an index into jobs, -1: no more), fetches what comes back
    int run_device(std::vector<SeqJob>& jobs, const std::function<long()>& next_job);
public:
    int pack_sequence(const BatchView& b, int seq, PackedTile& tile, int& tile_first, int& tile_last);
};

// Drop-in for MultisampleVariantsDetector.run (MSVD:421-462) over in-memory alignments of all samples (one coordinate-sorted
// stream: the single-file input, or the k-way merge of APG:212-289 done by the caller).
class GpuMultisampleVariantsDetector : public GpuSingleSampleVariantsDetector {
public:
    GpuMultisampleVariantsDetector(const DetectorOptions& o, std::vector<SampleDef> samples, int n_read_groups);
    static std::string vcf_header(const std::vector<SampleDef>& samples, bool psp);
    const std::vector<SampleDef>& samples() const { return samples_; }
protected:
    std::vector<SampleDef> samples_;
    int n_read_groups_;
    int create_ctx(int device, void** ctx) override;
    std::string header_text() const override { return vcf_header(samples_, opt_.print_sample_ploidy); }
    bool with_groups() const override { return true; }
    int fetch_tile(void* ctx, SeqJob& job, ngsep_tile_stats& ts) override;
    void emit_text(SeqJob& job) override;
    std::vector<ngsep_pop_site> sbuf_;
    std::vector<ngsep_sample_call> cbuf_;
    void on_span_event(int e, const std::string& ref_upper, const EventOutcome& ev, std::vector<HostRead*>& pileup, SeqPlan& plan, int& last_indel_end) override;
};

std::string pop_snv_vcf_line(const ngsep_pop_site& site, const ngsep_sample_call* calls, int n_samples, const std::string& chrom, int ploidy, bool embedded);
std::string pop_indel_vcf_line(const PopIndelRecord& r, const std::string& chrom, int ploidy);
std::string english_fmt(double v);   // main/io/ParseUtils.java:29, DecimalFormat("##0.0#")
// CalledGenomicVariantImpl.updateAllelesCopyNumberFromCounts (CalledGenomicVariantImpl.java:228-286)
std::vector<int> impl_copy_numbers(const std::vector<int>& called, const std::vector<int>& allele_counts, bool has_counts, int ploidy, size_t n_alleles);
void append_gt(std::string& s, const std::vector<int>& called, int ploidy);   // VCFFileWriter.java:171-199

// RelativeAlleleCountsCalculator.runProcess + printResults over a batch (allele_counts.cpp): the report text goes to `out`
// regions_to_filter / regions_to_select: the -frs / -srs region files ("" = none)
int run_relative_allele_counts(const Genome& genome, const BatchView& b, int min_rd, int max_rd, int min_base_qs, bool secondary, int device, std::string& out,
                               std::string& err, const std::string& regions_to_filter = std::string(), const std::string& regions_to_select = std::string());
// CoverageStatisticsCalculator.processFile + printCoverageStats over a batch (coverage.cpp): the table text goes to `out`
int run_coverage_stats(const Genome& genome, const BatchView& b, int min_mq, int device, std::string& out, std::string& err);

std::string snv_vcf_line(const ngsep_snv_call& c, const std::string& chrom, int ploidy, bool csb, bool embedded);
std::string indel_vcf_line(const IndelRecord& r, const std::string& chrom, int ploidy);
void host_tables(double het_rate, double* t95);   // CountsHelper.updateProbabilitiesCache, CH:135-187,414-415

}  // namespace host
}  // namespace ngsep_b200
