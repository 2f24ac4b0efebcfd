// Host pipeline above the C ABI: event pre-pass (realigner + host indel calls), tile packing, device calls,
// ordered merge and VCF text.  Mirrors the public surface of the reference's detector
// (src/ngsep/discovery/SingleSampleVariantsDetector.java: setters of :141-560, run() :589-656) and of
// AlignmentsPileupGenerator (APG:71-193) for this path only.
#pragma once
#include <cstdint>
#include <functional>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../../../include/ngsep_pileup.h"
#include "model.hpp"

namespace ngsep_b200 {
namespace host {

struct EventOutcome { bool fired = false; int span = 1; bool str = false, new_str = false; };

struct Realigner {
    const Genome* genome = nullptr;
    int seq = 0;
    EventOutcome on_event(int pos, std::vector<HostRead*>& pileup);            // IRPL:84-216
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
};

struct IndelCallerOptions { int max_base_qs = 30; double het_rate = 0.001; int min_quality = 40; bool csb = false; };

// SingleSampleVariantPileupListener.discoverVariantWithSpan (SSVPL:257-273) up to the SNV fall-back:
// AlleleCallClustersBuilder + CountsHelper.calculateCountsIndel + VariantDiscoverySNVQAlgorithm.callIndel.
bool call_indel_at(int pos, const std::string& ref_allele, bool is_str, std::vector<HostRead*>& pileup, const IndelCallerOptions& opt, IndelRecord& out);

// Phred / rounding helpers shared by the host writer (math/PhredScoreHelper.java:31-40)
long jround(double x);
int phred(double p);
double fisher_pvalue(int a, int b, int c, int d);   // math/FisherExactTest.java:65-136

struct PackedTile {
    std::vector<ngsep_read_hdr> hdr;
    std::vector<uint32_t> payload;
    std::vector<uint8_t> qual;   // phred+33 bytes, one run per read (padded to 4), addressed by the records' qoff
    int max_span = 1, max_back = 0;
    uint64_t aligned_bases = 0;
    ngsep_tile view(int seq_idx, int first, int last) const;
};
// Packs reads[lo,hi) (all their versions) in pileup order.
void pack_reads(const std::vector<HostRead>& reads, size_t lo, size_t hi, PackedTile& out);

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
    // device plumbing (not reference options)
    int device = 0;
    int tile_len = 0;          // 0 = one tile per sequence; otherwise region tiles of this many positions
    int shard = 0, n_shards = 1;   // region sharding across ranks: this process handles tiles t % n_shards == shard
    bool host_only = false;    // test hook: run admission + event pre-pass only and report the host-made indel/STR records
                               // (no SNV calling happens anywhere in this mode; it is not a fallback)
};

struct RunStats {
    uint64_t records = 0, admitted = 0, versions = 0, events = 0, indel_calls = 0, snv_calls = 0;
    uint64_t covered_positions = 0, candidate_positions = 0, aligned_bases = 0, tiles = 0, kernel_launches = 0;
    double kernel_ms = 0, prepass_s = 0, pack_s = 0, device_s = 0, total_s = 0;
};

// Outcome of the host stages for one sequence (admission + event pre-pass).
struct SeqPlan {
    std::vector<HostRead> reads;                 // pileup order, with versions
    std::vector<IndelRecord> indels;             // host-made indel / STR calls
    std::set<int> suppressed;                    // positions whose device SNV decision must not be reported
    std::vector<std::pair<int, int>> embedded;   // closed intervals of embedded positions (SSVPL:150-151)
    int win_first = 1, win_last = 0, max_reach = 1, max_back = 0;
};

// Drop-in for SingleSampleVariantsDetector.findSNVs over in-memory alignments.
class GpuSingleSampleVariantsDetector {
public:
    explicit GpuSingleSampleVariantsDetector(const DetectorOptions& o);
    ~GpuSingleSampleVariantsDetector();
    void set_genome(const Genome* g);                 // setGenome, SSVD:161-163
    // Processes a coordinate-sorted batch; appends VCF text (header included when with_header) to out.
    int run(const BatchView& batch, std::string& out, bool with_header = true);
    const RunStats& stats() const { return stats_; }
    const std::string& error() const { return err_; }
    static std::string vcf_header(const std::string& sample_id, int ploidy, bool psp);   // VCFFileHeader.java:45-71,219-246
private:
    DetectorOptions opt_;
    const Genome* genome_ = nullptr;
    void* ctx_ = nullptr;
    std::vector<bool> ref_uploaded_;
    RunStats stats_;
    std::string err_;
    int ensure_ctx();
    int run_sequence(const BatchView& b, int seq, int64_t lo, int64_t hi, std::string& out);
    int plan_sequence(const BatchView& b, int seq, int64_t lo, int64_t hi, SeqPlan& plan);
    void tile_records(const SeqPlan& plan, int tf, int tl, size_t& r_lo, size_t& r_hi) const;
public:
    int pack_sequence(const BatchView& b, int seq, PackedTile& tile, int& tile_first, int& tile_last);
};

std::string snv_vcf_line(const ngsep_snv_call& c, const std::string& chrom, int ploidy, bool csb, bool embedded);
std::string indel_vcf_line(const IndelRecord& r, const std::string& chrom, int ploidy);
void host_tables(double het_rate, double* t95);   // CountsHelper.updateProbabilitiesCache, CH:135-187,414-415

}  // namespace host
}  // namespace ngsep_b200
