// =============================================================================
// TEST INFRASTRUCTURE ONLY — NOT PART OF THE PRODUCT PATH.
//
// CPU oracle: a C++17 restatement of the NGSEP 4.3.2 Java pileup / SNVQ path
// (acastem15/NGSEPcore).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may build, link, import or execute
// anything under oracle/.  The product library (ngsepcore_b200/csrc) never does.
//
// Parity status: the SNV genotype arithmetic (CountsHelper / SNVQ / VCF fields)
// is PINNED against the reference's own demo call set
// training/yeastDemo_ann_q40_s_fi_I2_noREP_noCNV.vcf.gz (tests/golden/
// yeast_demo_kat.tsv), and so is the population INFO arithmetic (NS, AN, AFS,
// MAF: tests/golden/yeast_demo_diversity.tsv).  The pileup walk, indel realigner, allele clustering and
// multi-sample logic have NO golden vectors in the reference ("parity unpinned"
// for those rows): fidelity rests on this line-by-line restatement.  No JVM
// exists in this environment, so the reference jar itself cannot be run.
//
// Every class/function cites the reference file:line it follows, relative to
// /root/reference/src/ngsep/.
// =============================================================================
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <ostream>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

namespace ngsep_oracle {

// ---- alignments/ReadAlignment.java:43-69 -----------------------------------
constexpr int FLAG_PAIRED = 0x1, FLAG_PROPER = 0x2, FLAG_READ_UNMAPPED = 0x4,
              FLAG_MATE_UNMAPPED = 0x8, FLAG_READ_REVERSE_STRAND = 0x10,
              FLAG_FIRST_OF_PAIR = 0x40, FLAG_SECOND_OF_PAIR = 0x80,
              FLAG_SECONDARY = 0x100, FLAG_FAILS_QC = 0x200, FLAG_PCR_DUP = 0x400,
              FLAG_SUPPLEMENTARY = 0x800, FLAG_MULTIPLE_ALN = 0x1000,
              FLAG_MATE_DIFFERENT_SEQUENCE = 0x2000;
constexpr int OP_H = 0, OP_D = 1, OP_I = 2, OP_M = 3, OP_P = 4, OP_N = 5, OP_S = 6, OP_X = 7;
constexpr int DEF_MIN_MQ_UNIQUE_ALIGNMENT = 20;

// ---- variants/GenomicVariant.java:30-59 -------------------------------------
constexpr int8_t TYPE_UNDETERMINED = 0, TYPE_BIALLELIC_SNV = 1, TYPE_MULTIALLELIC_SNV = 2,
                 TYPE_EMBEDDED_SNV = 3, TYPE_INDEL = 4, TYPE_STR = 5;
constexpr int MAX_NUM_ALLELES = 100;
constexpr int8_t INVALID_STRAND_BIAS_SCORE = -1;
constexpr int8_t MAX_STRAND_BIAS_SCORE = 100;

// Java semantics helpers
long java_round(double x);                 // Math.round(double)
short phred_score(double p);               // math/PhredScoreHelper.java:31-40
double phred_probability(short q);         // math/PhredScoreHelper.java:46-51
double log_sum10(double l1, double l2);    // math/LogMath.java:38-44
double fisher_pvalue(int a, int b, int c, int d);  // math/FisherExactTest.java:65-136
std::string hamming_consensus(const std::vector<std::string>& seqs);  // sequences/HammingSequenceDistanceMeasure.java:88-101
// iteration order of a java.util.HashMap<Integer,?> given insertion order of DISTINCT keys
std::vector<int> java_int_hashmap_order(const std::vector<int>& keys_in_insertion_order);

// ---- genome/ReferenceGenome.java + sequences/DNAMaskedSequence.java ---------
struct ReferenceGenome {
    std::vector<std::string> names;
    std::vector<std::string> seqs;  // alphabet AaCcNngGtT, everything else -> N (DNAMaskedSequence.java:33-59)
    std::unordered_map<std::string, int> index;
    void add(const std::string& name, const std::string& chars);
    static ReferenceGenome load_fasta(const std::string& path);
    int find(const std::string& name) const;
    // genome/ReferenceGenome.java:222-245: 1-based inclusive; ok=false when out of bounds/unknown
    bool get(const std::string& name, int first, int last, std::string& out) const;
    long total_length() const;
};

// variants/GenomicVariantImpl.java / variants/SNV.java (merged)
struct GenomicVariant {
    std::string seq;
    int first = 0, last = 0, len = 0;
    std::vector<std::string> alleles;
    int8_t type = TYPE_UNDETERMINED;
    short qs = 0;
    std::string id;          // VCF ID column ("" = none, GenomicVariantImpl.java:139-144)
    bool snv_class = false;  // instance of variants/SNV.java
    bool is_snv() const;     // GenomicVariantImpl.java:251-258 / SNV.java:193
    bool is_biallelic() const { return alleles.size() == 2; }
};

struct IndelCall { int first, last, length; };  // GenomicVariantImpl(seq,first,last,TYPE_INDEL)+setLength, RA:812-813

// ---- alignments/ReadAlignment.java ------------------------------------------
struct ReadAlignment {
    int seq_idx = -1;
    std::string seq_name;
    int first = 0, last = 0;
    int flags = 0;
    int mapq = 0;
    std::vector<int> aln;       // len*8+op, RA:99,1180-1198
    int read_length = 0;
    std::string name;
    std::string bases; bool has_bases = false;
    std::vector<uint8_t> quals; bool has_quals = false;   // phred+33 bytes, RA:581-595
    std::string read_group;     // "" default, RA:73,120
    int8_t ignore_close_to_indel = 2;  // RA:115
    short ignore_start = 0, ignore_end = 0;
    // quick retrieval, RA:103-106
    std::vector<short> allele_call_length;
    std::map<int, IndelCall> indel_calls; bool has_indel_calls = false;
    bool calls_updated = false;

    bool is_negative() const { return flags & FLAG_READ_REVERSE_STRAND; }
    bool is_secondary() const { return flags & FLAG_SECONDARY; }
    bool is_unique() const { return !(flags & FLAG_MULTIPLE_ALN); }
    void set_cigar_string(const std::string& cigar);        // RA:1222-1266
    void set_quality_scores(const std::string& qs);          // RA:581-595
    void set_ignore_5p(short v);                             // RA:613-626
    void set_ignore_3p(short v);                             // RA:632-644
    void set_ignore_start(short v);                          // RA:657-662
    void set_ignore_end(short v);                            // RA:675-680
    void update_allele_calls_info();                         // RA:747-834
    int aligned_read_position(int ref_pos) const;            // RA:842-871
    bool allele_call(int ref_pos, std::string& out);         // RA:989-1000
    bool allele_call(int ref_first, int ref_last, std::string& out);  // RA:1008-1016
    char base_quality(int ref_pos) const;                    // RA:1022-1027
    bool base_qualities(int ref_first, int ref_last, std::string& out) const;  // RA:1034-1041
    const std::map<int, IndelCall>* get_indel_calls();       // RA:1050-1054
    const IndelCall* get_indel_call(int ref_pos);            // RA:1101-1106
    bool has_indel_calls_in(int ref_first, int ref_last);    // RA:1471-1478
    bool move_indel_start(int indel_ref_pos, int new_ref_pos);  // RA:1114-1153
    void realign_start(int new_first, int first_match_len, int ref_pos_after, int aln_read_pos_after);  // RA:1372-1418
    void realign_end(int ref_pos_before, int aln_pos_before, int final_match_ref_start, int final_match_len);  // RA:1427-1469
    int soft_clip_start() const;                             // RA:1351-1356
    int soft_clip_end() const;                               // RA:1358-1363
    std::string cigar_string() const;                        // RA:1208-1217
};
using ReadPtr = std::shared_ptr<ReadAlignment>;

// One decoded SAM record (what htsjdk hands to alignments/io/ReadAlignmentFileReader.java:219-272)
struct SamRecord {
    std::string qname; int flag = 0; std::string rname; int pos = 0; int mapq = 0;
    std::string cigar; std::string rnext; int pnext = 0; int tlen = 0;
    std::string seq, qual;
    bool has_nh = false; int nh = 0; bool has_rg = false; std::string rg; bool has_nm = false; int nm = 0;
};
bool parse_sam_line(const std::string& line, SamRecord& rec);

// alignments/io/ReadAlignmentFileReader.java:219-354 restated over SamRecord
struct AlignmentReader {
    int filter_flags = 0, required_flags = 0, min_mq = DEF_MIN_MQ_UNIQUE_ALIGNMENT;
    const ReferenceGenome* genome = nullptr;
    bool have_last = false; SamRecord last;
    // returns nullptr when the record is dropped
    ReadPtr load(const SamRecord& rec);
};

// ---- discovery/PileupAlleleCall.java ----------------------------------------
struct PileupAlleleCall { std::string seq, quals, read_group; bool neg = false; std::string allele() const; };

// ---- discovery/PileupRecord.java ---------------------------------------------
struct PileupRecord {
    std::string seq; int pos = 0;
    std::vector<ReadAlignment*> alns;
    std::map<std::string, std::vector<ReadAlignment*>> by_rg;
    int span = 1, n_alns = 0, n_unique = 0, n_neg = 0;
    bool str = false, new_str = false, embedded = false;
    void add(ReadAlignment* a);                                   // PR:154-167
    std::vector<PileupAlleleCall> allele_calls(int span, const std::string* rg);  // PR:126-152
    std::vector<PileupAlleleCall> allele_calls(int span, const std::vector<std::string>* rgs);  // PR:104-111
    bool input_str() const { return str && !new_str; }
};

// ---- discovery/CountsHelper.java ---------------------------------------------
struct CountsHelper {
    int total = 0, low_q = 0;
    std::vector<int> counts;
    std::vector<std::array<int, 2>> strand;
    std::vector<std::vector<double>> logc;
    int8_t max_base_qs = 30;
    double het_prop = 0.5;
    std::vector<std::string> alleles;
    explicit CountsHelper(const std::vector<std::string>& alleles);
    CountsHelper();
    static CountsHelper snv(const std::vector<PileupAlleleCall>& calls, int8_t max_qs, double het_prop);   // CH:83-95
    static CountsHelper snv_gt(const std::vector<std::string>& alleles, const std::vector<PileupAlleleCall>& calls, int8_t max_qs, double het_prop);   // CH:86-95
    std::vector<double> posteriors_major(double het_rate, int major) const;        // CH:450-467
    static CountsHelper indel(const std::vector<std::string>& alleles, const std::vector<PileupAlleleCall>& calls, int8_t max_qs, double het_prop);  // CH:96-105
    void update(const std::string& allele, int8_t q, bool neg);                  // CH:209-251
    void update_indel(const std::string& call, const std::string& quals, bool neg);  // CH:253-304
    std::vector<std::vector<double>> posteriors(double het_rate) const;           // CH:410-443
    static void posterior_from_events(std::vector<double>& ev);                   // CH:472-495
    int8_t strand_bias_score(int i1, int i2) const;                               // CH:563-576
    // table access for tests / product table upload (CH:135-187)
    static double table_error(int q, int j);
    static double table_gt(int f, int q, int j);
    static double table_freq(int f, int k);
};

// variants/CalledSNV.java + variants/CalledGenomicVariantImpl.java (merged)
struct CalledVariant {
    GenomicVariant var;
    bool called_snv = false;            // instance of CalledSNV
    std::string sample_id;
    // CalledSNV state (CalledSNV.java:33-58)
    int8_t genotype = -1;
    int cnt[4] = {0, 0, 0, 0};
    float homo_ref = 0, homo_alt = 0, ref_alt = 0, alt_ref = 0;
    short ref_cn = 0;
    // CalledGenomicVariantImpl state (CalledGenomicVariantImpl.java:30-42)
    std::vector<int8_t> idx_called;
    std::vector<short> alleles_cn;
    bool has_report = false; std::vector<std::string> rep_alleles; std::vector<int> rep_counts; bool rep_has_counts = false;
    std::vector<std::vector<double>> rep_logc; bool rep_has_logc = false;
    bool has_all_counts = false; std::vector<int> all_counts;
    // shared
    int depth = 0; short gq = 0; short total_cn = 2; int8_t strand_bias = INVALID_STRAND_BIAS_SCORE;

    static CalledVariant make_snv(const GenomicVariant& snv, int8_t genotype);                // CalledSNV.java:78-82
    static CalledVariant make_impl(const GenomicVariant& v, const std::vector<int8_t>& idx);  // CalledGenomicVariantImpl.java:50-55
    bool undecided() const; bool hom_ref() const; bool homozygous() const; bool heterozygous() const;
    std::vector<int8_t> called_indexes() const;
    void set_all_base_counts(const std::vector<int>& c);          // CalledSNV.java:245-252
    void set_all_logc(const std::vector<std::vector<double>>& l); // CalledSNV.java:258-265
    void update_cn_from_counts(short total);                      // CalledSNV.java:134-160 / Impl:228-286
    void make_undecided();
    int report_count(const std::string& allele) const;            // VariantCallReport.java:22-27
    double report_logc(const std::string& a1, const std::string& a2) const;  // VariantCallReport.java:28-34
    int last() const { return var.last; }
};

// ---- discovery/VariantDiscoverySNVQAlgorithm.java ---------------------------
bool discover_snv(const CountsHelper& h, const std::string& seq, int pos, char ref, double het_rate, bool csb, CalledVariant& out);  // VD:100-222
CalledVariant genotype_snv(const GenomicVariant& v, const CountsHelper& h, double het_rate, bool csb);  // VD:21-97
bool call_indel(PileupRecord& p, const CountsHelper& h, const GenomicVariant* v, double het_rate, bool csb, CalledVariant& out);  // VD:265-361

// ---- discovery/AlleleCallClustersBuilder.java --------------------------------
std::vector<std::string> cluster_allele_calls(const std::vector<PileupAlleleCall>& calls, const std::string& reference, int8_t max_qs);  // ACCB:72-141

// ---- discovery/PileupListener.java:27-31 -------------------------------------
struct PileupListener {
    virtual ~PileupListener() {}
    virtual void on_pileup(PileupRecord& p) = 0;
    virtual void on_sequence_start(const std::string& seq) = 0;
    virtual void on_sequence_end(const std::string& seq) = 0;
};

// ---- discovery/IndelRealignerPileupListener.java ------------------------------
// An input variant of the realigner (IRPL:52-53, 62-66): with -knownSTRs the merged, padded tandem-repeat loci of
// SingleSampleVariantsDetector.makeNonRedundantSTRs (SSVD:843-894), GenomicVariantImpl of TYPE_STR with the reference allele only
struct InputVariant { int first = 0, last = 0; bool str = false; };
// -knownVariants: vcf/VCFFileReader.loadVariants(file, true, true) (VCFFileReader.java:585-600 over :152-306, load mode MINIMAL)
std::vector<GenomicVariant> load_variants(const std::string& path);
// the per-sequence, position-sorted view GenomicRegionSortedCollection.getSequenceRegions(name).asList() gives of them
// (genome/GenomicRegionSortedCollection.java:224-232, GenomicRegionPositionComparator.java:34-37: first, then last, stable)
std::map<std::string, std::vector<GenomicVariant>> sort_variants_by_sequence(const std::vector<GenomicVariant>& vars);
struct GenomicRegionRow { std::string seq; int first = 0, last = 0; };
std::vector<GenomicRegionRow> load_regions(const std::string& path);   // genome/io/SimpleGenomicRegionFileHandler.java:57-80
std::map<std::string, std::vector<InputVariant>> make_non_redundant_strs(const ReferenceGenome& genome, const std::vector<GenomicRegionRow>& strs);  // SSVD:843-894

struct IndelRealignerPileupListener : PileupListener {
    const ReferenceGenome* genome = nullptr;
    std::map<std::string, std::vector<InputVariant>> input_variants;   // per sequence, sorted (IRPL:129-134)
    int min_bp_good_ref_aln = 5, max_bp_realign_end = 50;   // IRPL:46-47
    double min_prop_alns_realign = 0;                       // IRPL:48
    long n_events = 0;
    void on_pileup(PileupRecord& p) override;               // IRPL:84-126
    void on_sequence_start(const std::string& seq) override {   // IRPL:129-134
        auto it = input_variants.find(seq);
        seq_variants = it == input_variants.end() ? nullptr : &it->second;
        idx_next_variant = 0;
    }
    void on_sequence_end(const std::string&) override {}
private:
    const std::vector<InputVariant>* seq_variants = nullptr;
    size_t idx_next_variant = 0;
    const InputVariant* intersect_with_variants(int pos);   // IRPL:141-157
    int conciliate(PileupRecord& p, std::vector<ReadAlignment*>& alns, int event_end, bool fixed_event);       // IRPL:165-216
    int analyze(std::vector<ReadAlignment*>& alns, int ev_start, int ev_end, std::set<int>& lengths, std::vector<ReadAlignment*>& indel_alns, std::vector<int>& votes);  // IRPL:229-265
    int move_starts(std::vector<ReadAlignment*>& alns, int first, int last, int max_len, int offset);  // IRPL:274-313
    int look_for_new_str(PileupRecord& p, std::vector<ReadAlignment*>& alns, int max_len);  // IRPL:315-349
    void process_ends(std::vector<ReadAlignment*>& alns, const std::string& seq, int ev_first, int ev_last);  // IRPL:400-526
};
int check_mononucleotide(const std::string& seq);  // IRPL:365-391

struct Sample { std::string id; short ploidy = 2; std::vector<std::string> read_groups; };

// ---- discovery/SingleSampleVariantPileupListener.java -------------------------
struct SingleSampleVariantPileupListener : PileupListener {
    const ReferenceGenome* genome = nullptr;
    double het_rate = 0.001; bool csb = false; int8_t max_base_qs = 30;
    bool ignore_lower = false, call_embedded = false; short min_quality = 40;
    Sample sample{"Sample", 2, {}};
    std::vector<CalledVariant> called;
    int last_indel_end = 0;
    // -knownVariants (SSVPL:77, 135-137): non-empty = genotype these variants instead of discovering
    std::map<std::string, std::vector<GenomicVariant>> input_variants;
    size_t n_input_variants = 0;
    // optional per-position count dump (test hook; not in the reference): pos -> A,C,G,T,DP
    std::vector<std::array<int, 6>>* counts_dump = nullptr;
    void on_pileup(PileupRecord& p) override;                    // SSVPL:146-177
    void on_sequence_start(const std::string& seq) override {    // SSVPL:183-187
        auto it = input_variants.find(seq);
        seq_input_variants = it == input_variants.end() ? nullptr : &it->second;
        next_siv_index = 0;
        last_indel_end = 0;
    }
    void on_sequence_end(const std::string&) override {}
private:
    const std::vector<GenomicVariant>* seq_input_variants = nullptr;
    size_t next_siv_index = 0;
public:
    static bool reference_allele_discovery(PileupRecord& p, const ReferenceGenome& g, bool call_embedded, bool ignore_lower, std::string& out);  // SSVPL:191-206
    bool discover_variant(PileupRecord& p, const std::string& ref_allele, CalledVariant& out);   // SSVPL:213-232
    bool discover_snv_at(PileupRecord& p, char ref, CalledVariant& out);                         // SSVPL:234-255
    bool discover_with_span(PileupRecord& p, const std::string& ref_allele, CalledVariant& out); // SSVPL:257-273
    CalledVariant genotype_variant_sample(const GenomicVariant& v, PileupRecord& p, const Sample* s, double h);  // SSVPL:361-391
    // pools / polyploids (ploidy >= DEF_MIN_PLOIDY_POOL_ALGORITHM = 3)
    CalledVariant genotype_variant_pool(const GenomicVariant& v, short haplotypes, const std::vector<PileupAlleleCall>& calls, double h);   // SSVPL:402-503
    bool discover_indel(PileupRecord& p, const std::vector<std::string>& alleles, const std::vector<PileupAlleleCall>& calls, CalledVariant& out);   // SSVPL:275-296
};

// ---- discovery/AlignmentsPileupGenerator.java ---------------------------------
struct AlignmentsPileupGenerator {
    std::vector<PileupListener*> listeners;
    std::string query_seq; bool has_query = false; int query_first = 0, query_last = 1000000000;
    int max_alns_per_start = 5; bool non_unique_primary = false, secondary = false;
    int8_t ignore5 = 0, ignore3 = 0; int min_mq = DEF_MIN_MQ_UNIQUE_ALIGNMENT;
    long n_pileups = 0, n_nonempty = 0;
    int filter_flags() const;                 // APG:367-371
    void process_alignment(const ReadPtr& a); // APG:377-403
    void notify_end();                        // APG:447-452
    // query-window filter of APG:310-322; returns 0 = process, 1 = skip, 2 = stop
    int query_filter(const ReadAlignment& a);
private:
    bool have_seq = false; std::string cur_seq; int cur_pos = 0, cur_last = 0; bool query_found = false;
    std::vector<ReadPtr> pending, same_primary, same_secondary; int last_start = 0;
    void same_start_alns();                   // APG:407-433
    void process_pileups(int start);          // APG:453-462
    void update_pending();                    // APG:464-471
    bool process_current();                   // APG:475-498
};

// ---- vcf/VCFFileHeader.java + vcf/VCFFileWriter.java --------------------------
void print_vcf_header(std::ostream& out, const std::vector<Sample>& samples, bool print_sample_ploidy);  // VCFFileHeader.java:45-71,219-246
void print_vcf_record_single(std::ostream& out, CalledVariant& call, short ploidy);  // SSVD:944-963 + VCFFileWriter.java:44-70

// ---- discovery/SingleSampleVariantsDetector.java (findSNVs part) --------------
struct DetectorOptions {
    std::string sample_id = "Sample"; short ploidy = 2; bool psp = false; int min_mq = 20;
    std::string query_seq; bool has_query = false; int query_first = 0, query_last = 1000000000;
    bool ignore_lower = false; int max_alns_per_start = 5; bool non_unique_primary = false, secondary = false;
    int ignore5 = 0, ignore3 = 0; double het_rate = 0.001; bool het_modified = false; int max_base_qs = 30;
    short min_quality = 40; bool call_embedded = false, csb = false;
    std::string known_strs;             // -knownSTRs (SSVD:906-913)
    std::string known_variants;         // -knownVariants (SSVD:897-905)
    double min_allele_depth_freq = 0;   // MSVD:58 (setter only, no CLI option id; oracle switch --minAlleleDepthFrequency)
};
struct SingleSampleVariantsDetector : PileupListener {
    DetectorOptions opt;
    const ReferenceGenome* genome = nullptr;
    std::ostream* out = nullptr;
    AlignmentsPileupGenerator gen;
    IndelRealignerPileupListener realigner;
    SingleSampleVariantPileupListener var_listener;
    long covered = 0;
    void begin();                                  // SSVD:589-593,896-928 (header + listener chain)
    void add_record(const SamRecord& rec);         // APG.processFile loop body, APG:307-327
    void finish();                                 // APG:330
    void on_pileup(PileupRecord&) override { covered++; }         // SSVD:1011-1018
    void on_sequence_start(const std::string&) override {}
    void on_sequence_end(const std::string& seq) override;        // SSVD:933-968,1026-1032
private:
    AlignmentReader reader; bool stopped = false;
};

// ---- discovery/MultisampleVariantsDetector.java (oracle/ngsep_oracle_multi.cpp) ----------------------------------
std::vector<std::string> java_string_hashset_order(const std::vector<std::string>& keys_in_insertion_order);
bool load_samples_from_headers(const std::vector<std::vector<std::string>>& header_lines_per_file, short ploidy, std::vector<Sample>& out, std::string& err);  // MSVD:492-516
GenomicVariant make_new_variant(const GenomicVariant& v, const std::vector<std::string>& new_alleles);   // SSVPL:346-359 (alleles in iteration order)
bool alleles_same_length(const std::vector<std::string>& alleles);   // SSVPL:339-345
bool create_snv_variant_pool(const PileupRecord& p, const CountsHelper& h, char reference, double min_allele_depth_freq, GenomicVariant& out);  // SSVPL:297-332
std::string english_fmt(double v);   // main/io/ParseUtils.java:29
std::string diversity_info(size_t n_alleles, const std::vector<std::vector<int8_t>>& called_idx, const std::vector<std::vector<short>>& copy_numbers);  // VCFRecord.java:288-301
void print_genotype_field(std::ostream& out, CalledVariant& call, bool snv_format, short ploidy);  // VCFFileWriter.java:159-308
void print_vcf_record_population(std::ostream& out, const GenomicVariant& v, std::vector<CalledVariant>& calls, const std::vector<Sample>& samples);  // VCFRecord.java:277-301

struct MultisampleVariantsDetector : PileupListener {
    DetectorOptions opt;                       // MSVD option set (XML:480-563); sample_id / csb unused
    double min_allele_depth_freq = 0;          // MSVD:58 (no CLI option)
    const ReferenceGenome* genome = nullptr;
    std::ostream* out = nullptr;
    std::vector<Sample> samples;               // MSVD:492-516 order
    AlignmentsPileupGenerator gen;
    IndelRealignerPileupListener realigner;
    int last_indel_end = 0;
    long n_records = 0, n_input_records = 0;
    std::vector<std::array<int, 6>>* counts_dump = nullptr;   // test hook: pooled pos,A,C,G,T,DP wherever discoverPopulationSNV ran
    void begin();                                          // MSVD:421-456
    void add_alignment(const ReadPtr& a);                  // APG:243-262
    void finish();
    void process_files(std::vector<std::istream*>& sams);  // APG:212-289
    void on_pileup(PileupRecord& p) override;              // MSVD:521-537
    void on_sequence_start(const std::string&) override { last_indel_end = 0; }   // MSVD:561-565
    void on_sequence_end(const std::string&) override {}
    bool discover_population_snv(PileupRecord& p, char reference, GenomicVariant& out);                   // MSVD:585-597
    bool discover_population_with_span(PileupRecord& p, const std::string& ref_allele, GenomicVariant& out);  // MSVD:599-618
    bool discover_population_indel(PileupRecord& p, const CountsHelper& h, GenomicVariant& out);          // MSVD:620-634
    std::vector<CalledVariant> genotype_variant(GenomicVariant& v, PileupRecord& p, double h);            // MSVD:674-693
    bool make_new_variant(GenomicVariant& v, const std::vector<CalledVariant>& calls);                    // MSVD:642-656
private:
    bool stopped = false;
};

}  // namespace ngsep_oracle
