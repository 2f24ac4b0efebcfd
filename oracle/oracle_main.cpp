// TEST INFRASTRUCTURE ONLY — see ngsep_oracle.hpp.
// CLI + C API around the oracle.  The CLI accepts the option ids of the reference module
// (main/CommandsDescriptor.xml:565-703, SingleSampleVariantsDetector) that affect the SNV/indel path:
//   ngsep_oracle SingleSampleVariantsDetector -i in.sam -r ref.fa -o prefix [options]
// Extra, oracle-only switches (not in the reference): --dump-counts FILE, --stats FILE.
#include "ngsep_oracle.hpp"

#include <chrono>
#include <cmath>
#include <map>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>

using namespace ngsep_oracle;

namespace {

struct RunResult { long n_records = 0, n_pileups = 0, n_nonempty = 0, n_events = 0; double seconds = 0; };

RunResult run_ssvd(const ReferenceGenome& genome, std::istream& sam, std::ostream& vcf, const DetectorOptions& opt,
                   std::vector<std::array<int, 6>>* dump) {
    SingleSampleVariantsDetector det;
    det.opt = opt;
    det.genome = &genome;
    det.out = &vcf;
    det.var_listener.counts_dump = dump;
    auto t0 = std::chrono::steady_clock::now();
    det.begin();
    std::string line;
    SamRecord rec;
    RunResult r;
    while (std::getline(sam, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (!parse_sam_line(line, rec)) continue;
        r.n_records++;
        det.add_record(rec);
    }
    det.finish();
    auto t1 = std::chrono::steady_clock::now();
    r.seconds = std::chrono::duration<double>(t1 - t0).count();
    r.n_pileups = det.gen.n_pileups;
    r.n_nonempty = det.gen.n_nonempty;
    r.n_events = det.realigner.n_events;
    return r;
}

struct MsvdResult { long n_records = 0, n_pileups = 0, n_nonempty = 0, n_events = 0, n_vcf = 0; double seconds = 0; };

// MultisampleVariantsDetector.run (MSVD:421-462) over SAM text files: samples from the @RG headers, k-way merged streams.
bool run_msvd(const ReferenceGenome& genome, const std::vector<std::string>& sam_paths, std::ostream& vcf, const DetectorOptions& opt,
              std::vector<std::array<int, 6>>* dump, MsvdResult& r, std::string& err) {
    std::vector<std::vector<std::string>> headers;
    for (auto& path : sam_paths) {
        std::ifstream f(path);
        if (!f) { err = "cannot open " + path; return false; }
        std::vector<std::string> lines;
        std::string line;
        while (std::getline(f, line)) { if (line.empty() || line[0] != '@') break; if (line.back() == '\r') line.pop_back(); lines.push_back(line); }
        headers.push_back(lines);
    }
    MultisampleVariantsDetector det;
    det.opt = opt;
    det.genome = &genome;
    det.out = &vcf;
    det.counts_dump = dump;
    det.min_allele_depth_freq = opt.min_allele_depth_freq;
    if (!load_samples_from_headers(headers, opt.ploidy, det.samples, err)) return false;
    std::vector<std::unique_ptr<std::ifstream>> files;
    std::vector<std::istream*> streams;
    for (auto& path : sam_paths) { files.emplace_back(new std::ifstream(path)); streams.push_back(files.back().get()); }
    auto t0 = std::chrono::steady_clock::now();
    det.begin();
    det.process_files(streams);
    auto t1 = std::chrono::steady_clock::now();
    r.seconds = std::chrono::duration<double>(t1 - t0).count();
    r.n_records = det.n_input_records; r.n_pileups = det.gen.n_pileups; r.n_nonempty = det.gen.n_nonempty; r.n_events = det.realigner.n_events;
    r.n_vcf = det.n_records;
    return true;
}

int parse_options(int argc, char** argv, int start, DetectorOptions& opt, std::string& in, std::string& ref, std::string& out,
                  std::string& dump, std::string& stats) {
    for (int i = start; i < argc; i++) {
        std::string a = argv[i];
        auto need = [&](const char* what) -> const char* {
            if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", what); std::exit(1); }
            return argv[++i];
        };
        if (a == "-i") in = need("-i");
        else if (a == "-r") ref = need("-r");
        else if (a == "-o") out = need("-o");
        else if (a == "-sampleId") opt.sample_id = need("-sampleId");
        else if (a == "-ploidy") opt.ploidy = (short)std::atoi(need("-ploidy"));
        else if (a == "-psp") opt.psp = true;
        else if (a == "-minMQ") opt.min_mq = std::atoi(need("-minMQ"));
        else if (a == "-querySeq") { opt.query_seq = need("-querySeq"); opt.has_query = true; }
        else if (a == "-first") opt.query_first = std::atoi(need("-first"));
        else if (a == "-last") opt.query_last = std::atoi(need("-last"));
        else if (a == "-ignoreLowerCaseRef") opt.ignore_lower = true;
        else if (a == "-maxAlnsPerStartPos") opt.max_alns_per_start = std::atoi(need("-maxAlnsPerStartPos"));
        else if (a == "-p") opt.non_unique_primary = true;
        else if (a == "-s") opt.secondary = true;
        else if (a == "-ignore5") opt.ignore5 = std::atoi(need("-ignore5"));
        else if (a == "-ignore3") opt.ignore3 = std::atoi(need("-ignore3"));
        else if (a == "-h") { opt.het_rate = std::atof(need("-h")); opt.het_modified = true; }
        else if (a == "-maxBaseQS") opt.max_base_qs = std::atoi(need("-maxBaseQS"));
        else if (a == "-minQuality") opt.min_quality = (short)std::atoi(need("-minQuality"));
        else if (a == "-embeddedSNVs") opt.call_embedded = true;
        else if (a == "-csb") opt.csb = true;
        else if (a == "-knownSTRs") opt.known_strs = need("-knownSTRs");
        else if (a == "-knownVariants") opt.known_variants = need("-knownVariants");
        else if (a == "--minAlleleDepthFrequency") opt.min_allele_depth_freq = std::atof(need("--minAlleleDepthFrequency"));
        else if (a == "--dump-counts") dump = need("--dump-counts");
        else if (a == "--stats") stats = need("--stats");
        else { std::fprintf(stderr, "Unrecognized option %s\n", a.c_str()); return 1; }  // CommandsDescriptor.java:441-445
    }
    return 0;
}

}  // namespace

#ifndef NGSEP_ORACLE_NO_MAIN
int main(int argc, char** argv) {
    if (argc >= 2 && std::string(argv[1]) == "MultisampleVariantsDetector") {   // MSVD.main, MSVD:412-419: options, then the input files
        DetectorOptions opt;
        std::string in, ref, out = "variants.vcf", dump, stats;
        int i = 2;
        std::vector<char*> oargs;
        static const char* flags[] = {"-psp", "-ignoreLowerCaseRef", "-p", "-s", "-embeddedSNVs"};
        while (i < argc && argv[i][0] == '-') {
            bool is_flag = false;
            for (auto f : flags) if (!std::strcmp(f, argv[i])) is_flag = true;
            oargs.push_back(argv[i++]);
            if (!is_flag && i < argc) oargs.push_back(argv[i++]);
        }
        if (parse_options((int)oargs.size(), oargs.data(), 0, opt, in, ref, out, dump, stats)) return 1;
        std::vector<std::string> files;
        for (; i < argc; i++) files.push_back(argv[i]);
        if (ref.empty()) { std::fprintf(stderr, "The reference genome file is a required parameter\n"); return 2; }   // MSVD:423
        ReferenceGenome genome = ReferenceGenome::load_fasta(ref);
        std::ofstream vcf(out);   // MSVD:452
        MsvdResult r; std::string err;
        if (!run_msvd(genome, files, vcf, opt, nullptr, r, err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 2; }
        if (!stats.empty()) {
            std::ofstream s(stats);
            s << "{\"records\": " << r.n_records << ", \"pileups\": " << r.n_pileups << ", \"genotyped_positions\": " << r.n_nonempty
              << ", \"indel_events\": " << r.n_events << ", \"vcf_records\": " << r.n_vcf << ", \"seconds\": " << r.seconds << "}\n";
        }
        return 0;
    }
    if (argc < 2 || std::string(argv[1]) != "SingleSampleVariantsDetector") {
        std::fprintf(stderr, "usage: ngsep_oracle SingleSampleVariantsDetector -i in.sam -r ref.fa -o prefix [options]\n"
                             "       ngsep_oracle MultisampleVariantsDetector -r ref.fa -o out.vcf [options] a.sam b.sam ...\n");
        return 1;
    }
    DetectorOptions opt;
    std::string in, ref, out, dump, stats;
    if (parse_options(argc, argv, 2, opt, in, ref, out, dump, stats)) return 1;
    if (in.empty()) { std::fprintf(stderr, "The input file with alignments is a required parameter\n"); return 2; }   // SSVD:595
    if (ref.empty()) { std::fprintf(stderr, "The reference genome file is a required parameter\n"); return 2; }      // SSVD:596
    ReferenceGenome genome = ReferenceGenome::load_fasta(ref);
    std::ifstream sam(in);
    if (!sam) { std::fprintf(stderr, "cannot open %s\n", in.c_str()); return 2; }
    std::ofstream vcf(out + ".vcf");   // SSVD:926
    std::vector<std::array<int, 6>> counts;
    RunResult r = run_ssvd(genome, sam, vcf, opt, dump.empty() ? nullptr : &counts);
    vcf.close();
    if (!dump.empty()) {
        std::ofstream d(dump, std::ios::binary);
        d.write((const char*)counts.data(), (std::streamsize)(counts.size() * sizeof(counts[0])));
    }
    if (!stats.empty()) {
        std::ofstream s(stats);
        s << "{\"records\": " << r.n_records << ", \"pileups\": " << r.n_pileups << ", \"genotyped_positions\": " << r.n_nonempty
          << ", \"indel_events\": " << r.n_events << ", \"seconds\": " << r.seconds << "}\n";
    }
    return 0;
}
#endif

// ---------------------------------------------------------------------------------------------
// C API (ctypes) — function-level access for the known-answer tests
// ---------------------------------------------------------------------------------------------
extern "C" {

// Evaluate the SNV path on a synthetic column: counts[b] calls of base b ("ACGT") all with phred+33 quality
// char qual_char, plus n_lowq extra reference-base calls at quality '#' (q2).  discovery=1: VD.discoverSNV;
// discovery=0: VD.genotypeSNV for the biallelic SNV ref>alt.  Outputs: gt[2] (allele indexes into {ref,alt}; -1 none),
// pl[3], gq, qual, dp, called (0 = no call object).
int oracle_snv_column(const int* counts, int qual_char, int n_lowq, char ref, char alt, double het_rate, int max_base_qs,
                      int discovery, int* gt, int* pl, int* gq, int* qual, int* dp, char* alt_out) {
    std::vector<PileupAlleleCall> calls;
    for (int b = 0; b < 4; b++) for (int k = 0; k < counts[b]; k++) {
        PileupAlleleCall c; c.seq = std::string(1, "ACGT"[b]); c.quals = std::string(1, (char)qual_char); calls.push_back(c);
    }
    for (int k = 0; k < n_lowq; k++) { PileupAlleleCall c; c.seq = std::string(1, ref); c.quals = "#"; calls.push_back(c); }
    CountsHelper h = CountsHelper::snv(calls, (int8_t)max_base_qs, 0.5);
    CalledVariant cv;
    gt[0] = gt[1] = -1;
    if (discovery) {
        if (!discover_snv(h, "seq", 1, ref, het_rate, false, cv)) return 0;
    } else {
        GenomicVariant v; v.seq = "seq"; v.first = v.last = 1; v.len = 1; v.alleles = {std::string(1, ref), std::string(1, alt)};
        v.type = TYPE_BIALLELIC_SNV; v.snv_class = true;
        cv = genotype_snv(v, h, het_rate, false);
    }
    auto idx = cv.called_indexes();
    if (idx.size() == 1) gt[0] = gt[1] = idx[0];
    else if (idx.size() == 2) { gt[0] = idx[0]; gt[1] = idx[1]; }
    const auto& al = cv.var.alleles;
    int k = 0;
    for (size_t j = 0; j < al.size() && j < 2; j++) for (size_t i = 0; i <= j; i++) pl[k++] = (int)java_round(-10 * cv.report_logc(al[i], al[j]));
    *gq = cv.gq; *qual = cv.var.qs; *dp = cv.depth;
    if (alt_out) *alt_out = al.size() > 1 ? al[1][0] : '.';
    return 1;
}

double oracle_table_error(int q, int j) { return CountsHelper::table_error(q, j); }
double oracle_table_gt(int f, int q, int j) { return CountsHelper::table_gt(f, q, j); }
double oracle_fisher(int a, int b, int c, int d) { return fisher_pvalue(a, b, c, d); }
int oracle_phred(double p) { return phred_score(p); }
int oracle_mononucleotide(const char* s) { return check_mononucleotide(s); }

// Whole-path run on in-memory text.  Returns malloc'ed VCF text (caller frees with oracle_free).
// counts_out (optional): malloc'ed int32 [n][6] = pos,A,C,G,T,DP for every position where discoverSNV ran.
char* oracle_run_ssvd(const char* fasta_path, const char* sam_path, int argc, char** argv, int** counts_out, long* n_counts,
                      long* genotyped_positions, double* seconds) {
    DetectorOptions opt;
    std::string in, ref, out, dump, stats;
    if (parse_options(argc, argv, 0, opt, in, ref, out, dump, stats)) return nullptr;
    ReferenceGenome genome = ReferenceGenome::load_fasta(fasta_path);
    std::ifstream sam(sam_path);
    if (!sam) return nullptr;
    std::ostringstream vcf;
    std::vector<std::array<int, 6>> counts;
    RunResult r = run_ssvd(genome, sam, vcf, opt, counts_out ? &counts : nullptr);
    if (counts_out) {
        *counts_out = (int*)std::malloc(counts.size() * 6 * sizeof(int) + 8);
        std::memcpy(*counts_out, counts.data(), counts.size() * 6 * sizeof(int));
        *n_counts = (long)counts.size();
    }
    if (genotyped_positions) *genotyped_positions = r.n_nonempty;
    if (seconds) *seconds = r.seconds;
    std::string s = vcf.str();
    char* res = (char*)std::malloc(s.size() + 1);
    std::memcpy(res, s.c_str(), s.size() + 1);
    return res;
}


// ---- discovery/CoverageStatisticsCalculator.java -------------------------------------------------------------------------------
// The listener (processPileup :186-199, printCoverageStats :250-256) on the generator processFile configures for it (:130-138):
// secondary alignments processed (so only unmapped records are filtered, APG:367-371), 100 alignments per start position.
struct CoverageStatisticsCalculator : PileupListener {
    int max_coverage = 300;
    std::vector<long> counts, counts_unique;
    long high = 0, high_unique = 0;
    CoverageStatisticsCalculator() : counts(300, 0), counts_unique(300, 0) {}
    void on_pileup(PileupRecord& p) override {
        if (p.n_alns < (int)counts.size()) counts[p.n_alns]++; else high++;
        if (p.n_unique < (int)counts_unique.size()) counts_unique[p.n_unique]++; else high_unique++;
    }
    void on_sequence_start(const std::string&) override {}
    void on_sequence_end(const std::string&) override {}
    void print(std::ostream& out) const {
        for (size_t i = 1; i < counts.size(); i++) out << i << "\t" << counts[i] << "\t" << counts_unique[i] << "\n";
        out << "More\t" << high << "\t" << high_unique << "\n";
    }
};

char* oracle_run_coverage_stats(const char* fasta_path, const char* sam_path, int min_mq, double* seconds) {
    ReferenceGenome genome = ReferenceGenome::load_fasta(fasta_path);
    std::ifstream sam(sam_path);
    if (!sam) return nullptr;
    auto t0 = std::chrono::steady_clock::now();
    CoverageStatisticsCalculator calc;
    AlignmentsPileupGenerator gen;
    gen.secondary = true;              // setProcessSecondaryAlignments(true)
    gen.max_alns_per_start = 100;      // setMaxAlnsPerStartPos(100)
    gen.min_mq = min_mq;
    gen.listeners = {&calc};
    AlignmentReader reader;
    reader.filter_flags = gen.filter_flags();
    reader.min_mq = gen.min_mq;
    reader.genome = &genome;
    std::string line;
    SamRecord rec;
    while (std::getline(sam, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (!parse_sam_line(line, rec)) continue;
        ReadPtr a = reader.load(rec);
        if (a) gen.process_alignment(a);
    }
    gen.notify_end();
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::ostringstream out;
    calc.print(out);
    std::string s = out.str();
    char* res = (char*)std::malloc(s.size() + 1);
    std::memcpy(res, s.c_str(), s.size() + 1);
    return res;
}

// ---- math/Distribution.java (what RelativeAlleleCountsCalculator uses of it) ----------------------------------------------------
struct Distribution {
    double sum = 0, sum_sq = 0, count = 0, min_v, max_v, bin;
    std::vector<double> dist;
    long less = 0, more = 0;
    Distribution(double mn, double mx, double b) : min_v(mn), max_v(mx), bin(b), dist((size_t)((int)((mx - mn) / b) + 1), 0.0) {}   // :46-55
    void process(double value) {                                        // processDatapoint(1, value), :59-78
        sum += value; sum_sq += value * value; count += 1;
        if (value >= min_v && value <= max_v) dist[(size_t)(int)((value - min_v) / bin)] += 1;
        else if (value < min_v) less++;
        else if (value > max_v) more++;
    }
    void print(std::ostream& out, bool integer_counts) const {          // printDistribution(out, integerCounts, maxValueDistribution), :302-345
        if (less > 0) { out << "Less\t"; if (integer_counts) out << less << "\n"; else out << english_fmt((double)less) << "\n"; }
        const int max_idx = (int)((max_v - min_v) / bin);   // getBinIndex(double maxValue), :268-270
        for (int i = 0; i < (int)dist.size() && i <= max_idx; i++) {
            if (integer_counts) out << (int)(min_v + i * bin) << "\t" << (int)java_round(dist[(size_t)i]) << "\n";
            else out << english_fmt(min_v + i * bin) << "\t" << english_fmt(dist[(size_t)i]) << "\n";
        }
        double more_count = (double)more;
        for (int i = max_idx + 1; i < (int)dist.size(); i++) more_count += dist[(size_t)i];
        if (more_count > 0) { out << "More\t"; if (integer_counts) out << (int)java_round(more_count) << "\n"; else out << english_fmt(more_count) << "\n"; }
        if (count > 2147483647.0) out << "Count\t" << java_round(count) << "\n"; else out << "Count\t" << (int)java_round(count) << "\n";
        if (integer_counts) { if (sum > 2147483647.0) out << "Sum\t" << java_round(sum) << "\n"; else out << "Sum\t" << (int)java_round(sum) << "\n"; }
        else out << "Sum\t" << english_fmt(sum) << "\n";
        if (count > 0) out << "Average\t" << english_fmt(sum / count) << "\n";
        if (count > 1) {
            double variance = (sum_sq - sum * sum / count) / (count - 1);
            out << "Variance\t" << english_fmt(variance) << "\n";
            out << "STDev\t" << english_fmt(std::sqrt(variance)) << "\n";
        }
    }
};

// ---- discovery/RelativeAlleleCountsCalculator.java -----------------------------------------------------------------------------
// The listener (onPileup :246-290, onSequenceStart :309-320, printResults :221-243) on the generator runProcess configures for it
// (:209-219): secondary alignments on request, maxRD alignments per start position.
//
// Regions to filter / select (-frs, -srs; :131-161, onPileup :247-254): region files read by
// genome/io/SimpleGenomicRegionFileHandler.loadRegions (:57-80: items = line.split(" |\t"); sequence, Integer.parseInt(first),
// Integer.parseInt(last); a line that does not parse is skipped with a warning), looked up with findSpanningRegions(sequence,
// position): the regions of that sequence with first <= position <= last.
struct OracleRegions {
    std::map<std::string, std::vector<std::pair<int, int>>> of_seq;
    bool given = false;
    static bool parse_int(const std::string& t, int& v) {   // Integer.parseInt: optional sign, decimal digits, int range
        if (t.empty()) return false;
        size_t i = (t[0] == '-' || t[0] == '+') ? 1 : 0;
        if (i == t.size()) return false;
        long long x = 0;
        for (; i < t.size(); i++) { if (t[i] < '0' || t[i] > '9') return false; x = x * 10 + (t[i] - '0'); if (x > 2147483648LL) return false; }
        if (t[0] == '-') x = -x;
        if (x > 2147483647LL || x < -2147483648LL) return false;
        v = (int)x;
        return true;
    }
    bool load(const char* path) {
        if (!path || !*path) return true;
        std::ifstream in(path);
        if (!in) return false;
        given = true;
        std::string line;
        while (std::getline(in, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            std::vector<std::string> items(1);
            for (char c : line) { if (c == ' ' || c == '\t') items.emplace_back(); else items.back() += c; }
            while (!items.empty() && items.back().empty()) items.pop_back();   // String.split drops trailing empty strings
            int first, last;
            if (items.size() < 3 || !parse_int(items[1], first) || !parse_int(items[2], last)) continue;
            of_seq[items[0]].push_back({first, last});
        }
        return true;
    }
    bool spanning(const std::string& seq, int pos) const {
        auto it = of_seq.find(seq);
        if (it == of_seq.end()) return false;
        for (auto& r : it->second) if (r.first <= pos && pos <= r.second) return true;
        return false;
    }
};

struct RelativeAlleleCountsCalculator : PileupListener {
    OracleRegions filter, select;
    int min_rd = 10, min_bq = 20;
    const ReferenceGenome* genome = nullptr;
    Distribution dist_prop{0, 0.5, 0.01}, dist_num{1, 10, 1};
    std::vector<std::string> seq_names;
    std::vector<Distribution> per_seq;
    bool cur = false;
    void on_pileup(PileupRecord& p) override {
        if (filter.given && filter.spanning(p.seq, p.pos)) return;    // :247-250
        if (select.given && !select.spanning(p.seq, p.pos)) return;   // :251-254
        auto calls = p.allele_calls(1, (const std::string*)nullptr);
        if ((int)calls.size() < min_rd) return;
        std::map<std::string, int> counts;   // TreeMap
        for (auto& c : calls) {
            int qs = (int)(int8_t)c.quals[0] - 33;
            // (htsjdk hands the bases over in upper case — SAMRecord.setReadString normalises them — so the listener never sees a
            // lower-case allele; this oracle's SAM reader keeps the file's case, hence the upper-casing here)
            if (qs >= min_bq) counts[c.allele()]++;
        }
        if (counts.empty()) return;
        dist_num.process((double)counts.size());
        auto maximum = [&](const std::string* ignore) -> const std::string* {   // getMaximum, :292-305
            const std::string* mx = nullptr; int mc = -1;
            for (auto& kv : counts) { if (ignore && kv.first == *ignore) continue; if (!mx || mc < kv.second) { mx = &kv.first; mc = kv.second; } }
            return mx;
        };
        const std::string* a1 = maximum(nullptr);
        int c1 = 0, c2 = 0;
        if (a1) { c1 = counts[*a1]; const std::string* a2 = maximum(a1); if (a2) c2 = counts[*a2]; }
        if (c1 > 0) {
            double prop = (double)c2 / (c1 + c2);
            dist_prop.process(prop);
            if (cur) per_seq.back().process(prop);
        }
    }
    void on_sequence_start(const std::string& seq) override {
        int i = genome->find(seq);
        if (i >= 0 && genome->seqs[(size_t)i].size() > 100000) { seq_names.push_back(seq); per_seq.emplace_back(0, 0.5, 0.01); cur = true; }
        else cur = false;
    }
    void on_sequence_end(const std::string&) override {}
    void print(std::ostream& out) const {
        out << "Distribution of allele proportions\n";
        dist_prop.print(out, false);
        out << "Distribution of number of alleles\n";
        dist_num.print(out, true);
        if (seq_names.empty()) return;
        out << "Distribution of allele proportions per sequence\nProportion";
        for (auto& n : seq_names) out << "\t" << n;
        out << "\n";
        double mn = 0;
        for (size_t i = 0; i < dist_prop.dist.size(); i++) {
            out << english_fmt(mn);
            for (auto& d : per_seq) out << "\t" << english_fmt(d.dist[i]);
            out << "\n";
            mn += 0.01;
        }
    }
};

char* oracle_run_relative_allele_counts_regions(const char* fasta_path, const char* sam_path, int min_rd, int max_rd, int min_bq, int secondary, const char* frs,
                                                const char* srs, double* seconds);
char* oracle_run_relative_allele_counts(const char* fasta_path, const char* sam_path, int min_rd, int max_rd, int min_bq, int secondary, double* seconds) {
    return oracle_run_relative_allele_counts_regions(fasta_path, sam_path, min_rd, max_rd, min_bq, secondary, nullptr, nullptr, seconds);
}
char* oracle_run_relative_allele_counts_regions(const char* fasta_path, const char* sam_path, int min_rd, int max_rd, int min_bq, int secondary, const char* frs,
                                                const char* srs, double* seconds) {
    ReferenceGenome genome = ReferenceGenome::load_fasta(fasta_path);
    std::ifstream sam(sam_path);
    if (!sam) return nullptr;
    auto t0 = std::chrono::steady_clock::now();
    RelativeAlleleCountsCalculator calc;
    calc.min_rd = min_rd; calc.min_bq = min_bq; calc.genome = &genome;
    if (!calc.filter.load(frs) || !calc.select.load(srs)) return nullptr;
    AlignmentsPileupGenerator gen;
    gen.secondary = secondary != 0;      // setProcessSecondaryAlignments(secondaryAlns)
    gen.max_alns_per_start = max_rd;     // setMaxAlnsPerStartPos(maxRD)
    gen.listeners = {&calc};
    AlignmentReader reader;
    reader.filter_flags = gen.filter_flags();
    reader.min_mq = gen.min_mq;
    reader.genome = &genome;
    std::string line;
    SamRecord rec;
    while (std::getline(sam, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (!parse_sam_line(line, rec)) continue;
        ReadPtr a = reader.load(rec);
        if (a) gen.process_alignment(a);
    }
    gen.notify_end();
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::ostringstream out;
    calc.print(out);
    std::string s = out.str();
    char* res = (char*)std::malloc(s.size() + 1);
    std::memcpy(res, s.c_str(), s.size() + 1);
    return res;
}

// MultisampleVariantsDetector on SAM text files.  counts_out: pooled pos,A,C,G,T,DP wherever discoverPopulationSNV ran.
char* oracle_run_msvd(const char* fasta_path, int n_files, char** sam_paths, int argc, char** argv, int** counts_out, long* n_counts,
                      long* genotyped_positions, double* seconds) {
    DetectorOptions opt;
    std::string in, ref, out, dump, stats;
    if (parse_options(argc, argv, 0, opt, in, ref, out, dump, stats)) return nullptr;
    ReferenceGenome genome = ReferenceGenome::load_fasta(fasta_path);
    std::vector<std::string> files;
    for (int i = 0; i < n_files; i++) files.push_back(sam_paths[i]);
    std::ostringstream vcf;
    std::vector<std::array<int, 6>> counts;
    MsvdResult r; std::string err;
    if (!run_msvd(genome, files, vcf, opt, counts_out ? &counts : nullptr, r, err)) return nullptr;
    if (counts_out) {
        *counts_out = (int*)std::malloc(counts.size() * 6 * sizeof(int) + 8);
        std::memcpy(*counts_out, counts.data(), counts.size() * 6 * sizeof(int));
        *n_counts = (long)counts.size();
    }
    if (genotyped_positions) *genotyped_positions = r.n_nonempty;
    if (seconds) *seconds = r.seconds;
    std::string s = vcf.str();
    char* res = (char*)std::malloc(s.size() + 1);
    std::memcpy(res, s.c_str(), s.size() + 1);
    return res;
}

// DiversityStatistics + ParseUtils.ENGLISHFMT known-answer hooks
void oracle_english_fmt(double v, char* out16) { std::string s = english_fmt(v); std::strncpy(out16, s.c_str(), 15); out16[15] = 0; }
int oracle_string_hashset_order(int n, char** keys, int* order_out) {
    std::vector<std::string> k;
    for (int i = 0; i < n; i++) k.push_back(keys[i]);
    auto o = java_string_hashset_order(k);
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) if (o[i] == k[j]) order_out[i] = j;
    return n;
}

// DiversityStatistics known-answer hook: called[2k], called[2k+1] = allele indexes of sample k (-1 = none; equal = homozygous),
// cn[k * n_alleles + j] = its copy number of allele j.  out: "NS=..;AN=..;AFS=..;OH=..[;MAF=..]".
void oracle_diversity_info(int n_alleles, int n_calls, const int* called, const int* cn, char* out256) {
    std::vector<std::vector<int8_t>> idx;
    std::vector<std::vector<short>> cns;
    for (int k = 0; k < n_calls; k++) {
        std::vector<int8_t> i;
        if (called[2 * k] >= 0) { i.push_back((int8_t)called[2 * k]); if (called[2 * k + 1] != called[2 * k]) i.push_back((int8_t)called[2 * k + 1]); }
        idx.push_back(i);
        std::vector<short> c;
        for (int j = 0; j < n_alleles; j++) c.push_back((short)cn[k * n_alleles + j]);
        cns.push_back(c);
    }
    std::string s = diversity_info((size_t)n_alleles, idx, cns);
    std::strncpy(out256, s.c_str(), 255); out256[255] = 0;
}

void oracle_free(void* p) { std::free(p); }

}  // extern "C"
