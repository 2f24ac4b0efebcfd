// TEST INFRASTRUCTURE ONLY — see ngsep_oracle.hpp.
// Restatement of discovery/MultisampleVariantsDetector.java (MSVD), the static pool helpers of
// discovery/SingleSampleVariantPileupListener.java (SSVPL:297-359), vcf/VCFRecord.createDefaultPopulationVCFRecord
// (vcf/VCFRecord.java:277-301) and variants/DiversityStatistics.calculateDiversityStatistics
// (variants/DiversityStatistics.java:123-218), discovery mode only (no -knownVariants / -knownSTRs).
// Parity unpinned: the reference holds no golden vector for the multi-sample logic.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <istream>
#include <map>
#include <set>
#include <sstream>

#include "ngsep_oracle.hpp"

namespace ngsep_oracle {

static inline int dna_index_m(char c) {
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; }
    return -1;
}

// java.util.HashSet<String> iteration order for distinct keys given in insertion order: String.hashCode, HashMap's
// spreading and table growth (16, doubling past 0.75), bins in insertion order (treeified bins — 8+ collisions in a table
// of 64+ — are not modelled).
std::vector<std::string> java_string_hashset_order(const std::vector<std::string>& keys) {
    size_t cap = 16;
    while (keys.size() > cap * 3 / 4) cap *= 2;
    std::vector<std::pair<unsigned, size_t>> tagged;
    for (size_t i = 0; i < keys.size(); i++) {
        unsigned h = 0;
        for (unsigned char c : keys[i]) h = 31u * h + c;
        h ^= (h >> 16);
        tagged.push_back({h & (unsigned)(cap - 1), i});
    }
    std::stable_sort(tagged.begin(), tagged.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
    std::vector<std::string> out;
    for (auto& t : tagged) out.push_back(keys[t.second]);
    return out;
}

// MSVD.loadSamplesFromAlignmentHeaders (MSVD:492-516) + RAFR.loadHeader (RAFR:184-197) over the @RG lines of SAM texts,
// files in the given order.  Samples come back in TreeMap (id) order; each sample's read groups in HashSet order.
bool load_samples_from_headers(const std::vector<std::vector<std::string>>& header_lines_per_file, short ploidy, std::vector<Sample>& out, std::string& err) {
    std::map<std::string, std::vector<std::string>> by_sample;   // insertion order of read groups per sample
    std::map<std::string, std::string> rg_sample;
    for (auto& lines : header_lines_per_file) {
        std::vector<std::pair<std::string, std::string>> file_rgs;   // htsjdk keeps @RG records in file order; RAFR keeps the first of an id
        std::map<std::string, std::string> file_map;
        for (auto& l : lines) {
            if (l.compare(0, 3, "@RG") != 0) continue;
            std::string id, sm; bool has_sm = false;
            std::stringstream ss(l);
            std::string tok;
            while (std::getline(ss, tok, '\t')) {
                if (!tok.compare(0, 3, "ID:")) id = tok.substr(3);
                else if (!tok.compare(0, 3, "SM:")) { sm = tok.substr(3); has_sm = true; }
            }
            if (!has_sm) sm = id;                                       // RAFR:188
            if (!file_map.count(id)) { file_map[id] = sm; file_rgs.push_back({id, sm}); }
            else if (file_map[id] != sm) { err = "The read group ID: " + id + " is associated to two different samples"; return false; }
        }
        // getSampleIdsByReadGroup returns a HashMap: key iteration in hash order (MSVD:498)
        std::vector<std::string> ids;
        for (auto& p : file_rgs) ids.push_back(p.first);
        for (auto& id : java_string_hashset_order(ids)) {
            const std::string& sm = file_map[id];
            auto it = rg_sample.find(id);
            if (it != rg_sample.end() && it->second != sm) { err = "The read group ID: " + id + " is associated to two different samples"; return false; }
            rg_sample[id] = sm;
            auto& v = by_sample[sm];
            if (std::find(v.begin(), v.end(), id) == v.end()) v.push_back(id);   // Sample.readGroups is a HashSet
        }
    }
    out.clear();
    for (auto& kv : by_sample) out.push_back(Sample{kv.first, ploidy, java_string_hashset_order(kv.second)});
    return true;
}

// ---- SSVPL static pool helpers ---------------------------------------------------------------------------------
static GenomicVariant make_impl_variant(const std::string& seq, int first, const std::vector<std::string>& alleles) {  // GenomicVariantImpl.java:99-106
    GenomicVariant v;
    v.seq = seq; v.first = first; v.len = (int)alleles[0].size(); v.last = first + v.len - 1;
    v.alleles = alleles;
    return v;
}
static GenomicVariant make_snv_class(const std::string& seq, int pos, char ref, char alt) {  // SNV.java:36-51
    GenomicVariant v;
    v.seq = seq; v.first = pos; v.last = pos; v.len = 1;
    v.alleles = {std::string(1, ref), std::string(1, alt)};
    v.type = TYPE_BIALLELIC_SNV; v.snv_class = true;
    return v;
}

bool create_snv_variant_pool(const PileupRecord& p, const CountsHelper& h, char reference, double min_allele_depth_freq, GenomicVariant& out) {  // SSVPL:297-332
    if (h.total == 0) return false;
    int ref_idx = dna_index_m(reference);
    if (ref_idx < 0) return false;
    int sum = 0;
    for (int c : h.counts) sum += c;
    double min_count = std::max(1.0, min_allele_depth_freq * sum);
    std::vector<std::string> alleles = {std::string(1, "ACGT"[ref_idx])};
    for (int i = 0; i < (int)h.counts.size(); i++) if (h.counts[i] >= min_count && i != ref_idx) alleles.push_back(std::string(1, "ACGT"[i]));
    if (alleles.size() == 2) { out = make_snv_class(p.seq, p.pos, reference, alleles[1][0]); return true; }
    if (alleles.size() > 2) { out = make_impl_variant(p.seq, p.pos, alleles); out.type = TYPE_MULTIALLELIC_SNV; return true; }
    return false;
}

bool alleles_same_length(const std::vector<std::string>& a) {  // SSVPL:339-345
    for (auto& s : a) if (s.size() != a[0].size()) return false;
    return true;
}

GenomicVariant make_new_variant(const GenomicVariant& v, const std::vector<std::string>& new_alleles) {  // SSVPL:346-359
    std::vector<std::string> alleles = {v.alleles[0]};
    for (auto& a : new_alleles) if (a != v.alleles[0]) alleles.push_back(a);
    if (v.is_snv() && alleles.size() == 2) return make_snv_class(v.seq, v.first, v.alleles[0][0], alleles[1][0]);
    return make_impl_variant(v.seq, v.first, alleles);   // type stays TYPE_UNDETERMINED
}

static GenomicVariant make_new_variant_static(const GenomicVariant& v, const std::set<std::string>& new_alleles) {  // SSVPL:346-359
    std::vector<std::string> alleles = {v.alleles[0]};
    for (auto& a : new_alleles) if (a != v.alleles[0]) alleles.push_back(a);
    if (v.is_snv() && alleles.size() == 2) return make_snv_class(v.seq, v.first, v.alleles[0][0], alleles[1][0]);
    return make_impl_variant(v.seq, v.first, alleles);   // type stays TYPE_UNDETERMINED
}

static std::vector<std::string> called_alleles(const CalledVariant& c) {  // CalledSNV.java:365-381 / Impl getCalledAlleles
    std::vector<std::string> out;
    for (int8_t i : c.called_indexes()) out.push_back(c.var.alleles[i]);
    return out;
}

// ---- MultisampleVariantsDetector ---------------------------------------------------------------------------------
std::vector<CalledVariant> MultisampleVariantsDetector::genotype_variant(GenomicVariant& v, PileupRecord& p, double h) {  // MSVD:674-693
    std::vector<CalledVariant> calls;
    SingleSampleVariantPileupListener sd;   // a fresh listener: minQuality 40, only maxBaseQS forwarded (MSVD:676-678)
    sd.max_base_qs = (int8_t)opt.max_base_qs;
    short qs = 0;
    for (auto& s : samples) {
        CalledVariant c = sd.genotype_variant_sample(v, p, &s, h);
        if (!c.undecided() && !c.hom_ref() && c.gq > qs) qs = c.gq;
        calls.push_back(std::move(c));
    }
    v.qs = qs;
    for (auto& c : calls) c.var.qs = qs;
    return calls;
}

bool MultisampleVariantsDetector::make_new_variant(GenomicVariant& v, const std::vector<CalledVariant>& calls) {  // MSVD:642-656; true when replaced
    if (v.alleles.size() <= 2) return false;
    std::set<std::string> called = {v.alleles[0]};
    for (auto& c : calls) for (auto& a : called_alleles(c)) called.insert(a);
    if (v.alleles.size() != called.size()) { v = make_new_variant_static(v, called); return true; }
    return false;
}

bool MultisampleVariantsDetector::discover_population_snv(PileupRecord& p, char reference, GenomicVariant& out) {  // MSVD:585-597
    auto calls = p.allele_calls(1, (const std::string*)nullptr);
    CountsHelper h = CountsHelper::snv(calls, (int8_t)opt.max_base_qs, 0.5);
    if (counts_dump) counts_dump->push_back({p.pos, h.counts[0], h.counts[1], h.counts[2], h.counts[3], h.total});
    if (!create_snv_variant_pool(p, h, reference, min_allele_depth_freq, out)) return false;
    while (out.alleles.size() > 2) {
        auto gc = genotype_variant(out, p, opt.het_rate);
        if (!make_new_variant(out, gc)) break;
    }
    return true;
}

bool MultisampleVariantsDetector::discover_population_indel(PileupRecord& p, const CountsHelper& h, GenomicVariant& out) {  // MSVD:620-634
    if (h.alleles.size() == 1) return false;       // createIndelVariantPool SSVPL:333-338
    if (h.total == 0) return false;
    out = make_impl_variant(p.seq, p.pos, h.alleles);
    while (out.alleles.size() > 2) {
        if (!p.input_str() && alleles_same_length(out.alleles)) return false;
        auto gc = genotype_variant(out, p, opt.het_rate);
        if (out.qs < opt.min_quality) return false;
        if (!make_new_variant(out, gc)) break;
    }
    out.type = p.str ? TYPE_STR : TYPE_INDEL;
    return true;
}

bool MultisampleVariantsDetector::discover_population_with_span(PileupRecord& p, const std::string& ref_allele, GenomicVariant& out) {  // MSVD:599-618
    auto calls = p.allele_calls((int)ref_allele.size(), (const std::string*)nullptr);
    auto alleles = cluster_allele_calls(calls, ref_allele, (int8_t)opt.max_base_qs);
    CountsHelper h = CountsHelper::indel(alleles, calls, (int8_t)opt.max_base_qs, 0.5);
    bool have = discover_population_indel(p, h, out);
    if (!p.input_str() && !have) {
        if (p.new_str) { p.str = false; p.new_str = false; }
        have = discover_population_snv(p, ref_allele[0], out);
    }
    return have;
}

void MultisampleVariantsDetector::on_pileup(PileupRecord& p) {  // MSVD:521-537 (inputVariants empty)
    if (p.input_str()) last_indel_end = p.pos + p.span - 1;
    else if (p.pos <= last_indel_end) p.embedded = true;
    std::string ref;
    if (!SingleSampleVariantPileupListener::reference_allele_discovery(p, *genome, opt.call_embedded, opt.ignore_lower, ref)) return;
    GenomicVariant v;
    bool have = (ref.size() > 1) ? discover_population_with_span(p, ref, v) : discover_population_snv(p, ref[0], v);
    if (!have) return;
    if (v.is_snv() && p.embedded) v.type = TYPE_EMBEDDED_SNV;   // MSVD:580-582
    auto calls = genotype_variant(v, p, opt.het_rate);
    if (v.qs == 0 || v.qs < opt.min_quality) return;
    if (out) print_vcf_record_population(*out, v, calls, samples);
    n_records++;
    if (!v.is_snv()) last_indel_end = v.last;
}

void MultisampleVariantsDetector::begin() {  // MSVD:421-456
    gen.query_seq = opt.query_seq; gen.has_query = opt.has_query; gen.query_first = opt.query_first; gen.query_last = opt.query_last;
    gen.max_alns_per_start = opt.max_alns_per_start; gen.non_unique_primary = opt.non_unique_primary; gen.secondary = opt.secondary;
    gen.ignore5 = (int8_t)opt.ignore5; gen.ignore3 = (int8_t)opt.ignore3; gen.min_mq = opt.min_mq;
    realigner.genome = genome;
    gen.listeners = {&realigner, this};   // MSVD:449-450
    if (out) print_vcf_header(*out, samples, opt.psp);
}

void MultisampleVariantsDetector::add_alignment(const ReadPtr& a) {  // APG:243-262 loop body
    if (stopped) return;
    int q = gen.query_filter(*a);
    if (q == 2) { stopped = true; return; }
    if (q == 1) return;
    gen.process_alignment(a);
}

void MultisampleVariantsDetector::finish() { gen.notify_end(); if (out) out->flush(); }

// k-way merge of per-file alignment streams, APG:212-289: genome/GenomicRegionComparator.java:51-67 order
// (sequence index, first, last), ties to the lowest file index.
void MultisampleVariantsDetector::process_files(std::vector<std::istream*>& sams) {
    const size_t n = sams.size();
    std::vector<AlignmentReader> readers(n);
    std::vector<ReadPtr> cur(n);
    for (auto& r : readers) { r.filter_flags = gen.filter_flags(); r.min_mq = gen.min_mq; r.genome = genome; }
    auto next = [&](size_t i) -> ReadPtr {
        std::string line;
        SamRecord rec;
        while (std::getline(*sams[i], line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            if (!parse_sam_line(line, rec)) continue;
            n_input_records++;
            ReadPtr a = readers[i].load(rec);
            if (a) return a;
        }
        return nullptr;
    };
    for (size_t i = 0; i < n; i++) cur[i] = next(i);
    auto cmp = [&](const ReadAlignment& a, const ReadAlignment& b) {
        if (a.seq_idx != b.seq_idx) return a.seq_idx < b.seq_idx ? -3 : 3;
        if (a.first != b.first) return a.first < b.first ? -1 : 1;
        if (a.last != b.last) return a.last < b.last ? -1 : 1;
        return 0;
    };
    while (!stopped) {
        int best = -1;
        for (size_t i = 0; i < n; i++) if (cur[i] && (best < 0 || cmp(*cur[best], *cur[i]) > 0)) best = (int)i;
        if (best < 0) break;
        ReadPtr a = cur[best];
        cur[best] = next((size_t)best);
        add_alignment(a);
    }
    finish();
}

// ---- population VCF record ------------------------------------------------------------------------------------
std::string english_fmt(double v) {  // main/io/ParseUtils.java:29 "##0.0#", HALF_EVEN on the exact binary value
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.2f", v);
    std::string s = buf;
    if (s.size() >= 2 && s.back() == '0') s.pop_back();   // second decimal is optional
    return s;
}

// VCFRecord.updateDiversityStatistics (vcf/VCFRecord.java:288-301) over DiversityStatistics.calculateDiversityStatistics
// (variants/DiversityStatistics.java:123-218, assumeAlwaysDiploid = false): "NS=..;AN=..;AFS=..;OH=..[;MAF=..]".
// idx[k]: called allele indexes of sample k (empty = undecided); cn[k]: its allele copy numbers.
std::string diversity_info(size_t na, const std::vector<std::vector<int8_t>>& idx, const std::vector<std::vector<short>>& cn) {
    std::vector<int> counts(na, 0);
    double sum = 0;
    int genotyped = 0, n_het = 0;
    for (size_t k = 0; k < idx.size(); k++) {
        if (idx[k].empty()) continue;
        genotyped++;
        if (idx[k].size() > 1) n_het++;
        for (int8_t j : idx[k]) if (j >= 0 && (size_t)j < na) { counts[j] += cn[k][j]; sum += cn[k][j]; }
    }
    int n_called = 0, min_ac = 0;
    for (size_t i = 0; i < na; i++) if (counts[i] > 0) { n_called++; if (min_ac == 0 || min_ac > counts[i]) min_ac = counts[i]; }
    double oh = genotyped > 0 ? (double)n_het / genotyped : 0.0;
    double maf = n_called < 2 ? 0.0 : (double)min_ac / sum;
    std::ostringstream out;
    out << "NS=" << genotyped << ";AN=" << n_called << ";AFS=";
    for (size_t i = 0; i < na; i++) { if (i) out << ","; out << counts[i]; }
    out << ";OH=" << english_fmt(oh);
    if (na == 2) out << ";MAF=" << english_fmt(maf);
    return out.str();
}

void print_vcf_record_population(std::ostream& out, const GenomicVariant& v, std::vector<CalledVariant>& calls, const std::vector<Sample>& samples) {
    // vcf/VCFRecord.java:277-301 + vcf/VCFFileWriter.java:44-70
    const bool snv_format = v.is_snv();
    out << v.seq << "\t" << v.first << "\t.\t" << v.alleles[0] << "\t";
    if (v.alleles.size() == 1) out << ".";
    else for (size_t i = 1; i < v.alleles.size(); i++) { if (i > 1) out << ","; out << v.alleles[i]; }
    out << "\t" << v.qs << "\t.\t";
    // DiversityStatistics.calculateDiversityStatistics(calls, false), variants/DiversityStatistics.java:123-218
    const size_t na = v.alleles.size();
    std::vector<std::vector<int8_t>> idxs;
    std::vector<std::vector<short>> cns;
    for (auto& c : calls) {
        idxs.push_back(c.called_indexes());
        if (c.called_snv) cns.push_back({c.ref_cn, (short)(c.total_cn - c.ref_cn)});   // CalledSNV.getAllelesCopyNumber :162-171
        else cns.push_back(c.alleles_cn);
    }
    out << diversity_info(na, idxs, cns);
    static const char* names[] = {nullptr, "SNV", "MULTISNV", "EMBEDDED", "INDEL", "STR"};
    if (v.type != TYPE_UNDETERMINED && v.type != TYPE_BIALLELIC_SNV && v.type >= 0 && v.type <= TYPE_STR) out << ";TYPE=" << names[v.type];
    out << "\t" << (snv_format ? "GT:PL:GQ:DP:BSDP:ACN" : "GT:PL:GQ:DP:ADP:ACN");
    for (size_t i = 0; i < calls.size(); i++) print_genotype_field(out, calls[i], snv_format, samples[i].ploidy);
    out << "\n";
}

}  // namespace ngsep_oracle
