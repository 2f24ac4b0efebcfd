// -knownVariants (SingleSampleVariantsDetector.java:897-905): the detector genotypes the variants of an input VCF instead of
// discovering variants.  Three things change on the path:
//   * the input variants go to the indel realigner (IRPL:89-99, 141-157): a variant is a fixed event at its first position (span =
//     its reference allele), positions behind it are embedded and left alone;
//   * the variant listener genotypes every input variant whose first position has a pileup (SSVPL:162-176, 361-391): SNVs on the
//     device as known sites (ngsep_tile_sites: VD.genotypeSNV, VD:21-97), the others here on the host (VD.callIndel with the variant
//     given, VD:265-361) — homozygous-reference and undecided genotypes are reported as well;
//   * records carry the ID and QUAL columns of the input (vcf/VCFFileWriter.java:77-98).
// This file: the VCF loader (vcf/VCFFileReader.loadVariants, VCFFileReader.java:152-306, 585-600, load mode MINIMAL), the realigner's
// view of the variants and the record text.
#include <algorithm>
#include <cerrno>
#include <climits>
#include <cstdlib>
#include <fstream>

#include "pipeline.hpp"

namespace ngsep_b200 {
namespace host {

namespace {

std::vector<std::string> split(const std::string& s, char delim) {   // main/io/ParseUtils.java:38-53: empty tokens are kept
    std::vector<std::string> out(1);
    for (char c : s) { if (c == delim) out.emplace_back(); else out.back().push_back(c); }
    return out;
}

bool parse_int(const std::string& s, long& v) {   // Integer.parseInt
    if (s.empty()) return false;
    size_t i = (s[0] == '-' || s[0] == '+') ? 1 : 0;
    if (i == s.size()) return false;
    for (size_t k = i; k < s.size(); k++) if (s[k] < '0' || s[k] > '9') return false;
    errno = 0;
    v = std::strtol(s.c_str(), nullptr, 10);
    return errno == 0 && v >= INT_MIN && v <= INT_MAX;
}

int type_id(const std::string& name) {   // GenomicVariantImpl.java:44-58
    static const char* names[] = {"SNV", "MULTISNV", "EMBEDDED", "INDEL", "STR"};
    for (int k = 0; k < 5; k++) if (name == names[k]) return k + 1;
    static const char* sv[] = {"CNV", "REPEAT", "DEL", "INS", "INV", "DUP"};
    for (int k = 0; k < 6; k++) if (name == sv[k]) return 10 + k;
    static const char* old_names[] = {"Deletion", "Insertion", "Inversion", "Duplication"};
    for (int k = 0; k < 4; k++) if (name == old_names[k]) return 12 + k;
    return 0;
}

const char* type_name(int t) {
    switch (t) { case 2: return "MULTISNV"; case 3: return "EMBEDDED"; case 4: return "INDEL"; case 5: return "STR"; }
    return nullptr;   // undetermined and SNV are not written (VCFFileWriter.java:47-49)
}

bool acgt(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }   // DNASequence.isInAlphabeth

}  // namespace

bool load_known_variants(const std::string& path, const Genome& genome, std::vector<std::vector<KnownVariant>>& out, std::string& err) {
    std::ifstream in(path);
    if (!in) { err = "cannot open " + path; return false; }
    out.assign(genome.seqs.size(), {});
    std::string line;
    bool have_samples = false;
    size_t n_samples = 0;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (!have_samples) {   // loadHeader (:142-149) + VCFFileHeader.loadSampleIds (VCFFileHeader.java:155-167)
            if (line.compare(0, 2, "##") == 0) continue;
            if (line.empty() || line[0] != '#') { err = "VCF file does not have line with sample ids"; return false; }
            const size_t cols = split(line, '\t').size();
            n_samples = cols > 9 ? cols - 9 : 0;
            have_samples = true;
            continue;
        }
        const std::vector<std::string> it = split(line, '\t');
        if (it.size() < 8) continue;                                      // :154-157
        long position = 0;
        if (!parse_int(it[1], position) || position <= 0) continue;       // :201-208
        if (it[3].empty() || it[4].empty()) continue;                     // :213-220
        KnownVariant v;
        v.alleles.push_back(it[3]);
        if (it[4][0] != '.') for (auto& a : split(it[4], ',')) v.alleles.push_back(a);
        if (v.alleles.size() > 100) continue;                             // GenomicVariant.MAX_NUM_ALLELES
        if (!it[5].empty() && it[5][0] != '.') {                          // :229-241
            char* end = nullptr;
            double d = std::strtod(it[5].c_str(), &end);
            if (!end || *end) continue;
            if (d > 32767) d = 255;
            v.qs = (int)(short)jround(d);
            if (v.qs < 0) continue;
        }
        v.first = (int)position;
        if (it[2] != ".") v.id = it[2];
        if (v.alleles.size() == 2 && it[3].size() == 1 && it[4].size() == 1 && acgt(it[3][0]) && acgt(it[4][0])) {
            v.snv_class = true; v.type = 1; v.last = v.first;             // variants/SNV.java:36-51
        } else {
            v.last = v.first + (int)v.alleles[0].size() - 1;              // GenomicVariantImpl.java:99-106
        }
        if (it[7] != ".") {                                               // loadInfoField, :261-306: TYPE, END and SVTYPE change the variant
            std::vector<std::string> info(1);
            bool text = false;
            for (char c : it[7]) { if (c == '"') text = !text; if (c == ';' && !text) info.emplace_back(); else info.back().push_back(c); }
            for (auto& f : info) {
                const size_t eq = f.find('=');
                if (eq == std::string::npos) continue;
                const std::string attr = f.substr(0, eq), value = f.substr(eq + 1);
                if (attr == "TYPE") { const int t = type_id(value); if (t > 0 && t <= 14) v.type = t; }
                else if (attr == "END" && !v.snv_class) { long l; if (!parse_int(value, l)) { err = "END is not a number: " + value; return false; } v.last = (int)l; }
                else if (attr == "SVLEN" && !v.snv_class) { long l; if (!parse_int(value, l)) { err = "SVLEN is not a number: " + value; return false; } }
                else if (attr == "SVTYPE") { const int t = type_id(value); if (t >= 10) v.type = t; }
            }
        }
        if (it.size() == 8) { if (n_samples > 0) continue; }              // :166-181
        else if (it.size() - 9 != n_samples) continue;
        if (v.alleles.size() < 2) continue;                               // reference sites of a gVCF (:593)
        if (v.type >= 10) continue;                                       // structural variants (:594)
        v.is_snv = v.snv_class;
        if (!v.is_snv) {                                                  // GenomicVariantImpl.isSNV, :254-261
            v.is_snv = true;
            for (auto& a : v.alleles) if (a.size() != 1 || !acgt(a[0])) v.is_snv = false;
        }
        if (v.is_snv && v.alleles.size() > 4) { err = "an SNV with more than four alleles at " + it[0] + ":" + it[1] + " (repeated bases) is not supported"; return false; }
        if (v.is_snv) {
            unsigned code = (unsigned)v.alleles.size() << 8;
            for (size_t k = 0; k < v.alleles.size(); k++) code |= (unsigned)dna_index(v.alleles[k][0]) << (2 * k);
            v.site_code = (uint16_t)code;
        }
        int seq = -1;
        for (size_t i = 0; i < genome.names.size(); i++) if (genome.names[i] == it[0]) { seq = (int)i; break; }
        if (seq < 0) continue;   // a sequence without alignments: its variants never meet a pileup
        out[(size_t)seq].push_back(std::move(v));
    }
    if (!have_samples) { err = "VCF file does not have line with sample ids"; return false; }
    for (auto& vs : out)
        std::stable_sort(vs.begin(), vs.end(), [](const KnownVariant& a, const KnownVariant& b) { return a.first != b.first ? a.first < b.first : a.last < b.last; });
    return true;
}

// IRPL:141-157 walks the sorted list with a cursor that only passes a variant once the position lies behind its end, and returns the
// variant under the cursor when it contains the position.  So variant k is returned exactly for the positions of
// [max(first_k, reach_k + 1), last_k], reach_k = the largest end of the variants before it; it STARTS an event (IRPL:90) iff
// first_k > reach_k, otherwise its visible tail is embedded only.
std::vector<StrLocus> known_variant_loci(const std::vector<KnownVariant>& vars) {
    std::vector<StrLocus> loci;
    int reach = 0;
    for (const KnownVariant& v : vars) {
        if (v.last > reach) {
            StrLocus l;
            l.str = v.type == 5;
            l.starter = v.first > reach;
            l.first = l.starter ? v.first : reach + 1;
            l.last = v.last;
            loci.push_back(l);
            reach = v.last;
        }
    }
    return loci;
}

namespace {

void record_head(std::string& s, const KnownVariant& v, const std::string& chrom) {   // VCFFileWriter.java:77-98, 44-52
    s = chrom;
    s += "\t"; s += std::to_string(v.first); s += "\t"; s += v.id.empty() ? "." : v.id; s += "\t"; s += v.alleles[0]; s += "\t";
    for (size_t i = 1; i < v.alleles.size(); i++) { if (i > 1) s += ","; s += v.alleles[i]; }
    s += "\t"; s += std::to_string(v.qs); s += "\t.\t";
    const char* tn = type_name(v.type);
    if (tn) { s += "TYPE="; s += tn; } else s += ".";
}

void append_copy_numbers(std::string& f, std::vector<int> cn, bool undecided, int ploidy) {   // VCFFileWriter.java:236-248
    if (ploidy == 0) { f += "."; return; }
    if (undecided) cn[0] = ploidy;
    for (size_t j = 0; j < cn.size(); j++) { if (j) f += ","; f += std::to_string(cn[j]); }
}

}  // namespace

std::string known_snv_vcf_line(const KnownVariant& v, const ngsep_snv_call* c, const std::string& chrom, int ploidy) {
    std::string s;
    record_head(s, v, chrom);
    s += "\tGT:PL:GQ:DP:BSDP:ACN\t";   // CalledSNV, or a CalledGenomicVariantImpl with all counts (SSVD:947-952)
    const int nA = (int)v.alleles.size();
    int A[4] = {0, 0, 0, 0};
    for (int k = 0; k < nA; k++) A[k] = dna_index(v.alleles[k][0]);
    const bool no_data = !c || c->depth == 0;   // VD:23-27: undecided, the counts and no report
    std::vector<int> called;
    if (!no_data && c->a0 >= 0) { if (c->a0 == c->a1) called = {c->a0}; else called = {c->a0, c->a1}; }
    append_gt(s, called, ploidy);
    s += ":";
    auto L = [&](int i, int j) { return c->logcond[A[i] * 4 + A[j]]; };
    std::vector<int> cn;
    if (v.snv_class && !no_data) {
        const float hr = (float)L(0, 0), ra = (float)L(0, 1), ar = (float)L(1, 0), ha = (float)L(1, 1);   // CalledSNV.java:259-266
        const bool present = (float)(hr + ra + ar + ha) != 0.0f;                                        // CalledSNV.java:422
        const double pl[3] = {hr, ra, ha};
        for (int k = 0; k < 3; k++) { if (k) s += ","; s += std::to_string(present ? jround(-10 * pl[k]) : 0L); }
        if (called.empty()) cn = {0, 0};                                                                // CalledSNV.java:134-147
        else if (called.size() == 1) cn = (called[0] == 0) ? std::vector<int>{ploidy, 0} : std::vector<int>{0, ploidy};
        else cn = {1, 1};
    } else {
        bool firstv = true;
        for (int j = 0; j < nA; j++) for (int i = 0; i <= j; i++) {
            if (!firstv) s += ",";
            firstv = false;
            s += std::to_string(no_data ? 0L : jround(-10 * L(i, j)));
        }
        std::vector<int> rep((size_t)nA, 0);
        if (!no_data) for (int i = 0; i < nA; i++) rep[(size_t)i] = c->count[A[i]];
        cn = impl_copy_numbers(called, rep, !no_data, ploidy, (size_t)nA);
    }
    s += ":"; s += std::to_string(no_data ? 0 : (int)c->gq);
    s += ":"; s += std::to_string(no_data ? 0 : c->depth);
    s += ":";
    for (int i = 0; i < 4; i++) { if (i) s += ","; s += std::to_string(c ? c->count[i] : 0); }
    s += ":";
    append_copy_numbers(s, cn, called.empty(), ploidy);
    s += "\n";
    return s;
}

std::string known_indel_vcf_line(const KnownVariant& v, const SampleIndelCall& c, const std::string& chrom, int ploidy) {
    std::string s;
    record_head(s, v, chrom);
    s += "\tGT:PL:GQ:DP:ADP:ACN\t";
    const size_t n = v.alleles.size();
    // VariantCallReport resolves alleles through a TreeMap name->index: with duplicate names the last index wins
    auto index_of = [&](const std::string& a) { int k = -1; for (size_t i = 0; i < n; i++) if (v.alleles[i] == a) k = (int)i; return k; };
    append_gt(s, c.called, ploidy);
    s += ":";
    bool firstv = true;
    for (size_t j = 0; j < n; j++) for (size_t i = 0; i <= j; i++) {
        if (!firstv) s += ",";
        firstv = false;
        s += std::to_string(c.has_report ? jround(-10 * c.logc[(size_t)index_of(v.alleles[i])][(size_t)index_of(v.alleles[j])]) : 0L);
    }
    s += ":"; s += std::to_string(c.gq);
    s += ":"; s += std::to_string(c.depth);
    s += ":";
    std::vector<int> rep(n, 0);
    for (size_t i = 0; i < n; i++) { rep[i] = c.has_report ? c.counts[(size_t)index_of(v.alleles[i])] : 0; if (i) s += ","; s += std::to_string(rep[i]); }
    std::vector<int> cn = impl_copy_numbers(c.called, rep, c.has_report, ploidy, n);
    s += ":";
    append_copy_numbers(s, cn, c.called.empty(), ploidy);
    s += "\n";
    return s;
}

// A pool / polyploid SNV call (SSVPL:240-254): a CalledGenomicVariantImpl with a call report and no base counts of its own, so the record
// takes the ADP format (SSVD:947-952); the variant has no quality; its type is SNV (not written), MULTISNV while three or more alleles
// stand (SSVPL:326-328), EMBEDDED inside an indel (SSVPL:229).
std::string pool_snv_vcf_line(int pos, const std::vector<int>& alleles, const PoolCall& c, const std::string& chrom, int ploidy, bool embedded) {
    static const char B[] = "ACGT";
    KnownVariant v;
    v.first = v.last = pos;
    for (int a : alleles) v.alleles.emplace_back(1, B[a]);
    v.type = embedded ? 3 : (alleles.size() > 2 ? 2 : 1);
    SampleIndelCall sc;
    sc.called = c.called; sc.gq = c.gq; sc.depth = c.depth; sc.has_report = true; sc.counts = c.counts; sc.logc = c.logc;
    return known_indel_vcf_line(v, sc, chrom, ploidy);
}

// A known SNV genotyped as a pool (SSVPL:366-370): a CalledGenomicVariantImpl with a call report AND all base counts (setAllCounts), so the
// record takes the BSDP format; an undecided result of genotypeVariantPool carries neither depth nor report (SSVPL:432-436).
std::string known_pool_snv_vcf_line(const KnownVariant& v, const ngsep_snv_call* col, const SampleIndelCall& c, const std::string& chrom, int ploidy) {
    std::string s;
    record_head(s, v, chrom);
    s += "\tGT:PL:GQ:DP:BSDP:ACN\t";
    const size_t n = v.alleles.size();
    append_gt(s, c.called, ploidy);
    s += ":";
    bool firstv = true;
    for (size_t j = 0; j < n; j++) for (size_t i = 0; i <= j; i++) {
        if (!firstv) s += ",";
        firstv = false;
        s += std::to_string(c.has_report ? jround(-10 * c.logc[i][j]) : 0L);
    }
    s += ":"; s += std::to_string(c.gq);
    s += ":"; s += std::to_string(c.depth);
    s += ":";
    for (int i = 0; i < 4; i++) { if (i) s += ","; s += std::to_string(col ? col->count[i] : 0); }
    std::vector<int> rep(n, 0);
    if (c.has_report) rep = c.counts;
    std::vector<int> cn = impl_copy_numbers(c.called, rep, c.has_report, ploidy, n);
    s += ":";
    append_copy_numbers(s, cn, c.called.empty(), ploidy);
    s += "\n";
    return s;
}

}  // namespace host
}  // namespace ngsep_b200
