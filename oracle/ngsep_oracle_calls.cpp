// TEST INFRASTRUCTURE ONLY — see ngsep_oracle.hpp.  Part 2: PileupRecord, CountsHelper,
// VariantDiscoverySNVQAlgorithm, AlleleCallClustersBuilder, called-variant objects.
#include "ngsep_oracle.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>

namespace ngsep_oracle {

static const std::vector<std::string> BASES_ARRAY = {"A", "C", "G", "T"};  // sequences/DNASequence.java:33
static inline int dna_index(char c) {
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; }
    return -1;
}
static std::string to_upper(const std::string& s) {
    std::string o = s;
    for (auto& c : o) if (c >= 'a' && c <= 'z') c = (char)(c - 32);
    return o;
}

std::string PileupAlleleCall::allele() const { return to_upper(seq); }  // PAC:41-43

// ---------------------------------------------------------------------------------------------
// PileupRecord
// ---------------------------------------------------------------------------------------------
void PileupRecord::add(ReadAlignment* a) {  // PR:154-167
    if (a->first > pos) return;
    if (a->last < pos) return;
    alns.push_back(a);
    by_rg[a->read_group].push_back(a);
    n_alns++;
    if (a->is_unique()) n_unique++;
    if (a->is_negative()) n_neg++;
}

std::vector<PileupAlleleCall> PileupRecord::allele_calls(int ref_span, const std::string* rg) {  // PR:126-152
    std::vector<PileupAlleleCall> out;
    const std::vector<ReadAlignment*>* list;
    if (!rg) list = &alns;
    else {
        auto it = by_rg.find(*rg);
        if (it == by_rg.end()) return out;
        list = &it->second;
    }
    for (ReadAlignment* a : *list) {
        std::string call;
        if (!a->allele_call(pos, call)) continue;
        std::string qs(1, a->base_quality(pos));
        if (ref_span > 1) {
            int last_base = pos + ref_span - 1;
            if (!a->allele_call(pos, last_base, call)) continue;
            if (!a->base_qualities(pos, last_base, qs)) qs.clear();  // Java: null string; never dereferenced because call!=null implies qs!=null
        } else if (call.size() > 1) continue;
        PileupAlleleCall c;
        c.seq = call; c.quals = qs; c.read_group = a->read_group; c.neg = a->is_negative();
        out.push_back(std::move(c));
    }
    return out;
}

std::vector<PileupAlleleCall> PileupRecord::allele_calls(int ref_span, const std::vector<std::string>* rgs) {  // PR:104-111
    if (!rgs) return allele_calls(ref_span, (const std::string*)nullptr);
    std::vector<PileupAlleleCall> out;
    for (auto& rg : *rgs) {
        auto part = allele_calls(ref_span, &rg);
        out.insert(out.end(), part.begin(), part.end());
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// CountsHelper tables (CH:135-187).  Column count n+1 with n = 100 (largest cache the reference ever builds);
// rows q = 3..30; frequencies f = 0..500.
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int NQ = 31, NJ = MAX_NUM_ALLELES + 1, NF = 501;
double g_err[NQ][NJ];
double g_freq[NF][2];
std::vector<double> g_gt;  // [f][q][j]
bool build_tables() {
    std::memset(g_err, 0, sizeof(g_err));
    for (int i = 3; i < NQ; i++) {
        g_err[i][0] = -0.1 * i;
        for (int j = 2; j < NJ; j++) g_err[i][j] = g_err[i][0] - std::log10((double)(j - 1));
    }
    g_gt.assign((size_t)NF * NQ * NJ, 0.0);
    for (int f = 0; f < NF; f++) {
        double af = (double)f / (NF - 1);
        g_freq[f][0] = std::log10(af);
        g_freq[f][1] = std::log10(1 - af);
        for (int i = 3; i < NQ; i++) {
            double e = phred_probability((short)i);
            double s = 1 - e;
            g_gt[((size_t)f * NQ + i) * NJ + 0] = std::log10(s);
            for (int j = 2; j < NJ; j++) {
                double het = af * s + (1 - af) * e / (j - 1);
                g_gt[((size_t)f * NQ + i) * NJ + j] = std::log10(het);
            }
        }
    }
    return true;
}
inline void ensure_tables() { static const bool built = build_tables(); (void)built; }
inline double GT(int f, int q, int j) { return g_gt[((size_t)f * NQ + q) * NJ + j]; }
const double LOG_ERR_INDEL = std::log10(0.0001);  // CH:48
}  // namespace

double CountsHelper::table_error(int q, int j) { ensure_tables(); return g_err[q][j]; }
double CountsHelper::table_gt(int f, int q, int j) { ensure_tables(); return GT(f, q, j); }
double CountsHelper::table_freq(int f, int k) { ensure_tables(); return g_freq[f][k]; }

CountsHelper::CountsHelper(const std::vector<std::string>& al) : alleles(al) {  // CH:125-134
    ensure_tables();
    size_t n = al.size();
    counts.assign(n, 0);
    strand.assign(n, {0, 0});
    logc.assign(n, std::vector<double>(n, 0.0));
}
CountsHelper::CountsHelper() : CountsHelper(BASES_ARRAY) {}

CountsHelper CountsHelper::snv(const std::vector<PileupAlleleCall>& calls, int8_t max_qs, double hp) {  // CH:83-95
    CountsHelper h(BASES_ARRAY);
    if (max_qs > 0) h.max_base_qs = max_qs;
    h.het_prop = hp;
    for (auto& c : calls) {
        int8_t q = (int8_t)std::min(30, (int)(unsigned char)c.quals[0] - 33);
        h.update(c.allele().substr(0, 1), q, c.neg);
    }
    return h;
}

CountsHelper CountsHelper::snv_gt(const std::vector<std::string>& al, const std::vector<PileupAlleleCall>& calls, int8_t max_qs, double hp) {  // CH:86-95
    CountsHelper h(al);
    if (max_qs > 0) h.max_base_qs = max_qs;
    h.het_prop = hp;
    for (auto& c : calls) {
        int8_t q = (int8_t)std::min(30, (int)(unsigned char)c.quals[0] - 33);
        h.update(c.allele().substr(0, 1), q, c.neg);
    }
    return h;
}

std::vector<double> CountsHelper::posteriors_major(double het_rate, int major) const {  // CH:450-467
    int n = (int)alleles.size();
    double lp_het = std::log10(het_rate / (n - 1));
    double lp_hom = std::log10(1 - het_rate);
    std::vector<double> ev((size_t)n);
    for (int j = 0; j < n; j++) ev[(size_t)j] = logc[(size_t)major][(size_t)j] + (j == major ? lp_hom : lp_het);
    posterior_from_events(ev);
    return ev;
}

CountsHelper CountsHelper::indel(const std::vector<std::string>& al, const std::vector<PileupAlleleCall>& calls, int8_t max_qs, double hp) {  // CH:96-105
    CountsHelper h(al);
    if (max_qs > 0) h.max_base_qs = max_qs;
    h.het_prop = hp;
    for (auto& c : calls) h.update_indel(c.allele(), c.quals, c.neg);
    return h;
}

void CountsHelper::update(const std::string& allele, int8_t q, bool neg) {  // CH:209-251
    total++;
    int f = (int)java_round(het_prop * (NF - 1));
    int g = (int)java_round((1 - het_prop) * (NF - 1));
    if (q <= 3) { low_q++; return; }
    else if (q > max_base_qs) q = max_base_qs;
    int index = -1;
    for (size_t i = 0; i < alleles.size(); i++) if (alleles[i] == allele) { index = (int)i; break; }
    if (index >= 0) {
        counts[index]++;
        if (neg) strand[index][0]++; else strand[index][1]++;
        int n = (int)alleles.size();
        int m = (int)logc.size();
        for (int i = 0; i < m; i++) {
            if (i == index) logc[i][i] += GT(f, q, 0);
            else logc[i][i] += g_err[q][n];
            for (int j = 0; j < m; j++) {
                if (i != j) {
                    if (j == index) logc[i][j] += GT(f, q, n);
                    else if (i == index) logc[i][j] += GT(g, q, n);
                    else logc[i][j] += g_err[q][n];
                }
            }
        }
    }
}

static double calculate_log_cond(const std::string& allele, const std::string& call, const std::string& quals, int8_t max_qs) {  // CH:384-396
    double lc = 0;
    for (size_t i = 0; i < allele.size(); i++) {
        char c = call[i];
        int8_t q = (int8_t)std::min((int)max_qs, (int)(unsigned char)quals[i] - 33);
        if (q < 3) continue;
        if (allele[i] == c) lc += GT(0, q, 0);
        else lc += g_err[q][4];
    }
    return lc;
}

void CountsHelper::update_indel(const std::string& call, const std::string& quals, bool neg) {  // CH:253-304
    total++;
    int index = -1;
    for (size_t i = 0; i < alleles.size(); i++) if (alleles[i] == call) { index = (int)i; break; }
    int f = (int)java_round(het_prop * NF);  // 501 quirk, CH:256
    int n = (int)alleles.size();
    std::vector<double> lca(n, 0.0);
    int best = -1;
    for (int i = 0; i < n; i++) {
        if (alleles[i].size() == call.size()) {
            lca[i] = std::max(LOG_ERR_INDEL, calculate_log_cond(alleles[i], call, quals, max_base_qs));
            if (lca[i] > LOG_ERR_INDEL) {
                if (best == -1 || lca[best] < lca[i]) best = i;
            }
        } else {
            lca[i] = LOG_ERR_INDEL;
        }
    }
    if (index >= 0 && best >= 0 && best != index) index = std::min(index, best);
    else if (index < 0 && best >= 0) index = best;
    if (index >= 0) {
        counts[index]++;
        if (neg) strand[index][0]++; else strand[index][1]++;
    }
    int m = (int)logc.size();
    for (int i = 0; i < m; i++) {
        logc[i][i] += lca[i];
        for (int j = 0; j < m; j++) {
            if (i != j) {
                if (j == index) logc[i][j] += log_sum10(g_freq[f][0] + lca[index], g_freq[f][1] + LOG_ERR_INDEL);
                else if (i == index) logc[i][j] += log_sum10(g_freq[f][1] + lca[index], g_freq[f][0] + LOG_ERR_INDEL);
                else logc[i][j] += LOG_ERR_INDEL;
            }
        }
    }
}

void CountsHelper::posterior_from_events(std::vector<double>& ev) {  // CH:472-495
    double log_max = 1;
    for (double e : ev) if (log_max > 0 || log_max < e) log_max = e;
    double total = 0;
    for (double& e : ev) {
        e -= log_max;
        if (e < -20) e = 0.0;
        else e = std::pow(10.0, e);
        total += e;
    }
    for (double& e : ev) e = e / total;
}

std::vector<std::vector<double>> CountsHelper::posteriors(double het_rate) const {  // CH:410-443
    int n = (int)alleles.size();
    int hetero = n * (n - 1);
    double lp_het = std::log10(het_rate / hetero);
    double lp_hom = std::log10((1 - het_rate) / n);
    std::vector<double> ev((size_t)n * n);
    int k = 0;
    for (int i = 0; i < n; i++) {
        ev[k++] = logc[i][i] + lp_hom;
        for (int j = 0; j < n; j++) if (i != j) ev[k++] = logc[i][j] + lp_het;
    }
    posterior_from_events(ev);
    std::vector<std::vector<double>> post(n, std::vector<double>(n, 0.0));
    k = 0;
    for (int i = 0; i < n; i++) {
        post[i][i] = ev[k++];
        for (int j = 0; j < n; j++) if (i != j) post[i][j] = ev[k++];
    }
    return post;
}

int8_t CountsHelper::strand_bias_score(int i1, int i2) const {  // CH:563-576
    double p = fisher_pvalue(strand[i1][0], strand[i2][0], strand[i1][1], strand[i2][1]);
    return (int8_t)std::min((int)MAX_STRAND_BIAS_SCORE, (int)phred_score(p));
}

// ---------------------------------------------------------------------------------------------
// Called variants
// ---------------------------------------------------------------------------------------------
CalledVariant CalledVariant::make_snv(const GenomicVariant& snv, int8_t gt) {  // CalledSNV.java:78-82
    CalledVariant c;
    c.var = snv;
    c.called_snv = true;
    c.ref_cn = 2;
    c.total_cn = 0;  // CalledSNV.totalCopyNumber starts at 0 (:50)
    // setGenotype (:108-114): genotype field starts at UNDECIDED(-1)
    if (c.genotype != gt) {
        c.genotype = gt;
        c.update_cn_from_counts(c.total_cn);
    }
    return c;
}

CalledVariant CalledVariant::make_impl(const GenomicVariant& v, const std::vector<int8_t>& idx) {  // Impl:50-55
    CalledVariant c;
    c.var = v;
    c.called_snv = false;
    c.total_cn = 2;
    c.alleles_cn.assign(v.alleles.size(), 0);
    c.idx_called = idx;  // setIndexesCalledAlleles :288-292
    short t = (short)std::max((int)idx.size(), (int)c.total_cn);
    c.update_cn_from_counts(t);
    return c;
}

bool CalledVariant::undecided() const { return called_snv ? genotype == -1 : idx_called.empty(); }
bool CalledVariant::homozygous() const { return called_snv ? (genotype == 0 || genotype == 2) : idx_called.size() == 1; }
bool CalledVariant::hom_ref() const { return called_snv ? genotype == 0 : (idx_called.size() == 1 && idx_called[0] == 0); }
bool CalledVariant::heterozygous() const { return called_snv ? genotype == 1 : idx_called.size() > 1; }

std::vector<int8_t> CalledVariant::called_indexes() const {  // CalledSNV.java:388-400
    if (!called_snv) return idx_called;
    if (undecided()) return {};
    if (homozygous()) return {(int8_t)(genotype == 0 ? 0 : 1)};
    return {0, 1};
}

void CalledVariant::set_all_base_counts(const std::vector<int>& c) {  // CalledSNV.java:245-252
    for (int i = 0; i < 4; i++) cnt[i] = c[i];
    int sum = cnt[0] + cnt[1] + cnt[2] + cnt[3];
    if (sum > depth) depth = sum;
}

void CalledVariant::set_all_logc(const std::vector<std::vector<double>>& l) {  // CalledSNV.java:258-265
    int r = dna_index(var.alleles[0][0]), a = dna_index(var.alleles[1][0]);
    homo_ref = (float)l[r][r];
    homo_alt = (float)l[a][a];
    ref_alt = (float)l[r][a];
    alt_ref = (float)l[a][r];
}

int CalledVariant::report_count(const std::string& allele) const {  // VariantCallReport.java:22-27 (+CalledSNV.getCallReport :413-430)
    if (called_snv) {
        if (cnt[0] + cnt[1] + cnt[2] + cnt[3] <= 0) return 0;
        if (allele == var.alleles[0]) return cnt[dna_index(var.alleles[0][0])];
        if (allele == var.alleles[1]) return cnt[dna_index(var.alleles[1][0])];
        return 0;
    }
    if (!has_report || !rep_has_counts) return 0;
    // TreeMap allele->index: later duplicates overwrite earlier
    int idx = -1;
    for (size_t i = 0; i < rep_alleles.size(); i++) if (rep_alleles[i] == allele) idx = (int)i;
    return idx < 0 ? 0 : rep_counts[idx];
}

double CalledVariant::report_logc(const std::string& a1, const std::string& a2) const {  // VariantCallReport.java:28-34
    if (called_snv) {
        if ((float)(homo_ref + ref_alt + alt_ref + homo_alt) == 0.0f) return 0;  // CalledSNV.java:422
        int i = a1 == var.alleles[0] ? 0 : (a1 == var.alleles[1] ? 1 : -1);
        int j = a2 == var.alleles[0] ? 0 : (a2 == var.alleles[1] ? 1 : -1);
        if (i < 0 || j < 0) return 0;
        if (i == 0 && j == 0) return homo_ref;
        if (i == 0 && j == 1) return ref_alt;
        if (i == 1 && j == 0) return alt_ref;
        return homo_alt;
    }
    if (!has_report || !rep_has_logc) return 0;
    int i = -1, j = -1;
    for (size_t k = 0; k < rep_alleles.size(); k++) { if (rep_alleles[k] == a1) i = (int)k; if (rep_alleles[k] == a2) j = (int)k; }
    if (i < 0 || j < 0) return 0;
    return rep_logc[i][j];
}

void CalledVariant::update_cn_from_counts(short total) {
    if (called_snv) {  // CalledSNV.java:134-160
        total_cn = total;
        ref_cn = 0;
        if (undecided()) return;
        if (hom_ref()) { ref_cn = total; return; }
        else if (homozygous()) return;
        else if (total <= 2) { total_cn = 2; ref_cn = 1; return; }
        int r = dna_index(var.alleles[0][0]), a = dna_index(var.alleles[1][0]);
        double count_ref = cnt[r];
        double sum = cnt[r] + cnt[a];
        double prop = 0.5;
        if (sum > 0) prop = count_ref / sum;
        if (prop > 1) prop = 1;
        ref_cn = (short)java_round(prop * total);
        if (ref_cn == 0) ref_cn = 1;
        else if (ref_cn >= total) ref_cn = (short)(total - 1);
        return;
    }
    // CalledGenomicVariantImpl.java:228-286
    total_cn = total;
    std::fill(alleles_cn.begin(), alleles_cn.end(), (short)0);
    if (undecided()) return;
    if (hom_ref()) { alleles_cn[0] = total; return; }
    int nc = (int)idx_called.size();
    if (total <= nc) { for (int i = 0; i < nc; i++) alleles_cn[idx_called[i]] = 1; return; }
    if (!has_report || !rep_has_counts) {
        short def = (short)(total / nc);
        for (int i = 0; i < nc; i++) alleles_cn[idx_called[i]] = def;
        int diff = total - def * nc;
        alleles_cn[idx_called[0]] += (short)diff;
        return;
    }
    int total_reads = 0;
    std::vector<int> rc(nc);
    for (int i = 0; i < nc; i++) {
        rc[i] = report_count(var.alleles[idx_called[i]]);
        if (rc[i] == 0) rc[i] = 1;
        total_reads += rc[i];
    }
    int tc = 0;
    for (int i = 0; i < nc; i++) {
        int j = idx_called[i];
        alleles_cn[j] = (short)std::max(1L, java_round((double)total * rc[i] / total_reads));
        tc += alleles_cn[j];
    }
    if (tc < total) alleles_cn[idx_called[0]] += (short)(total - tc);
    else {
        int excess = tc - total;
        for (int i = nc - 1; excess > 0 && i >= 0; i--) {
            int j = idx_called[i];
            int rm = std::min(excess, alleles_cn[j] - 1);
            alleles_cn[j] -= (short)rm;
            excess -= rm;
        }
    }
}

void CalledVariant::make_undecided() {
    if (called_snv) { genotype = -1; gq = 0; ref_cn = 0; return; }  // CalledSNV.java:276-281
    idx_called.clear(); gq = 0; update_cn_from_counts(total_cn);   // Impl:318-323
}

// ---------------------------------------------------------------------------------------------
// VariantDiscoverySNVQAlgorithm
// ---------------------------------------------------------------------------------------------
static void indexes_max_genotype(const std::vector<std::vector<double>>& post, int idx_default, int& oi, int& oj) {  // VD:223-243
    int n = (int)post.size();
    if (idx_default < 0 || idx_default >= n) idx_default = 0;
    oi = oj = idx_default;
    double pmax = post[idx_default][idx_default];
    for (int i = 0; i < n; i++) {
        for (int j = i; j < (int)post[0].size(); j++) {
            double gp = post[i][j];
            if (i != j) gp += post[j][i];
            if (gp > pmax + 0.01) { pmax = gp; oi = i; oj = j; }
        }
    }
}
static std::vector<int> report_counts(const std::vector<int>& counts, const std::vector<int>& idx) {  // VD:244-251
    std::vector<int> r(idx.size());
    for (size_t i = 0; i < idx.size(); i++) r[i] = (idx[i] < 0 || idx[i] >= (int)counts.size()) ? 0 : counts[idx[i]];
    return r;
}
static std::vector<std::vector<double>> report_probs(const std::vector<std::vector<double>>& all, const std::vector<int>& idx) {  // VD:252-262
    size_t n = idx.size();
    std::vector<std::vector<double>> a(n, std::vector<double>(n, 0.0));
    for (size_t i = 0; i < n; i++) for (size_t j = 0; j < n; j++) {
        if (idx[i] < 0 || idx[i] >= (int)all.size()) a[i][j] = 0;
        else if (idx[j] < 0 || idx[j] >= (int)all[0].size()) a[i][j] = 0;
        else a[i][j] = all[idx[i]][idx[j]];
    }
    return a;
}
static GenomicVariant make_variant(const std::string& seq, int first, const std::vector<std::string>& alleles) {  // GenomicVariantImpl.java:99-106
    GenomicVariant v;
    v.seq = seq; v.first = first; v.len = (int)alleles[0].size(); v.last = first + v.len - 1;
    for (auto& a : alleles) v.alleles.push_back(to_upper(a));
    return v;
}
static GenomicVariant make_snv_variant(const std::string& seq, int pos, char ref, char alt) {  // SNV.java:41-46
    GenomicVariant v;
    v.seq = seq; v.first = pos; v.last = pos; v.len = 1;
    v.alleles = {std::string(1, ref), std::string(1, alt)};
    v.type = TYPE_BIALLELIC_SNV; v.snv_class = true;
    return v;
}

bool discover_snv(const CountsHelper& h, const std::string& seq, int pos, char ref, double het_rate, bool csb, CalledVariant& out) {  // VD:100-222
    if (h.total == 0) return false;
    int idx_ref = dna_index(ref);
    if (idx_ref < 0) return false;
    auto post = h.posteriors(het_rate);
    int ii, jj;
    indexes_max_genotype(post, idx_ref, ii, jj);
    double ref_prob = post[idx_ref][idx_ref];
    double max_p = post[ii][jj];
    if (ii != jj) max_p += post[jj][ii];
    short gq = phred_score(1 - max_p);
    int idx_alt = 0;
    if (ii != jj && ii != idx_ref && jj != idx_ref) {  // VD:128-177 (indexRef<0 unreachable here)
        int third;
        if (post[ii][ii] > post[jj][jj] + 0.01) { idx_alt = ii; third = jj; }
        else { idx_alt = jj; third = ii; }
        std::vector<std::string> alleles = {BASES_ARRAY[idx_ref], BASES_ARRAY[idx_alt], BASES_ARRAY[third]};
        std::vector<int> idx = {idx_ref, idx_alt, third};
        GenomicVariant gv = make_variant(seq, pos, alleles);
        gv.type = TYPE_MULTIALLELIC_SNV;
        gv.qs = phred_score(ref_prob);
        out = CalledVariant::make_impl(gv, {1, 2});
        out.gq = gq;
        out.depth = h.total;
        out.has_all_counts = true; out.all_counts = h.counts;
        out.has_report = true; out.rep_alleles = alleles; out.rep_counts = report_counts(h.counts, idx); out.rep_has_counts = true;
        out.rep_logc = report_probs(h.logc, idx); out.rep_has_logc = true;
        if (csb) out.strand_bias = h.strand_bias_score(idx_alt, third);
        return true;
    }
    int8_t genotype;
    if (ii != jj) {
        idx_alt = (idx_ref != ii) ? ii : jj;
        genotype = 1;
    } else if (idx_ref != ii) {
        idx_alt = ii;
        genotype = 2;
    } else {  // homozygous reference, VD:194-209
        std::vector<std::string> alleles = {BASES_ARRAY[idx_ref]};
        GenomicVariant gv = make_variant(seq, pos, alleles);
        gv.qs = phred_score(ref_prob);
        out = CalledVariant::make_impl(gv, {0});
        out.gq = gq;
        out.depth = h.total;
        out.has_all_counts = true; out.all_counts = h.counts;
        std::vector<int> idx = {idx_ref};
        out.has_report = true; out.rep_alleles = alleles; out.rep_counts = report_counts(h.counts, idx); out.rep_has_counts = true;
        out.rep_logc = report_probs(h.logc, idx); out.rep_has_logc = true;
        return true;
    }
    GenomicVariant snv = make_snv_variant(seq, pos, ref, "ACGT"[idx_alt]);
    snv.qs = phred_score(ref_prob);
    out = CalledVariant::make_snv(snv, genotype);
    out.gq = gq;
    out.depth = h.total;
    out.set_all_base_counts(h.counts);
    out.set_all_logc(h.logc);
    if (csb) out.strand_bias = h.strand_bias_score(idx_ref, idx_alt);
    return true;
}

CalledVariant genotype_snv(const GenomicVariant& v, const CountsHelper& h, double het_rate, bool csb) {  // VD:21-97
    if (h.total == 0) {
        CalledVariant u = CalledVariant::make_impl(v, {});
        u.has_all_counts = true; u.all_counts = h.counts;
        return u;
    }
    auto post = h.posteriors(het_rate);
    if (v.snv_class) {
        int r = dna_index(v.alleles[0][0]), a = dna_index(v.alleles[1][0]);
        double p_max = post[r][r];
        int8_t gt = 0;
        if (post[a][a] > p_max + 0.01) { p_max = post[a][a]; gt = 2; }
        double p_het = post[r][a] + post[a][r];
        if (p_het > p_max + 0.01) { p_max = p_het; gt = 1; }
        short gq = phred_score(1 - p_max);
        if (gq == 0) gt = -1;
        CalledVariant c = CalledVariant::make_snv(v, gt);
        c.gq = gq;
        c.depth = h.total;
        c.set_all_base_counts(h.counts);
        c.set_all_logc(h.logc);
        if (csb && gt != -1 && gt != 0) c.strand_bias = h.strand_bias_score(r, a);
        return c;
    }
    std::vector<int> idx(v.alleles.size());
    for (size_t i = 0; i < idx.size(); i++) idx[i] = dna_index(v.alleles[i][0]);
    auto rc = report_counts(h.counts, idx);
    auto rl = report_probs(h.logc, idx);
    auto rp = report_probs(post, idx);
    int mi, mj;
    indexes_max_genotype(rp, 0, mi, mj);
    double max_p = rp[mi][mj];
    std::vector<int8_t> gt;
    if (mi != mj) { max_p += rp[mj][mi]; gt = {(int8_t)mi, (int8_t)mj}; }
    else gt = {(int8_t)mi};
    short gq = phred_score(1 - max_p);
    CalledVariant c = CalledVariant::make_impl(v, gt);
    c.gq = gq;
    c.depth = h.total;
    c.has_all_counts = true; c.all_counts = h.counts;
    c.has_report = true; c.rep_alleles = v.alleles; c.rep_counts = rc; c.rep_has_counts = true; c.rep_logc = rl; c.rep_has_logc = true;
    if (csb) {
        if (gt.size() == 1 && gt[0] != 0) c.strand_bias = h.strand_bias_score(idx[0], idx[gt[0]]);
        else if (gt.size() == 2) c.strand_bias = h.strand_bias_score(idx[gt[0]], idx[gt[1]]);
    }
    return c;
}

bool call_indel(PileupRecord& p, const CountsHelper& h, const GenomicVariant* variant, double het_rate, bool csb, CalledVariant& out) {  // VD:265-361
    auto post = h.posteriors(het_rate);
    if (h.total == 0) {
        if (!variant) return false;
        out = CalledVariant::make_impl(*variant, {});
        return true;
    }
    int m0, m1;
    indexes_max_genotype(post, 0, m0, m1);
    std::vector<int8_t> called;
    GenomicVariant gv;
    std::vector<int> rcounts = h.counts;
    std::vector<std::vector<double>> rlogs = h.logc;
    if (!variant) {
        const auto& ha = h.alleles;
        std::vector<std::string> alleles = {ha[0]};
        std::vector<int> idx = {0};
        size_t ref_len = ha[0].size();
        bool length_change = false;
        if (m0 > 0 && m0 < (int)ha.size()) {
            alleles.push_back(ha[m0]); idx.push_back(m0);
            if (ha[m0].size() != ref_len) length_change = true;
        }
        if (m1 > 0 && m1 != m0 && m1 < (int)ha.size()) {
            alleles.push_back(ha[m1]); idx.push_back(m1);
            if (ha[m1].size() != ref_len) length_change = true;
            if (alleles.size() == 3 && ha[m1].size() != alleles[1].size()) length_change = true;
        }
        if (!length_change && !p.input_str()) return false;
        gv = make_variant(p.seq, p.pos, alleles);
        gv.type = p.str ? TYPE_STR : TYPE_INDEL;
        gv.qs = phred_score(post[0][0]);
        rcounts = report_counts(h.counts, idx);
        rlogs = report_probs(h.logc, idx);
        if (m1 != m0) {
            if (alleles.size() == 3) called = {1, 2};
            else called = {0, 1};
        } else {
            called = {(int8_t)(m0 == 0 ? 0 : 1)};
        }
    } else {
        gv = *variant;
        if (m0 > MAX_NUM_ALLELES || m1 > MAX_NUM_ALLELES) called = {};
        else if (m1 != m0) called = {(int8_t)m0, (int8_t)m1};
        else called = {(int8_t)m0};
    }
    int depth = h.total;
    out = CalledVariant::make_impl(gv, called);
    double max_p = post[m0][m1];
    if (m0 != m1) max_p += post[m1][m0];
    out.gq = phred_score(1 - max_p);
    out.depth = depth;
    if (depth > 0) { out.has_report = true; out.rep_alleles = gv.alleles; out.rep_counts = rcounts; out.rep_has_counts = true; out.rep_logc = rlogs; out.rep_has_logc = true; }
    if (csb) {
        if (called.size() == 1 && called[0] != 0) out.strand_bias = h.strand_bias_score(0, called[0]);
        else if (called.size() == 2) out.strand_bias = h.strand_bias_score(called[0], called[1]);
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// AlleleCallClustersBuilder
// ---------------------------------------------------------------------------------------------
static std::vector<double> het_posteriors(const std::vector<const PileupAlleleCall*>& calls, const std::string& consensus, int8_t max_qs) {  // ACCB:223-261
    size_t l = consensus.size();
    std::vector<double> ans(l, 0.0);
    for (size_t i = 0; i < l; i++) {
        char c = consensus[i];
        int idx_c = dna_index(c);
        if (idx_c < 0) continue;
        bool variable = false;
        for (size_t j = 0; j < calls.size() && !variable; j++) variable = calls[j]->allele()[i] != c;
        if (!variable) continue;
        CountsHelper helper;
        for (auto* call : calls) {
            std::string al = call->allele();
            int8_t q = (int8_t)std::min((int)max_qs, (int)(unsigned char)call->quals[i] - 33);
            helper.update(al.substr(i, 1), q, false);
        }
        auto post = helper.posteriors(0.001);
        for (int k = 0; k < 4; k++) {
            double hp = post[idx_c][k] + post[k][idx_c];
            if (k != idx_c && hp > ans[i]) ans[i] = hp;
        }
    }
    return ans;
}

static std::set<std::string> split_by_variant_sites(const std::vector<const PileupAlleleCall*>& calls, const std::string& consensus, int8_t max_qs) {  // ACCB:165-221
    std::set<std::string> answer;
    auto hp = het_posteriors(calls, consensus, max_qs);
    std::vector<int> sites;
    for (size_t i = 0; i < hp.size(); i++) if (hp[i] >= 0.51) sites.push_back((int)i);
    if (sites.empty()) { answer.insert(consensus); return answer; }
    size_t n = calls.size(), m = sites.size();
    std::vector<std::string> haps(n);
    for (size_t i = 0; i < n; i++) {
        std::string seq = calls[i]->allele();
        for (size_t j = 0; j < m; j++) haps[i].push_back(seq[sites[j]]);
    }
    std::map<std::string, int> hap_counts;  // CountsRankHelper<String> (TreeMap)
    for (auto& hps : haps) hap_counts[hps]++;
    int max_haps = 2;
    if (m > 3) max_haps = std::min(10, (int)m / 2 + 1);
    std::vector<std::pair<std::string, int>> ranked(hap_counts.begin(), hap_counts.end());
    std::stable_sort(ranked.begin(), ranked.end(), [](const auto& a, const auto& b) { return a.second > b.second; });
    std::vector<std::string> selected;
    for (int i = 0; i < max_haps && i < (int)ranked.size(); i++) selected.push_back(ranked[i].first);
    std::vector<std::vector<std::string>> sel_seqs(selected.size());
    for (size_t i = 0; i < n; i++) {
        auto it = std::find(selected.begin(), selected.end(), haps[i]);
        if (it != selected.end()) sel_seqs[it - selected.begin()].push_back(calls[i]->allele());
    }
    for (auto& seqs : sel_seqs) if (!seqs.empty()) answer.insert(hamming_consensus(seqs));
    return answer;
}

std::vector<std::string> cluster_allele_calls(const std::vector<PileupAlleleCall>& calls, const std::string& reference, int8_t max_qs) {  // ACCB:72-141
    std::set<std::string> alleles_set;
    std::map<int, std::vector<const PileupAlleleCall*>> by_len;  // HashMap in Java; iteration order is irrelevant (results go to a TreeSet)
    for (auto& c : calls) by_len[(int)c.allele().size()].push_back(&c);
    // filterLengthClusters ACCB:147-157
    std::map<int, std::vector<const PileupAlleleCall*>> filtered;
    if (by_len.size() < 3) filtered = by_len;
    else {
        double min_count = 0.2 * (double)calls.size();
        for (auto& kv : by_len) if (min_count <= (double)kv.second.size()) filtered[kv.first] = kv.second;
    }
    for (auto& kv : filtered) {
        int l = kv.first;
        auto& calls_l = kv.second;
        std::set<std::string> suggested;
        if (l == (int)reference.size()) suggested.insert(reference);
        std::set<std::string> length_alleles;
        if (calls_l.size() < 5 * suggested.size()) {
            length_alleles = suggested;
        } else {
            std::vector<std::string> alleles_l;
            for (auto* c : calls_l) alleles_l.push_back(c->allele());
            std::string consensus = hamming_consensus(alleles_l);
            suggested.insert(consensus);
            if (l < 4 || calls_l.size() < 10) length_alleles = suggested;
            else length_alleles = split_by_variant_sites(calls_l, consensus, max_qs);
        }
        alleles_set.insert(length_alleles.begin(), length_alleles.end());
    }
    alleles_set.insert(reference);
    std::vector<std::string> answer = {reference};
    for (auto& a : alleles_set) if (a != reference) answer.push_back(a);
    return answer;
}

}  // namespace ngsep_oracle
