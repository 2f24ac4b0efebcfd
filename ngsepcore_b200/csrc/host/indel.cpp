// Host indel / STR caller for event positions with a reference span > 1.  Restates
//   PileupRecord.getAlleleCalls(span)                    PR   = discovery/PileupRecord.java:126-152
//   AlleleCallClustersBuilder.clusterAlleleCalls          ACCB = discovery/AlleleCallClustersBuilder.java:72-261
//   CountsHelper.calculateCountsIndel / updateCountsIndel CH   = discovery/CountsHelper.java:96-105,253-304,384-396
//   VariantDiscoverySNVQAlgorithm.callIndel               VD   = discovery/VariantDiscoverySNVQAlgorithm.java:265-361
// (paths under /root/reference/src/ngsep/).  Events are ~1e-4 of positions, so this stays on the host in round 1
// (SURVEY.md section 8f rank 1 moves it to the device).
#include <algorithm>
#include <cmath>
#include <map>
#include <set>

#include "pipeline.hpp"

namespace ngsep_b200 {
namespace host {

long jround(double x) { double f = std::floor(x); return (long)((x - f >= 0.5) ? f + 1.0 : f); }

int phred(double p) {
    if (p == 0) return 255;
    double s = -10 * std::log10(p);
    if (s > 255) return 255;
    return (int)(short)jround(s);
}

double fisher_pvalue(int a, int b, int c, int d) {
    static std::vector<double> lf;
    auto exact = [&](int a, int b, int c, int d) {
        int n = a + b + c + d;
        if ((int)lf.size() <= n) {
            int m = std::max(n, 10000);
            lf.assign(m + 1, 0.0);
            for (int i = 2; i <= m; i++) lf[i] = lf[i - 1] + std::log10((double)i);
        }
        double v = lf[a + b];
        v += lf[c + d]; v += lf[a + c]; v += lf[b + d];
        v -= lf[a]; v -= lf[b]; v -= lf[c]; v -= lf[d]; v -= lf[n];
        return std::pow(10.0, v);
    };
    if (a > b) { std::swap(a, b); std::swap(c, d); }
    if (a > c) { std::swap(a, c); std::swap(b, d); }
    int e = std::min(a, d);
    double ans = 0;
    while (a >= 0 && d >= 0) {
        double p = exact(a, b, c, d);
        if (e >= 10 && ans > 100.0 * e * p) break;
        ans += p;
        a--; b++; c++; d--; e++;
    }
    return ans;
}

namespace {

struct QTables {   // the slices of CountsHelper's caches this caller needs
    double match[31], err4[31], lf251_0, lf251_1, log_err_indel;
    double a250[31], c250[31];
    QTables() {
        for (int q = 0; q < 31; q++) match[q] = err4[q] = a250[q] = c250[q] = 0;
        for (int q = 3; q < 31; q++) {
            double e = std::pow(10.0, -0.1 * q), s = 1 - e;
            match[q] = std::log10(s);                        // logProbCacheGT[0][q][0]
            err4[q] = -0.1 * q - std::log10(3.0);             // logProbCacheError[q][4]
            a250[q] = match[q];
            c250[q] = std::log10(0.5 * s + 0.5 * e / 3);
        }
        double af = 251.0 / 500.0;                            // f = round(0.5 * 501), CH:256
        lf251_0 = std::log10(af);
        lf251_1 = std::log10(1 - af);
        log_err_indel = std::log10(0.0001);
    }
};
const QTables& qt() { static const QTables t; return t; }

double log_sum(double a, double b) {  // math/LogMath.java:38-44
    if (a - b > 20) return a;
    if (b - a > 20) return b;
    return a + std::log10(1 + std::pow(10.0, b - a));
}

std::string column_consensus(const std::vector<std::string>& seqs) {
    std::string out;
    for (size_t i = 0; i < seqs[0].size(); i++) {
        int cnt[256] = {0};
        for (auto& s : seqs) cnt[(unsigned char)s[i]]++;
        int best = 0;
        for (int c = 1; c < 256; c++) if (cnt[c] > cnt[best]) best = c;
        out.push_back((char)best);
    }
    return out;
}

// 16-event posterior of CH:410-495 for n alleles
std::vector<std::vector<double>> posteriors(const std::vector<std::vector<double>>& L, double het_rate) {
    int n = (int)L.size();
    double lp_het = std::log10(het_rate / (n * (n - 1)));
    double lp_hom = std::log10((1 - het_rate) / n);
    std::vector<double> ev;
    for (int i = 0; i < n; i++) {
        ev.push_back(L[i][i] + lp_hom);
        for (int j = 0; j < n; j++) if (i != j) ev.push_back(L[i][j] + lp_het);
    }
    double mx = 1;
    for (double e : ev) if (mx > 0 || mx < e) mx = e;
    double tot = 0;
    for (double& e : ev) { e -= mx; e = (e < -20) ? 0.0 : std::pow(10.0, e); tot += e; }
    for (double& e : ev) e /= tot;
    std::vector<std::vector<double>> P(n, std::vector<double>(n, 0.0));
    int k = 0;
    for (int i = 0; i < n; i++) { P[i][i] = ev[k++]; for (int j = 0; j < n; j++) if (i != j) P[i][j] = ev[k++]; }
    return P;
}

// ACCB.calculateHetPosteriors (ACCB:223-261): a 4-allele SNV column per consensus position
double het_posterior_column(const std::vector<const SpanCall*>& calls, size_t i, int idx_c, int max_qs) {
    const QTables& t = qt();
    std::vector<std::vector<double>> L(4, std::vector<double>(4, 0.0));
    for (auto* c : calls) {
        int q = std::min(max_qs, (int)(unsigned char)c->quals[i] - 33);
        q = (int)(int8_t)q;
        if (q <= 3) continue;              // CH:214-216
        if (q > 30) q = 30;                // fresh CountsHelper: maxBaseQS 30 (CH:56)
        int b = -1;
        switch (c->allele[i]) { case 'A': b = 0; break; case 'C': b = 1; break; case 'G': b = 2; break; case 'T': b = 3; break; }
        if (b < 0) continue;
        for (int x = 0; x < 4; x++) {
            L[x][x] += (x == b) ? t.a250[q] : t.err4[q];
            for (int y = 0; y < 4; y++) if (x != y) L[x][y] += (y == b || x == b) ? t.c250[q] : t.err4[q];
        }
    }
    auto P = posteriors(L, 0.001);
    double best = 0;
    for (int k = 0; k < 4; k++) { double hp = P[idx_c][k] + P[k][idx_c]; if (k != idx_c && hp > best) best = hp; }
    return best;
}

std::set<std::string> split_by_variant_sites(const std::vector<const SpanCall*>& calls, const std::string& consensus, int max_qs) {  // ACCB:165-221
    std::set<std::string> out;
    std::vector<int> sites;
    for (size_t i = 0; i < consensus.size(); i++) {
        int idx_c = (consensus[i] < 'a') ? dna_index(consensus[i]) : -1;
        if (idx_c < 0) continue;
        bool variable = false;
        for (auto* c : calls) if (c->allele[i] != consensus[i]) { variable = true; break; }
        if (!variable) continue;
        if (het_posterior_column(calls, i, idx_c, max_qs) >= 0.51) sites.push_back((int)i);
    }
    if (sites.empty()) { out.insert(consensus); return out; }
    std::vector<std::string> haps(calls.size());
    for (size_t i = 0; i < calls.size(); i++) for (int s : sites) haps[i].push_back(calls[i]->allele[s]);
    std::map<std::string, int> counts;
    for (auto& h : haps) counts[h]++;
    std::vector<std::pair<std::string, int>> ranked(counts.begin(), counts.end());
    std::stable_sort(ranked.begin(), ranked.end(), [](const auto& a, const auto& b) { return a.second > b.second; });
    int max_haps = 2;
    if (sites.size() > 3) max_haps = std::min(10, (int)sites.size() / 2 + 1);
    std::vector<std::string> chosen;
    for (int i = 0; i < max_haps && i < (int)ranked.size(); i++) chosen.push_back(ranked[i].first);
    std::vector<std::vector<std::string>> groups(chosen.size());
    for (size_t i = 0; i < calls.size(); i++) {
        auto it = std::find(chosen.begin(), chosen.end(), haps[i]);
        if (it != chosen.end()) groups[it - chosen.begin()].push_back(calls[i]->allele);
    }
    for (auto& g : groups) if (!g.empty()) out.insert(column_consensus(g));
    return out;
}

std::vector<std::string> cluster_alleles(const std::vector<SpanCall>& calls, const std::string& ref, int max_qs) {  // ACCB:72-141
    std::map<int, std::vector<const SpanCall*>> by_len;
    for (auto& c : calls) by_len[(int)c.allele.size()].push_back(&c);
    if (by_len.size() >= 3) {  // filterLengthClusters, ACCB:147-157
        double min_count = 0.2 * (double)calls.size();
        for (auto it = by_len.begin(); it != by_len.end();) { if (min_count <= (double)it->second.size()) ++it; else it = by_len.erase(it); }
    }
    std::set<std::string> all;
    for (auto& kv : by_len) {
        auto& cl = kv.second;
        std::set<std::string> suggested;
        if (kv.first == (int)ref.size()) suggested.insert(ref);
        if (cl.size() < 5 * suggested.size()) { all.insert(suggested.begin(), suggested.end()); continue; }
        std::vector<std::string> seqs;
        for (auto* c : cl) seqs.push_back(c->allele);
        std::string consensus = column_consensus(seqs);
        suggested.insert(consensus);
        if (kv.first < 4 || cl.size() < 10) all.insert(suggested.begin(), suggested.end());
        else { auto s = split_by_variant_sites(cl, consensus, max_qs); all.insert(s.begin(), s.end()); }
    }
    all.insert(ref);
    std::vector<std::string> out = {ref};
    for (auto& a : all) if (a != ref) out.push_back(a);
    return out;
}

}  // namespace

void gather_span_event(int pos, const std::string& ref_allele, bool is_str, bool input_str, std::vector<HostRead*>& pileup, int max_base_qs, SpanEventInput& out) {
    const int span = (int)ref_allele.size();
    out.pos = pos; out.ref_allele = ref_allele; out.is_str = is_str; out.input_str = input_str;
    out.calls.clear();
    for (HostRead* r : pileup) {
        SpanCall c;
        if (!r->span_call(pos, pos + span - 1, c.allele, c.quals)) continue;
        c.allele = upper(c.allele);
        c.neg = r->neg;
        out.calls.push_back(std::move(c));
    }
    out.alleles = cluster_alleles(out.calls, ref_allele, max_base_qs);
}

// The per-call part of CountsHelper.updateCountsIndel (CH:253-289) on the host: what k_indel_tallies computes on the device for the
// product's single-sample runs; this copy serves the sample axis, whose per-sample genotypes stay on the host, and the host-only
// test hook.
void tally_calls_host(const SpanEventInput& in, const IndelCallerOptions& opt, CallTallies& ct) {
    const QTables& q = qt();
    const std::vector<std::string>& alleles = in.alleles;
    const int n = (int)alleles.size();
    ct = CallTallies();
    ct.counts.assign(n, 0);
    ct.strand.assign(n, {0, 0});
    ct.diag.assign(n, 0.0);
    for (auto& c : in.calls) {
        ct.total++;
        int index = -1;
        for (int i = 0; i < n; i++) if (alleles[i] == c.allele) { index = i; break; }
        std::vector<double> lc(n, q.log_err_indel);
        int best = -1;
        for (int i = 0; i < n; i++) {
            if (alleles[i].size() != c.allele.size()) continue;
            double v = 0;
            for (size_t k = 0; k < alleles[i].size(); k++) {  // calculateLogCond, CH:384-396
                int qs = (int)(int8_t)std::min(opt.max_base_qs, (int)(unsigned char)c.quals[k] - 33);
                if (qs < 3) continue;
                v += (alleles[i][k] == c.allele[k]) ? q.match[qs] : q.err4[qs];
            }
            lc[i] = std::max(q.log_err_indel, v);
            if (lc[i] > q.log_err_indel && (best == -1 || lc[best] < lc[i])) best = i;
        }
        if (index >= 0 && best >= 0 && best != index) index = std::min(index, best);
        else if (index < 0 && best >= 0) index = best;
        if (index >= 0) { ct.counts[index]++; ct.strand[index][c.neg ? 0 : 1]++; }
        for (int i = 0; i < n; i++) ct.diag[i] += lc[i];
        ct.call_index.push_back(index);
        ct.call_fit.push_back(index >= 0 ? lc[index] : q.log_err_indel);
    }
}

// The rest of updateCountsIndel (the heterozygous mixture terms of the off-diagonal entries, CH:290-300: one log-sum per call and
// direction), getPosteriorProbabilities (CH:410-495) and VD.getIndexesMaxGenotype (VD:223-243) — the part that hangs on libm.
void finish_tally(const CallTallies& ct, size_t n_alleles, double het_rate, IndelTally& t) {
    const QTables& q = qt();
    const int n = (int)n_alleles;
    t = IndelTally();
    t.total = ct.total;
    t.counts = ct.counts;
    t.strand = ct.strand;
    t.L.assign(n, std::vector<double>(n, 0.0));
    for (int i = 0; i < n; i++) t.L[i][i] = ct.diag[i];
    for (size_t c = 0; c < ct.call_index.size(); c++) {
        const int index = ct.call_index[c];
        double t_to = 0, t_from = 0;
        if (index >= 0) {
            t_to = log_sum(q.lf251_0 + ct.call_fit[c], q.lf251_1 + q.log_err_indel);
            t_from = log_sum(q.lf251_1 + ct.call_fit[c], q.lf251_0 + q.log_err_indel);
        }
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                if (i == j) continue;
                if (j == index) t.L[i][j] += t_to;
                else if (i == index) t.L[i][j] += t_from;
                else t.L[i][j] += q.log_err_indel;
            }
    }
    auto P = posteriors(t.L, het_rate);
    t.p00 = P[0][0];
    int m0 = 0, m1 = 0;
    double pmax = P[0][0];
    for (int i = 0; i < n; i++) for (int j = i; j < n; j++) {
        double gp = P[i][j];
        if (i != j) gp += P[j][i];
        if (gp > pmax + 0.01) { pmax = gp; m0 = i; m1 = j; }
    }
    t.m0 = m0; t.m1 = m1;
    t.max_p = P[m0][m1];
    if (m0 != m1) t.max_p += P[m1][m0];
}

void tally_span_event_host(const SpanEventInput& in, const IndelCallerOptions& opt, IndelTally& t) {
    CallTallies ct;
    tally_calls_host(in, opt, ct);
    finish_tally(ct, in.alleles.size(), opt.het_rate, t);
}

double indel_log_error() { return qt().log_err_indel; }

// callIndel, VD:265-361 (variant == null) from the tallies, then the filter of SSVPL:263
bool finish_indel_call(const SpanEventInput& in, const IndelTally& t, const IndelCallerOptions& opt, IndelRecord& out) {
    const std::vector<std::string>& alleles = in.alleles;
    if (t.total == 0) return false;
    const int m0 = t.m0, m1 = t.m1;
    std::vector<int> idx = {0};
    std::vector<std::string> va = {alleles[0]};
    bool length_change = false;
    if (m0 > 0) { va.push_back(alleles[m0]); idx.push_back(m0); if (alleles[m0].size() != alleles[0].size()) length_change = true; }
    if (m1 > 0 && m1 != m0) {
        va.push_back(alleles[m1]); idx.push_back(m1);
        if (alleles[m1].size() != alleles[0].size()) length_change = true;
        if (va.size() == 3 && alleles[m1].size() != va[1].size()) length_change = true;
    }
    if (!length_change && !in.input_str) return false;   // VD:307
    out = IndelRecord();
    out.pos = in.pos;
    out.last = in.pos + (int)in.ref_allele.size() - 1;
    out.alleles = va;
    out.str = in.is_str;
    out.qual = phred(t.p00);
    for (int i : idx) out.counts.push_back(t.counts[i]);
    out.logc.assign(idx.size(), std::vector<double>(idx.size(), 0.0));
    for (size_t i = 0; i < idx.size(); i++) for (size_t j = 0; j < idx.size(); j++) out.logc[i][j] = t.L[idx[i]][idx[j]];
    if (m1 != m0) out.called = (va.size() == 3) ? std::vector<int>{1, 2} : std::vector<int>{0, 1};
    else out.called = {m0 == 0 ? 0 : 1};
    out.gq = phred(1 - t.max_p);
    out.depth = t.total;
    if (opt.csb) {
        // CountsHelper.getScoreStrandBiasFisher is indexed with the positions in calledAlleles (VD:356-357), i.e. into the
        // helper's allele list, not the reduced list — restated as is.
        auto score = [&](int i1, int i2) {
            double p = fisher_pvalue(t.strand[i1][0], t.strand[i2][0], t.strand[i1][1], t.strand[i2][1]);
            return std::min(100, phred(p));
        };
        if (out.called.size() == 1 && out.called[0] != 0) out.strand_bias = score(0, out.called[0]);
        else if (out.called.size() == 2) out.strand_bias = score(out.called[0], out.called[1]);
    }
    // filter of SSVPL:263: undecided / homozygous reference / low quality
    bool hom_ref = out.called.size() == 1 && out.called[0] == 0;
    if (hom_ref || opt.min_quality > out.gq) return false;
    return true;
}

bool call_indel_at(int pos, const std::string& ref_allele, bool is_str, std::vector<HostRead*>& pileup, const IndelCallerOptions& opt, IndelRecord& out, bool input_str) {
    SpanEventInput in;
    gather_span_event(pos, ref_allele, is_str, input_str, pileup, opt.max_base_qs, in);
    if (opt.ploidy >= 3) {   // pools / polyploids, SSVPL:281-295
        CallTallies ct;
        tally_calls_host(in, opt, ct);
        return finish_indel_pool(in, ct, opt, opt.ploidy, out);
    }
    IndelTally t;
    tally_span_event_host(in, opt, t);
    return finish_indel_call(in, t, opt, out);
}

// ---- sample axis --------------------------------------------------------------------------------------------------
namespace {

// CountsHelper.calculateCountsIndel over a call list (CH:96-105, 253-304): counts, total and the log-likelihood matrix.
struct IndelCounts { std::vector<int> counts; std::vector<std::vector<double>> L; int total = 0; };

IndelCounts indel_counts(const std::vector<std::string>& alleles, const std::vector<const SpanCall*>& calls, int max_base_qs) {
    const QTables& t = qt();
    const int n = (int)alleles.size();
    IndelCounts h;
    h.counts.assign(n, 0);
    h.L.assign(n, std::vector<double>(n, 0.0));
    for (const SpanCall* cp : calls) {
        const SpanCall& c = *cp;
        h.total++;
        int index = -1;
        for (int i = 0; i < n; i++) if (alleles[i] == c.allele) { index = i; break; }
        std::vector<double> lc(n, t.log_err_indel);
        int best = -1;
        for (int i = 0; i < n; i++) {
            if (alleles[i].size() != c.allele.size()) continue;
            double v = 0;
            for (size_t k = 0; k < alleles[i].size(); k++) {  // calculateLogCond, CH:384-396
                int q = (int)(int8_t)std::min(max_base_qs, (int)(unsigned char)c.quals[k] - 33);
                if (q < 3) continue;
                v += (alleles[i][k] == c.allele[k]) ? t.match[q] : t.err4[q];
            }
            lc[i] = std::max(t.log_err_indel, v);
            if (lc[i] > t.log_err_indel && (best == -1 || lc[best] < lc[i])) best = i;
        }
        if (index >= 0 && best >= 0 && best != index) index = std::min(index, best);
        else if (index < 0 && best >= 0) index = best;
        if (index >= 0) h.counts[index]++;
        for (int i = 0; i < n; i++) {
            h.L[i][i] += lc[i];
            for (int j = 0; j < n; j++) {
                if (i == j) continue;
                if (j == index) h.L[i][j] += log_sum(t.lf251_0 + lc[index], t.lf251_1 + t.log_err_indel);
                else if (i == index) h.L[i][j] += log_sum(t.lf251_1 + lc[index], t.lf251_0 + t.log_err_indel);
                else h.L[i][j] += t.log_err_indel;
            }
        }
    }
    return h;
}

// SSVPL.genotypeVariantSample for a variant that is not an SNV (SSVPL:378-384, 386-388): VD.callIndel with the variant
// given (VD:265-361) on the sample's calls, then the sample-level minQuality (40: the listener is a fresh one, MSVD:676-678).
SampleIndelCall genotype_indel_sample(const std::vector<std::string>& alleles, const std::vector<const SpanCall*>& calls, int max_base_qs, double het_rate,
                                      int sample_min_quality = 40) {
    SampleIndelCall sc;
    IndelCounts h = indel_counts(alleles, calls, max_base_qs);
    if (h.total == 0) return sc;                       // VD:273-276: undecided, no report
    auto P = posteriors(h.L, het_rate);
    const int n = (int)alleles.size();
    int m0 = 0, m1 = 0;
    double pmax = P[0][0];
    for (int i = 0; i < n; i++) for (int j = i; j < n; j++) {
        double gp = P[i][j];
        if (i != j) gp += P[j][i];
        if (gp > pmax + 0.01) { pmax = gp; m0 = i; m1 = j; }
    }
    if (m0 != m1) sc.called = {m0, m1}; else sc.called = {m0};
    double max_p = P[m0][m1];
    if (m0 != m1) max_p += P[m1][m0];
    sc.gq = phred(1 - max_p);
    sc.depth = h.total;
    sc.has_report = true;
    sc.counts = h.counts;
    sc.logc = h.L;
    if (sample_min_quality > sc.gq) { sc.called.clear(); sc.gq = 0; }   // SSVPL:388 makeUndecided (the report stays)
    return sc;
}

}  // namespace

SampleIndelCall genotype_known_indel_at(int pos, const std::vector<std::string>& alleles, std::vector<HostRead*>& pileup, const IndelCallerOptions& opt) {
    const int span = (int)alleles[0].size();
    std::vector<SpanCall> calls;
    for (HostRead* r : pileup) {
        SpanCall c;
        if (!r->span_call(pos, pos + span - 1, c.allele, c.quals)) continue;
        c.allele = upper(c.allele);
        c.neg = r->neg;
        calls.push_back(std::move(c));
    }
    std::vector<const SpanCall*> ptrs;
    for (auto& c : calls) ptrs.push_back(&c);
    return genotype_indel_sample(alleles, ptrs, opt.max_base_qs, opt.het_rate, opt.min_quality);
}

int call_population_indel_at(int pos, const std::string& ref_allele, bool is_str, std::vector<HostRead*>& pileup, const std::vector<SampleDef>& samples,
                             const IndelCallerOptions& opt, PopIndelRecord& rec) {
    const int span = (int)ref_allele.size();
    // PR.getAlleleCalls(span): every read of the column, pileup order (pooled discovery, MSVD:600)
    std::vector<SpanCall> calls;
    std::vector<int> call_rg;
    for (HostRead* r : pileup) {
        SpanCall c;
        if (!r->span_call(pos, pos + span - 1, c.allele, c.quals)) continue;
        c.allele = upper(c.allele);
        c.neg = r->neg;
        calls.push_back(std::move(c));
        call_rg.push_back(r->rg);
    }
    std::vector<std::string> alleles = cluster_alleles(calls, ref_allele, opt.max_base_qs);   // MSVD:601-602
    // createIndelVariantPool (SSVPL:333-338): the pooled helper is only asked for its allele list and total count
    if (alleles.size() == 1 || calls.empty()) return 0;
    // the sample's calls: PR.getAlleleCalls(span, sample.getReadGroups()) — read group by read group, pileup order inside (PR:104-111)
    std::vector<std::vector<const SpanCall*>> per_sample(samples.size());
    for (size_t s = 0; s < samples.size(); s++)
        for (int rg : samples[s].rgs)
            for (size_t k = 0; k < calls.size(); k++) if (call_rg[k] == rg) per_sample[s].push_back(&calls[k]);
    auto genotype_all = [&](const std::vector<std::string>& al, std::vector<SampleIndelCall>& out) {   // MSVD:674-693
        out.clear();
        int qs = 0;
        for (size_t s = 0; s < samples.size(); s++) {
            out.push_back(genotype_indel_sample(al, per_sample[s], opt.max_base_qs, opt.het_rate));
            const SampleIndelCall& c = out.back();
            const bool hom_ref = c.called.size() == 1 && c.called[0] == 0;
            if (!c.called.empty() && !hom_ref && c.gq > qs) qs = c.gq;
        }
        return qs;
    };
    std::vector<SampleIndelCall> gc;
    while (alleles.size() > 2) {   // discoverPopulationIndel, MSVD:623-630
        bool same_len = true;
        for (auto& a : alleles) if (a.size() != alleles[0].size()) same_len = false;
        if (same_len) return 0;                                   // (no input STRs in discovery mode)
        int qs = genotype_all(alleles, gc);
        if (qs < opt.min_quality) return 0;
        std::set<std::string> called = {alleles[0]};              // makeNewVariant, MSVD:642-656 (TreeSet)
        for (auto& c : gc) for (int k : c.called) called.insert(alleles[k]);
        if (called.size() == alleles.size()) break;
        std::vector<std::string> na = {alleles[0]};               // SSVPL:346-359: reference first, then set order
        for (auto& a : called) if (a != alleles[0]) na.push_back(a);
        alleles.swap(na);
    }
    rec = PopIndelRecord();
    rec.pos = pos;
    rec.last = pos + span - 1;
    rec.alleles = alleles;
    rec.str = is_str;
    rec.qs = genotype_all(alleles, rec.calls);                    // MSVD:532
    if (rec.qs == 0 || rec.qs < opt.min_quality) return 1;        // MSVD:533
    return 2;
}

}  // namespace host
}  // namespace ngsep_b200
