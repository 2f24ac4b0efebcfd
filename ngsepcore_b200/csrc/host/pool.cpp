// Pools / polyploids: SingleSampleVariantPileupListener with ploidy >= DEF_MIN_PLOIDY_POOL_ALGORITHM = 3 (discovery mode).
//   discoverSNV's pool branch            SSVPL:240-254     createSNVVariantPool SSVPL:297-332
//   discoverIndel's pool branch          SSVPL:281-295     createIndelVariantPool SSVPL:333-338
//   genotypeVariantPool                  SSVPL:402-503     (one CountsHelper per allele-frequency hypothesis k / ploidy < 0.51)
//   CountsHelper.updateCounts with an allele list and a heterozygous proportion   CH:86-95, 135-187, 209-251
//   CountsHelper.getPosteriorProbabilities(h, major)                               CH:450-467
// The device finds the SNV columns whose alternative allele has the support createSNVVariantPool asks for and hands over their
// counted bases in pileup order (k_pool_columns, NGSEP_CALL_POOL); the indel / STR events come with the device's allele tallies
// (k_indel_tallies).  What is done here is what hangs on libm and on the allele list: the helpers' table sums for every hypothesis,
// the posteriors, the genotype, the second round with the reduced allele list.
#include <algorithm>
#include <cmath>

#include "pipeline.hpp"

namespace ngsep_b200 {
namespace host {

namespace {

double log_sum10(double a, double b) {  // math/LogMath.java:38-44
    if (a - b > 20) return a;
    if (b - a > 20) return b;
    return a + std::log10(1 + std::pow(10.0, b - a));
}

// CountsHelper.calculatePosteriorProbabilities (CH:472-495) in place
void posterior_from_events(std::vector<double>& ev) {
    double mx = 1;
    for (double e : ev) if (mx > 0 || mx < e) mx = e;
    double tot = 0;
    for (double& e : ev) { e -= mx; e = (e < -20) ? 0.0 : std::pow(10.0, e); tot += e; }
    for (double& e : ev) e /= tot;
}

struct Helpers {                              // the CountsHelpers of genotypeVariantPool, one per allele-frequency hypothesis
    std::vector<double> freqs;
    std::vector<int> counts;                  // of the first helper (the counts do not depend on the hypothesis)
    int total = 0;
    std::vector<std::vector<std::vector<double>>> L;   // [hypothesis][i][j]
};

std::vector<double> hypotheses(int haplotypes) {   // SSVPL:410-416: the loop variable itself accumulates
    std::vector<double> f;
    const double step = 1.0 / (double)haplotypes;
    for (double freq = step; freq < 0.51; freq += step) f.push_back(freq);
    return f;
}

// CountsHelper.calculateCountsGTSNV (CH:86-95) + updateCounts (CH:209-251) over the counted bases of a column, in pileup order, for the
// allele list A (base indexes) and every hypothesis.  ev: quality << 2 | base, the quality already min(30, q), > 3 and capped at maxBaseQS.
Helpers snv_helpers(const uint8_t* ev, int n_ev, int depth, const std::vector<int>& A, int haplotypes) {
    Helpers H;
    H.freqs = hypotheses(haplotypes);
    const int n = (int)A.size();
    H.counts.assign((size_t)n, 0);
    H.total = depth;
    H.L.assign(H.freqs.size(), std::vector<std::vector<double>>((size_t)n, std::vector<double>((size_t)n, 0.0)));
    int index_of[4] = {-1, -1, -1, -1};
    for (int i = n - 1; i >= 0; i--) index_of[A[(size_t)i]] = i;   // List.indexOf: the first one
    // the table slices this allele count needs (CH:135-187): error towards one of n - 1 other alleles, a match, the mixtures
    double err[31], match[31];
    std::vector<std::array<double, 31>> mix_f(H.freqs.size()), mix_g(H.freqs.size());
    for (int q = 3; q <= 30; q++) {
        const double e = std::pow(10.0, -0.1 * q), s = 1 - e;
        err[q] = -0.1 * q - std::log10((double)(n - 1));
        match[q] = std::log10(s);
        for (size_t j = 0; j < H.freqs.size(); j++) {
            const int f = (int)jround(H.freqs[j] * 500), g = (int)jround((1 - H.freqs[j]) * 500);   // CH:212-213
            const double af = (double)f / 500, ag = (double)g / 500;
            mix_f[j][(size_t)q] = std::log10(af * s + (1 - af) * e / (n - 1));
            mix_g[j][(size_t)q] = std::log10(ag * s + (1 - ag) * e / (n - 1));
        }
    }
    for (int k = 0; k < n_ev; k++) {
        const int q = ev[k] >> 2, index = index_of[ev[k] & 3];
        if (index < 0) continue;                                      // CH:224: a base outside the allele list adds nothing
        H.counts[(size_t)index]++;
        for (size_t j = 0; j < H.freqs.size(); j++) {
            auto& L = H.L[j];
            for (int i = 0; i < n; i++) {
                L[(size_t)i][(size_t)i] += (i == index) ? match[q] : err[q];
                for (int c = 0; c < n; c++) {
                    if (i == c) continue;
                    if (c == index) L[(size_t)i][(size_t)c] += mix_f[j][(size_t)q];
                    else if (i == index) L[(size_t)i][(size_t)c] += mix_g[j][(size_t)q];
                    else L[(size_t)i][(size_t)c] += err[q];
                }
            }
        }
    }
    return H;
}

// CountsHelper.calculateCountsIndel with the hypothesis' heterozygous proportion (CH:96-105, 253-304) from the per-call tallies
Helpers indel_helpers(const CallTallies& ct, size_t n_alleles, int haplotypes) {
    Helpers H;
    H.freqs = hypotheses(haplotypes);
    const int n = (int)n_alleles;
    H.counts = ct.counts;
    H.total = ct.total;
    const double lerr = indel_log_error();
    for (double freq : H.freqs) {
        const int f = (int)jround(freq * 501);                       // CH:256: DEF_NUM_FREQUENCIES, not one less
        const double af = (double)f / 500;
        const double lf0 = std::log10(af), lf1 = std::log10(1 - af);
        std::vector<std::vector<double>> L((size_t)n, std::vector<double>((size_t)n, 0.0));
        for (int i = 0; i < n; i++) L[(size_t)i][(size_t)i] = ct.diag[(size_t)i];
        for (size_t c = 0; c < ct.call_index.size(); c++) {
            const int index = ct.call_index[c];
            double t_to = 0, t_from = 0;
            if (index >= 0) {
                t_to = log_sum10(lf0 + ct.call_fit[c], lf1 + lerr);
                t_from = log_sum10(lf1 + ct.call_fit[c], lf0 + lerr);
            }
            for (int i = 0; i < n; i++)
                for (int j = 0; j < n; j++) {
                    if (i == j) continue;
                    if (j == index) L[(size_t)i][(size_t)j] += t_to;
                    else if (i == index) L[(size_t)i][(size_t)j] += t_from;
                    else L[(size_t)i][(size_t)j] += lerr;
                }
        }
        H.L.push_back(std::move(L));
    }
    return H;
}

// genotypeVariantPool, SSVPL:402-503
PoolCall genotype_pool(const Helpers& H, int haplotypes, double h) {
    PoolCall call;
    const int n = (int)H.counts.size();
    int major = -1;
    for (int i = 0; i < n; i++) if (major == -1 || H.counts[(size_t)major] < H.counts[(size_t)i]) major = i;
    call.counts = H.counts;
    if (H.counts[(size_t)major] < haplotypes) { call.undecided = true; return call; }   // SSVPL:432-436
    const double lp_het = std::log10(h), lp_hom = std::log10(1 - h);
    std::vector<double> terms(H.freqs.size() + 1);
    const double term_hom = H.L[0][(size_t)major][(size_t)major] + lp_hom;
    double max_het = 0, min_hom = 1;
    int max_freq_idx = 0, max_alt = -1;
    for (int i = 0; i < n; i++) {
        if (i == major) continue;
        terms[0] = term_hom;
        for (size_t j = 0; j < H.freqs.size(); j++) terms[j + 1] = H.L[j][(size_t)major][(size_t)i] + lp_het;
        posterior_from_events(terms);
        int idx_max = -1;
        for (size_t k = 0; k < terms.size(); k++) if (idx_max == -1 || terms[(size_t)idx_max] < terms[k]) idx_max = (int)k;
        if (idx_max == 0) min_hom = std::min(min_hom, terms[0]);
        else if (max_alt == -1 || max_het < terms[(size_t)idx_max]) { max_het = terms[(size_t)idx_max]; max_freq_idx = idx_max - 1; max_alt = i; }
    }
    call.called = {major};
    if (max_alt != -1) call.called.push_back(max_alt);
    std::sort(call.called.begin(), call.called.end());
    call.depth = H.total;
    if (max_alt == -1) {
        call.gq = phred(1 - min_hom);
        call.logc = H.L[0];
    } else {
        const auto& L = H.L[(size_t)max_freq_idx];
        std::vector<double> ev((size_t)n);                          // getPosteriorProbabilities(h, major), CH:450-467
        const double lph = std::log10(h / (n - 1)), lpo = std::log10(1 - h);
        for (int j = 0; j < n; j++) ev[(size_t)j] = L[(size_t)major][(size_t)j] + (j == major ? lpo : lph);
        posterior_from_events(ev);
        call.gq = phred(1 - ev[(size_t)max_alt]);
        call.logc = L;
    }
    return call;
}

bool same_length(const std::vector<std::string>& a) { for (auto& s : a) if (s.size() != a[0].size()) return false; return true; }

}  // namespace

// discoverSNV for a pool (SSVPL:240-254) on a column the device found supported.  Returns false when no variant is called.
bool discover_pool_snv(const ngsep_snv_call& col, const uint8_t* ev, int ploidy, double het_rate, int min_quality, std::vector<int>& alleles, PoolCall& out) {
    const int ref = col.ref;
    const int sum = col.count[0] + col.count[1] + col.count[2] + col.count[3];
    const double min_count = std::max(1.0, (0.5 / (double)ploidy) * sum);        // createSNVVariantPool, SSVPL:306-319
    alleles = {ref};
    for (int b = 0; b < 4; b++) if (b != ref && (double)col.count[b] >= min_count) alleles.push_back(b);
    if (alleles.size() < 2) return false;
    const int n_ev = col.sample;
    PoolCall call = genotype_pool(snv_helpers(ev, n_ev, col.depth, alleles, ploidy), ploidy, het_rate);
    if (alleles.size() > 2) {                                                      // SSVPL:244-253
        const bool hom_ref = call.called.size() == 1 && call.called[0] == 0;
        if (call.undecided || hom_ref) return false;
        if (!(call.called.size() == 2 && call.called[0] != 0)) {
            std::vector<int> reduced = {ref};                                      // the reference, then the called alleles
            for (int i : call.called) if (alleles[(size_t)i] != ref) reduced.push_back(alleles[(size_t)i]);
            alleles = reduced;
            call = genotype_pool(snv_helpers(ev, n_ev, col.depth, alleles, ploidy), ploidy, het_rate);
        }
    }
    // discoverVariant's filter, SSVPL:223
    const bool hom_ref = call.called.size() == 1 && call.called[0] == 0;
    if (call.undecided || hom_ref || min_quality > call.gq) return false;
    out = std::move(call);
    return true;
}

// discoverIndel for a pool (SSVPL:281-295) + the filter of SSVPL:263.  `retally`: the allele tallies of the event for another allele list.
bool finish_indel_pool(const SpanEventInput& in, const CallTallies& ct, const IndelCallerOptions& opt, int ploidy, IndelRecord& out) {
    std::vector<std::string> alleles = in.alleles;
    if (alleles.size() == 1 || ct.total == 0) return false;                        // createIndelVariantPool, SSVPL:333-338
    if (same_length(alleles)) return false;
    PoolCall call = genotype_pool(indel_helpers(ct, alleles.size(), ploidy), ploidy, opt.het_rate);
    if (alleles.size() > 2) {
        const bool hom_ref = call.called.size() == 1 && call.called[0] == 0;
        if (call.undecided || hom_ref) return false;
        if (!(call.called.size() == 2 && call.called[0] != 0)) {
            std::vector<std::string> reduced = {alleles[0]};
            for (int i : call.called) if (alleles[(size_t)i] != alleles[0]) reduced.push_back(alleles[(size_t)i]);
            if (same_length(reduced)) return false;
            SpanEventInput in2 = in;
            in2.alleles = reduced;
            CallTallies ct2;
            tally_calls_host(in2, opt, ct2);   // the second round's allele list is only known here: a handful of events per run
            alleles = reduced;
            call = genotype_pool(indel_helpers(ct2, alleles.size(), ploidy), ploidy, opt.het_rate);
        }
    }
    const bool hom_ref = call.called.size() == 1 && call.called[0] == 0;
    if (call.undecided || hom_ref || opt.min_quality > call.gq) return false;
    out = IndelRecord();
    out.pos = in.pos;
    out.last = in.pos + (int)in.ref_allele.size() - 1;
    out.alleles = alleles;
    out.str = false;
    out.pool = true;
    out.qual = 0;
    out.gq = call.gq;
    out.depth = call.depth;
    out.called = call.called;
    out.counts = call.counts;
    out.logc = call.logc;
    return true;
}

// SSVPL.genotypeVariantSample's pool branch (SSVPL:366-377) + minQuality => undecided (SSVPL:386-388) for a known SNV: col / ev = the site's
// NGSEP_CALL_POOL column (nullptr: no record reaches the site), alleles = the variant's alleles as base indexes.
SampleIndelCall genotype_known_snv_pool(const ngsep_snv_call* col, const uint8_t* ev, const std::vector<int>& alleles, int ploidy, double het_rate, int min_quality) {
    static const uint8_t none = 0;
    PoolCall call = genotype_pool(snv_helpers(col ? ev : &none, col ? col->sample : 0, col ? col->depth : 0, alleles, ploidy), ploidy, het_rate);
    SampleIndelCall sc;
    if (call.undecided) return sc;                      // SSVPL:432-436: no depth, no report
    sc.called = call.called; sc.gq = call.gq; sc.depth = call.depth; sc.has_report = true; sc.counts = call.counts; sc.logc = call.logc;
    if (min_quality > sc.gq) { sc.called.clear(); sc.gq = 0; }
    return sc;
}

// ... and for a known variant that is not an SNV (SSVPL:376-377): PR.getAlleleCalls(reference allele length) against the variant's alleles
SampleIndelCall genotype_known_indel_pool_at(int pos, const std::vector<std::string>& alleles, std::vector<HostRead*>& pileup, const IndelCallerOptions& opt) {
    SpanEventInput in;
    in.pos = pos; in.ref_allele = alleles[0]; in.alleles = alleles;
    const int span = (int)alleles[0].size();
    for (HostRead* r : pileup) {
        SpanCall c;
        if (!r->span_call(pos, pos + span - 1, c.allele, c.quals)) continue;
        c.allele = upper(c.allele);
        c.neg = r->neg;
        in.calls.push_back(std::move(c));
    }
    CallTallies ct;
    tally_calls_host(in, opt, ct);
    PoolCall call = genotype_pool(indel_helpers(ct, alleles.size(), opt.ploidy), opt.ploidy, opt.het_rate);
    SampleIndelCall sc;
    if (call.undecided) return sc;
    sc.called = call.called; sc.gq = call.gq; sc.depth = call.depth; sc.has_report = true; sc.counts = call.counts; sc.logc = call.logc;
    if (opt.min_quality > sc.gq) { sc.called.clear(); sc.gq = 0; }
    return sc;
}

}  // namespace host
}  // namespace ngsep_b200
