// GpuMultisampleVariantsDetector: drop-in for MultisampleVariantsDetector.run (MSVD = discovery/MultisampleVariantsDetector.java:421-462,
// onPileup :521-537) over in-memory, coordinate-sorted alignment records of all samples.  Same four stages as the
// single-sample detector (detector.cpp); what changes is
//   * the device context carries the sample axis (ngsep_samples_set / ngsep_tile.group / ngsep_tile_fetch_pop): pooled
//     allele discovery + per-sample SNVQ genotypes for every span-1 position happen in k_genotype_columns_multi;
//   * span > 1 events go through the population indel caller (call_population_indel_at, indel.cpp);
//   * records are population VCF records (vcf/VCFRecord.java:277-301: NS, AN, AFS, OH, MAF).
// Other citations: SSVPL = discovery/SingleSampleVariantPileupListener.java, RAFR = alignments/io/ReadAlignmentFileReader.java,
// variants/DiversityStatistics.java:123-218, vcf/VCFFileWriter.java:44-308, variants/CalledSNV.java, variants/CalledGenomicVariantImpl.java.
#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdio>
#include <map>

#include "pipeline.hpp"

namespace ngsep_b200 {
namespace host {

namespace {
// Iteration order of a java.util.HashSet<String> holding `keys` (distinct, given in insertion order): String.hashCode,
// HashMap's hash spreading, table of 16 doubling past a load of 0.75, bins in insertion order.
std::vector<int> java_hashset_order(const std::vector<std::string>& keys) {
    size_t cap = 16;
    while (keys.size() > cap * 3 / 4) cap *= 2;
    std::vector<std::pair<unsigned, int>> tagged;
    for (size_t i = 0; i < keys.size(); i++) {
        unsigned h = 0;
        for (unsigned char c : keys[i]) h = 31u * h + c;
        h ^= (h >> 16);
        tagged.push_back({h & (unsigned)(cap - 1), (int)i});
    }
    std::stable_sort(tagged.begin(), tagged.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
    std::vector<int> out;
    for (auto& t : tagged) out.push_back(t.second);
    return out;
}
}  // namespace

std::vector<SampleDef> samples_from_read_groups(const std::vector<std::string>& rg_ids, const std::vector<std::string>& rg_samples, int ploidy) {
    // MSVD:492-516.  reader.getSampleIdsByReadGroup() is a HashMap (RAFR:70,184-197): read groups reach each sample's
    // HashSet in the map's key order; samples are listed in TreeMap (id) order.
    std::map<std::string, std::vector<int>> by_sample;
    for (int k : java_hashset_order(rg_ids)) {
        const std::string& sm = (k < (int)rg_samples.size() && !rg_samples[k].empty()) ? rg_samples[k] : rg_ids[k];   // RAFR:188
        by_sample[sm].push_back(k);
    }
    std::vector<SampleDef> out;
    for (auto& kv : by_sample) {
        std::vector<std::string> names;
        for (int k : kv.second) names.push_back(rg_ids[k]);
        SampleDef s;
        s.id = kv.first;
        s.ploidy = ploidy;
        for (int o : java_hashset_order(names)) s.rgs.push_back(kv.second[o]);
        out.push_back(std::move(s));
    }
    return out;
}

static DetectorOptions multi_options(DetectorOptions o) {
    o.het_modified = true;   // MSVD has no haploid default for the heterozygosity rate (MSVD:80)
    o.known_strs.clear(); o.known_variants.clear();   // SingleSampleVariantsDetector options
    return o;
}

GpuMultisampleVariantsDetector::GpuMultisampleVariantsDetector(const DetectorOptions& o, std::vector<SampleDef> samples, int n_read_groups)
    : GpuSingleSampleVariantsDetector(multi_options(o)), samples_(std::move(samples)), n_read_groups_(n_read_groups) {}

int GpuMultisampleVariantsDetector::create_ctx(int device, void** out) {
    void* ctx_ = nullptr;   // (local: the members are set by ensure_ctx)
    *out = nullptr;
    if (samples_.size() < 2) { err_ = "MultisampleVariantsDetector on the device needs at least two samples (use SingleSampleVariantsDetector for one)"; return NGSEP_EINVAL; }
    ngsep_params p{};
    p.max_base_qs = opt_.max_base_qs;
    p.min_quality = opt_.min_quality;
    p.het_rate = opt_.het_rate;
    p.ploidy = opt_.ploidy;
    p.flags = opt_.ignore_lower ? NGSEP_F_IGNORE_LOWER_REF : 0u;
    p.n_samples = (int32_t)samples_.size();
    double tables[NGSEP_N_TABLES];
    host_tables(opt_.het_rate, tables);
    int rc = ngsep_ctx_create(device, &p, tables, NGSEP_N_TABLES, &ctx_);
    if (rc != NGSEP_OK) { err_ = ngsep_last_error(nullptr); return rc; }
    std::vector<int32_t> off = {0};
    std::vector<uint16_t> groups;
    for (auto& s : samples_) {
        for (int g : s.rgs) groups.push_back((uint16_t)g);
        off.push_back((int32_t)groups.size());
    }
    if (groups.empty()) groups.push_back(0);
    rc = ngsep_samples_set(ctx_, std::max(1, n_read_groups_), off.data(), groups.data(), opt_.min_allele_depth_freq);
    if (rc != NGSEP_OK) { err_ = ngsep_last_error(ctx_); ngsep_ctx_destroy(ctx_); return rc; }
    *out = ctx_;
    return rc;
}

void GpuMultisampleVariantsDetector::on_span_event(int e, const std::string& ref_upper, const EventOutcome& ev, std::vector<HostRead*>& pileup, SeqPlan& plan,
                                                   int& last_indel_end) {
    IndelCallerOptions io;
    io.max_base_qs = opt_.max_base_qs; io.het_rate = opt_.het_rate; io.min_quality = opt_.min_quality; io.csb = false;
    PopIndelRecord rec;
    const int r = call_population_indel_at(e, ref_upper, ev.str, pileup, samples_, io, rec);
    if (r == 0) return;                 // discoverPopulationSNV decides (MSVD:608-616): the device's result for e stands
    plan.suppressed.insert(e);          // an indel variant was genotyped here: no SNV is tried (MSVD:531-533)
    if (r == 2) {
        last_indel_end = rec.last;      // MSVD:536-538
        if (rec.last > e) plan.embedded.push_back({e + 1, rec.last});
        plan.pop_indels.push_back(std::move(rec));
        plan.stats.indel_calls++;
    }
}

int GpuMultisampleVariantsDetector::fetch_tile(void* ctx, SeqJob& job, ngsep_tile_stats& ts) {
    const size_t n_samples = samples_.size();
    size_t n = 0;
    if (sbuf_.size() < 4096) { sbuf_.resize(4096); cbuf_.resize(4096 * n_samples); }
    int rc = ngsep_tile_fetch_pop(ctx, sbuf_.data(), sbuf_.size(), &n, cbuf_.data(), nullptr, &ts);
    if (rc == NGSEP_EOVERFLOW && n > sbuf_.size()) {
        sbuf_.resize(n + 16); cbuf_.resize((n + 16) * n_samples);
        rc = ngsep_tile_fetch_pop(ctx, sbuf_.data(), sbuf_.size(), &n, cbuf_.data(), nullptr, &ts);
    }
    if (rc != NGSEP_OK) { err_ = ngsep_last_error(ctx); return rc; }
    job.sites.insert(job.sites.end(), sbuf_.begin(), sbuf_.begin() + (ptrdiff_t)n);
    job.scalls.insert(job.scalls.end(), cbuf_.begin(), cbuf_.begin() + (ptrdiff_t)(n * n_samples));
    return NGSEP_OK;
}

void GpuMultisampleVariantsDetector::emit_text(SeqJob& job) {
    const std::string& chrom = genome_->names[job.seq];
    SeqPlan& plan = job.plan;
    const int win_first = plan.win_first;
    const int n_samples = (int)samples_.size();
    std::vector<ngsep_pop_site>& sites = job.sites;
    std::vector<ngsep_sample_call>& scalls = job.scalls;

    // ---- ordered merge + text: MSVD prints each record as its position is reached (MSVD:534-535) ----------------------
    std::vector<size_t> order(sites.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t c) { return sites[a].pos < sites[c].pos; });
    struct Line { int first, last, order; std::string text; };
    std::vector<Line> lines;
    size_t ei = 0;
    for (size_t oi : order) {
        const ngsep_pop_site& c = sites[oi];
        if (plan.suppressed.count(c.pos)) continue;
        while (ei < plan.embedded.size() && plan.embedded[ei].second < c.pos) ei++;
        bool emb = ei < plan.embedded.size() && plan.embedded[ei].first <= c.pos && c.pos <= plan.embedded[ei].second;
        if (emb && !opt_.call_embedded) continue;
        lines.push_back({c.pos, c.pos, (int)lines.size(), pop_snv_vcf_line(c, scalls.data() + oi * (size_t)n_samples, n_samples, chrom, opt_.ploidy, emb)});
        job.snv_records++;
    }
    for (auto& r : plan.pop_indels) {
        bool mine = true;
        if (opt_.n_shards > 1) mine = (((r.pos - win_first) / tile_len()) % opt_.n_shards) == opt_.shard;
        if (mine) lines.push_back({r.pos, r.last, (int)lines.size(), pop_indel_vcf_line(r, chrom, opt_.ploidy)});
    }
    std::stable_sort(lines.begin(), lines.end(), [](const Line& a, const Line& c) { return a.first < c.first; });   // one record per position
    for (auto& l : lines) job.text += l.text;
}

// ---- population VCF text ---------------------------------------------------------------------------------------------
std::string english_fmt(double v) {   // DecimalFormat("##0.0#"), HALF_EVEN on the exact binary value
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.2f", v);
    std::string s = buf;
    if (s.size() >= 2 && s.back() == '0') s.pop_back();
    return s;
}

std::vector<int> impl_copy_numbers(const std::vector<int>& called, const std::vector<int>& counts, bool has_counts, int ploidy, size_t n) {
    std::vector<int> cn(n, 0);
    const int nc = (int)called.size();
    if (nc == 0) return cn;
    if (nc == 1 && called[0] == 0) { cn[0] = ploidy; return cn; }
    if (ploidy <= nc) { for (int k : called) cn[k] = 1; return cn; }
    if (!has_counts) {
        int def = ploidy / nc;
        for (int k : called) cn[k] = def;
        cn[called[0]] += ploidy - def * nc;
        return cn;
    }
    std::vector<int> rc(nc);
    int tot = 0;
    for (int k = 0; k < nc; k++) { rc[k] = counts[called[k]]; if (rc[k] == 0) rc[k] = 1; tot += rc[k]; }
    int sum = 0;
    for (int k = 0; k < nc; k++) { cn[called[k]] = (int)std::max(1L, jround((double)ploidy * rc[k] / tot)); sum += cn[called[k]]; }
    if (sum < ploidy) cn[called[0]] += ploidy - sum;
    else { int ex = sum - ploidy; for (int k = nc - 1; ex > 0 && k >= 0; k--) { int rm = std::min(ex, cn[called[k]] - 1); cn[called[k]] -= rm; ex -= rm; } }
    return cn;
}

namespace {

struct SampleText { std::vector<int> called; std::vector<int> cn; std::string field; };

// INFO of a population record: DiversityStatistics.calculateDiversityStatistics(calls, false) as printed by
// VCFRecord.updateDiversityStatistics (NS, AN, AFS, OH, MAF for biallelic variants), then TYPE (VCFFileWriter.java:47-49).
std::string population_info(const std::vector<SampleText>& st, size_t n_alleles, const char* type_name) {
    std::vector<int> counts(n_alleles, 0);
    double sum = 0;
    int genotyped = 0, n_het = 0;
    for (auto& s : st) {
        if (s.called.empty()) continue;
        genotyped++;
        if (s.called.size() > 1) n_het++;
        for (int j : s.called) if (j >= 0 && (size_t)j < n_alleles) { counts[j] += s.cn[j]; sum += s.cn[j]; }
    }
    int n_called = 0, min_ac = 0;
    for (size_t i = 0; i < n_alleles; i++) if (counts[i] > 0) { n_called++; if (min_ac == 0 || min_ac > counts[i]) min_ac = counts[i]; }
    const double oh = genotyped > 0 ? (double)n_het / genotyped : 0.0;
    const double maf = n_called < 2 ? 0.0 : (double)min_ac / sum;
    std::string s = "NS=" + std::to_string(genotyped) + ";AN=" + std::to_string(n_called) + ";AFS=";
    for (size_t i = 0; i < n_alleles; i++) { if (i) s += ","; s += std::to_string(counts[i]); }
    s += ";OH=" + english_fmt(oh);
    if (n_alleles == 2) s += ";MAF=" + english_fmt(maf);
    if (type_name) { s += ";TYPE="; s += type_name; }
    return s;
}

void append_cn(std::string& f, std::vector<int> cn, bool undecided, int ploidy) {   // VCFFileWriter.java:236-248
    if (ploidy == 0) { f += "."; return; }
    if (undecided) cn[0] = ploidy;
    for (size_t j = 0; j < cn.size(); j++) { if (j) f += ","; f += std::to_string(cn[j]); }
}

}  // namespace

std::string pop_snv_vcf_line(const ngsep_pop_site& site, const ngsep_sample_call* calls, int n_samples, const std::string& chrom, int ploidy, bool embedded) {
    static const char B[] = "ACGT";
    const int nA = site.n_alleles;
    const bool snv_class = nA == 2;   // variant instanceof SNV: CalledSNV genotype calls (VD:33-60), else CalledGenomicVariantImpl (VD:61-95)
    auto Lij = [](const ngsep_sample_call& c, int i, int j) {
        if (i == j) return c.logcond[i];
        const int lo = std::min(i, j), hi = std::max(i, j);
        return c.logcond[4 + ((lo == 0) ? hi - 1 : ((lo == 1) ? hi + 1 : 5))];
    };
    std::vector<SampleText> st((size_t)n_samples);
    for (int s = 0; s < n_samples; s++) {
        const ngsep_sample_call& c = calls[s];
        SampleText& t = st[s];
        if (c.g0 >= 0) { if (c.g0 == c.g1) t.called = {c.g0}; else t.called = {c.g0, c.g1}; }
        std::string& f = t.field;
        append_gt(f, t.called, ploidy);
        f += ":";
        const bool no_data = c.depth == 0;   // VD:23-27: an undecided CalledGenomicVariantImpl with the counts and no report
        if (snv_class && !no_data) {
            const int r = site.allele[0], a = site.allele[1];
            const float hr = (float)Lij(c, r, r), ra = (float)Lij(c, r, a), ar = (float)Lij(c, a, r), ha = (float)Lij(c, a, a);   // CalledSNV.java:259-266
            const bool present = (float)(hr + ra + ar + ha) != 0.0f;                                                              // CalledSNV.java:422
            const double v[3] = {hr, ra, ha};
            for (int k = 0; k < 3; k++) { if (k) f += ","; f += std::to_string(present ? jround(-10 * v[k]) : 0L); }
            // CalledSNV.updateAllelesCopyNumberFromCounts at ploidy <= 2 (CalledSNV.java:134-147)
            if (t.called.empty()) t.cn = {0, 0};
            else if (t.called.size() == 1) t.cn = (t.called[0] == 0) ? std::vector<int>{ploidy, 0} : std::vector<int>{0, ploidy};
            else t.cn = {1, 1};
        } else {
            bool firstv = true;
            for (int j = 0; j < nA; j++) for (int i = 0; i <= j; i++) {
                if (!firstv) f += ",";
                firstv = false;
                f += std::to_string(no_data ? 0L : jround(-10 * Lij(c, site.allele[i], site.allele[j])));
            }
            std::vector<int> rep((size_t)nA);
            for (int i = 0; i < nA; i++) rep[i] = c.count[site.allele[i]];
            t.cn = impl_copy_numbers(t.called, rep, !no_data, ploidy, (size_t)nA);
        }
        f += ":"; f += std::to_string((int)c.gq);
        f += ":"; f += std::to_string(c.depth);
        f += ":";
        for (int i = 0; i < 4; i++) { if (i) f += ","; f += std::to_string(c.count[i]); }
        f += ":";
        append_cn(f, t.cn, t.called.empty(), ploidy);
    }
    // a variant that was reduced to three alleles is a plain GenomicVariantImpl: its type is undetermined (SSVPL:346-359)
    const char* type_name = nullptr;
    if (embedded) type_name = "EMBEDDED";                                    // MSVD:580-582
    else if (nA > 2 && site.n_alleles0 == nA) type_name = "MULTISNV";        // SSVPL:326-328
    std::string s = chrom;
    s += "\t"; s += std::to_string(site.pos); s += "\t.\t"; s += B[site.allele[0]]; s += "\t";
    for (int i = 1; i < nA; i++) { if (i > 1) s += ","; s += B[site.allele[i]]; }
    s += "\t"; s += std::to_string((int)site.qs); s += "\t.\t";
    s += population_info(st, (size_t)nA, type_name);
    s += "\tGT:PL:GQ:DP:BSDP:ACN";
    for (auto& t : st) { s += "\t"; s += t.field; }
    s += "\n";
    return s;
}

std::string pop_indel_vcf_line(const PopIndelRecord& r, const std::string& chrom, int ploidy) {
    const size_t n = r.alleles.size();
    // VariantCallReport resolves alleles through a TreeMap name->index: with duplicate names the last index wins
    auto index_of = [&](const std::string& a) { int k = -1; for (size_t i = 0; i < n; i++) if (r.alleles[i] == a) k = (int)i; return k; };
    std::vector<SampleText> st(r.calls.size());
    for (size_t s = 0; s < r.calls.size(); s++) {
        const SampleIndelCall& c = r.calls[s];
        SampleText& t = st[s];
        t.called = c.called;
        std::string& f = t.field;
        append_gt(f, t.called, ploidy);
        f += ":";
        bool firstv = true;
        for (size_t j = 0; j < n; j++) for (size_t i = 0; i <= j; i++) {
            if (!firstv) f += ",";
            firstv = false;
            f += std::to_string(c.has_report ? jround(-10 * c.logc[index_of(r.alleles[i])][index_of(r.alleles[j])]) : 0L);
        }
        f += ":"; f += std::to_string(c.gq);
        f += ":"; f += std::to_string(c.depth);
        f += ":";
        std::vector<int> rep(n, 0);
        for (size_t i = 0; i < n; i++) { rep[i] = c.has_report ? c.counts[index_of(r.alleles[i])] : 0; if (i) f += ","; f += std::to_string(rep[i]); }
        t.cn = impl_copy_numbers(t.called, rep, c.has_report, ploidy, n);
        f += ":";
        append_cn(f, t.cn, t.called.empty(), ploidy);
    }
    std::string s = chrom;
    s += "\t"; s += std::to_string(r.pos); s += "\t.\t"; s += r.alleles[0]; s += "\t";
    for (size_t i = 1; i < n; i++) { if (i > 1) s += ","; s += r.alleles[i]; }
    s += "\t"; s += std::to_string(r.qs); s += "\t.\t";
    s += population_info(st, n, r.str ? "STR" : "INDEL");
    s += "\tGT:PL:GQ:DP:ADP:ACN";
    for (auto& t : st) { s += "\t"; s += t.field; }
    s += "\n";
    return s;
}

}  // namespace host
}  // namespace ngsep_b200
