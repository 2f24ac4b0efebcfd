// RelativeAlleleCounts: drop-in for RelativeAlleleCountsCalculator.runProcess + printResults
// (discovery/RelativeAlleleCountsCalculator.java:209-243) over in-memory, coordinate-sorted alignment records.  The generator it
// configures admits maxRD alignments per start position (and secondary alignments on request); the listener (:246-290) looks at
// PileupRecord.getAlleleCalls(1) of EVERY position: the device counts the alleles of every position of a region tile
// (ngsep_tile_allele_counts: largest count, second largest, number of alleles), the host feeds them — in position order, as the
// listener would see them — to the three distributions of the report (math/Distribution.java).  Regions to filter / select (-frs,
// -srs; :131-161, 247-254): the positions inside a region to filter, and with regions to select the positions outside all of them,
// are left out before anything is counted — per sequence the union of the regions as sorted disjoint intervals, walked along with
// the positions.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <map>

#include "pipeline.hpp"

namespace ngsep_b200 {
namespace host {

namespace {

// A region file as SimpleGenomicRegionFileHandler.loadRegions reads it (genome/io/SimpleGenomicRegionFileHandler.java:57-80): fields
// separated by ONE blank or tab, sequence / first / last in the first three; a line whose numbers Integer.parseInt would refuse, or
// with fewer than three fields, is skipped.  Kept per sequence name as the union of its regions (findSpanningRegions only asks
// whether some region spans the position).
struct RegionUnion {
    bool given = false;
    std::map<std::string, std::vector<std::pair<int, int>>> spans;   // sorted, disjoint
    static bool java_int(const char* s, const char* e, long& v) {
        if (s == e) return false;
        const bool neg = *s == '-';
        if (*s == '-' || *s == '+') s++;
        if (s == e) return false;
        long x = 0;
        for (; s < e; s++) { if (*s < '0' || *s > '9') return false; x = x * 10 + (*s - '0'); if (x > 2147483648L) return false; }
        v = neg ? -x : x;
        return v >= -2147483648L && v <= 2147483647L;
    }
    bool load(const std::string& path, std::string& err) {
        if (path.empty()) return true;
        std::ifstream in(path);
        if (!in) { err = "cannot open " + path; return false; }
        given = true;
        std::string line;
        while (std::getline(in, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            size_t b[3], e[3];
            size_t at = 0;
            int k = 0;
            for (; k < 3; k++) {
                b[k] = at;
                const size_t sep = line.find_first_of(" \t", at);
                e[k] = sep == std::string::npos ? line.size() : sep;
                if (sep == std::string::npos) { k++; break; }
                at = sep + 1;
            }
            if (k < 3) continue;
            long f, l;
            if (!java_int(line.data() + b[1], line.data() + e[1], f) || !java_int(line.data() + b[2], line.data() + e[2], l)) continue;
            if (f <= l) spans[line.substr(b[0], e[0] - b[0])].push_back({(int)f, (int)l});   // (a region with first > last spans nothing)
        }
        for (auto& kv : spans) {
            auto& v = kv.second;
            std::sort(v.begin(), v.end());
            size_t w = 0;
            for (size_t i = 0; i < v.size(); i++) {
                if (w > 0 && v[i].first <= v[w - 1].second) v[w - 1].second = std::max(v[w - 1].second, v[i].second);
                else v[w++] = v[i];
            }
            v.resize(w);
        }
        return true;
    }
    const std::vector<std::pair<int, int>>* of(const std::string& seq) const { auto it = spans.find(seq); return it == spans.end() ? nullptr : &it->second; }
};
// positions come in ascending order: the cursor only moves forward
struct RegionCursor {
    const std::vector<std::pair<int, int>>* v = nullptr;
    size_t k = 0;
    bool spans(int pos) { if (!v) return false; while (k < v->size() && (*v)[k].second < pos) k++; return k < v->size() && (*v)[k].first <= pos; }
};

struct Distribution {   // math/Distribution.java:46-78, 302-345
    double sum = 0, sum_sq = 0, count = 0, min_v, max_v, bin;
    std::vector<double> dist;
    long less = 0, more = 0;
    Distribution(double mn, double mx, double b) : min_v(mn), max_v(mx), bin(b), dist((size_t)((int)((mx - mn) / b) + 1), 0.0) {}
    void add(double v) {
        sum += v; sum_sq += v * v; count += 1;
        if (v >= min_v && v <= max_v) dist[(size_t)(int)((v - min_v) / bin)] += 1;
        else if (v < min_v) less++;
        else more++;
    }
    void print(std::string& out, bool integers) const {
        auto num = [&](double v) { return integers ? std::to_string((long)jround(v)) : english_fmt(v); };
        if (less > 0) out += "Less\t" + num((double)less) + "\n";
        const int max_idx = (int)((max_v - min_v) / bin);
        for (int i = 0; i < (int)dist.size() && i <= max_idx; i++) {
            if (integers) out += std::to_string((int)(min_v + i * bin)) + "\t" + num(dist[(size_t)i]) + "\n";
            else out += english_fmt(min_v + i * bin) + "\t" + num(dist[(size_t)i]) + "\n";
        }
        double more_count = (double)more;
        for (int i = max_idx + 1; i < (int)dist.size(); i++) more_count += dist[(size_t)i];
        if (more_count > 0) out += "More\t" + num(more_count) + "\n";
        out += "Count\t" + std::to_string((long)jround(count)) + "\n";
        out += "Sum\t" + num(sum) + "\n";
        if (count > 0) out += "Average\t" + english_fmt(sum / count) + "\n";
        if (count > 1) {
            const double variance = (sum_sq - sum * sum / count) / (count - 1);
            out += "Variance\t" + english_fmt(variance) + "\n";
            out += "STDev\t" + english_fmt(std::sqrt(variance)) + "\n";
        }
    }
};

}  // namespace

int run_relative_allele_counts(const Genome& genome, const BatchView& b, int min_rd, int max_rd, int min_base_qs, bool secondary, int device, std::string& out,
                               std::string& err, const std::string& regions_to_filter, const std::string& regions_to_select) {
    RegionUnion filter, select;
    if (!filter.load(regions_to_filter, err) || !select.load(regions_to_select, err)) return NGSEP_EINVAL;
    ngsep_params p{};
    p.max_base_qs = 30; p.min_quality = 40; p.het_rate = 0.001; p.ploidy = 2; p.n_samples = 1;
    double tables[NGSEP_N_TABLES];
    host_tables(p.het_rate, tables);
    void* ctx = nullptr;
    int rc = ngsep_ctx_create(device, &p, tables, NGSEP_N_TABLES, &ctx);
    if (rc != NGSEP_OK) { err = ngsep_last_error(nullptr); return rc; }
    Distribution dist_prop(0, 0.5, 0.01), dist_num(1, 10, 1);
    std::vector<std::string> seq_names;
    std::vector<Distribution> per_seq;
    AdmissionOptions ao;
    ao.min_mq = 20; ao.max_alns_per_start = max_rd; ao.secondary = secondary;   // :213-214; the generator's own minMQ
    const int min_qual_bits = (min_base_qs >= 4 && min_base_qs <= 30) ? 0 : 8;   // coded qualities keep min(30, q), every q <= 3 alike
    PackedTile tile;
    std::vector<uint32_t> counts;
    int64_t i = 0;
    while (i < b.n && rc == NGSEP_OK) {
        const int seq = b.seq_idx[i];
        int64_t j = i;
        while (j < b.n && b.seq_idx[j] == seq) j++;
        if (seq >= 0 && seq < (int)genome.seqs.size()) {
            std::vector<HostRead> reads = admit_sequence(b, i, j, ao);
            const int seq_len = (int)genome.seqs[(size_t)seq].size();
            if (!reads.empty()) {
                Distribution* cur = nullptr;   // onSequenceStart, :309-320
                if (seq_len > 100000) { seq_names.push_back(genome.names[(size_t)seq]); per_seq.emplace_back(0, 0.5, 0.01); cur = &per_seq.back(); }
                rc = ngsep_ref_upload(ctx, seq, genome.seqs[(size_t)seq].data(), (int64_t)seq_len);
                if (rc != NGSEP_OK) { err = ngsep_last_error(ctx); break; }
                RegionCursor in_filter, in_select;
                in_filter.v = filter.of(genome.names[(size_t)seq]); in_select.v = select.of(genome.names[(size_t)seq]);
                int max_span = 1;
                for (auto& r : reads) { if (r.L > 0xffff || r.st.cigar.size() > 0xff) { err = "read longer than 65535 bases or with more than 255 CIGAR operations"; rc = NGSEP_EINVAL; break; } max_span = std::max(max_span, r.st.last - r.admit_first + 1); }
                for (int64_t tf = 1; tf <= seq_len && rc == NGSEP_OK; tf += kAutoTileLen) {
                    const int tl = (int)std::min<int64_t>(seq_len, tf + kAutoTileLen - 1);
                    const size_t lo = std::lower_bound(reads.begin(), reads.end(), (int)tf - max_span + 1, [](const HostRead& r, int v) { return r.admit_first < v; }) - reads.begin();
                    const size_t hi = std::upper_bound(reads.begin(), reads.end(), tl, [](int v, const HostRead& r) { return v < r.admit_first; }) - reads.begin();
                    if (lo >= hi) continue;
                    if (!pack_reads(reads, lo, hi, tile, false, &err, min_qual_bits)) { rc = NGSEP_EINVAL; break; }
                    if (tile.hdr.empty()) continue;
                    ngsep_tile tv = tile.view(seq, (int)tf, tl);
                    counts.resize(2 * (size_t)(tl - tf + 1));
                    rc = ngsep_tile_allele_counts(ctx, &tv, min_rd, min_base_qs, counts.data());
                    if (rc != NGSEP_OK) { err = ngsep_last_error(ctx); break; }
                    for (size_t k = 0; k < counts.size(); k += 2) {   // onPileup, :269-284, position by position
                        const uint32_t n_alleles = counts[k] >> 28;
                        if (n_alleles == 0) continue;
                        const int pos = (int)tf + (int)(k / 2);
                        if (filter.given && in_filter.spans(pos)) continue;     // :247-250
                        if (select.given && !in_select.spans(pos)) continue;    // :251-254
                        const uint32_t c1 = counts[k] & 0x0fffffffu, c2 = counts[k + 1];
                        dist_num.add((double)n_alleles);
                        const double prop = (double)c2 / (double)(c1 + c2);
                        dist_prop.add(prop);
                        if (cur) cur->add(prop);
                    }
                }
            }
        }
        i = j;
    }
    ngsep_ctx_destroy(ctx);
    if (rc != NGSEP_OK) return rc;
    out += "Distribution of allele proportions\n";      // printResults, :221-243
    dist_prop.print(out, false);
    out += "Distribution of number of alleles\n";
    dist_num.print(out, true);
    if (seq_names.empty()) return NGSEP_OK;
    out += "Distribution of allele proportions per sequence\nProportion";
    for (auto& n : seq_names) out += "\t" + n;
    out += "\n";
    double mn = 0;
    for (size_t k = 0; k < dist_prop.dist.size(); k++) {
        out += english_fmt(mn);
        for (auto& d : per_seq) out += "\t" + english_fmt(d.dist[k]);
        out += "\n";
        mn += 0.01;
    }
    return NGSEP_OK;
}

}  // namespace host
}  // namespace ngsep_b200
