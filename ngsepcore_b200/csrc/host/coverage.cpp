// CoverageStats: drop-in for CoverageStatisticsCalculator.processFile (discovery/CoverageStatisticsCalculator.java:119-143) over
// in-memory, coordinate-sorted alignment records.  The generator it configures processes secondary alignments (only unmapped
// records are filtered, APG:367-371) and admits up to 100 alignments per start position and read group; the listener itself only
// needs each position's number of (unique) alignments, which the device computes per region tile (ngsep_coverage_histogram).
#include <algorithm>
#include <cstdio>

#include "pipeline.hpp"

namespace ngsep_b200 {
namespace host {

int run_coverage_stats(const Genome& genome, const BatchView& b, int min_mq, int device, std::string& out, std::string& err) {
    ngsep_params p{};
    p.max_base_qs = 30; p.min_quality = 40; p.het_rate = 0.001; p.ploidy = 2; p.n_samples = 1;
    double tables[NGSEP_N_TABLES];
    host_tables(p.het_rate, tables);
    void* ctx = nullptr;
    int rc = ngsep_ctx_create(device, &p, tables, NGSEP_N_TABLES, &ctx);
    if (rc != NGSEP_OK) { err = ngsep_last_error(nullptr); return rc; }
    const int max_cov = 300;   // CoverageStatisticsCalculator.maxCoverage (:57): depths 1..299 are printed, anything above goes to "More"
    std::vector<uint64_t> all((size_t)max_cov + 1, 0), uniq((size_t)max_cov + 1, 0), ha((size_t)max_cov + 1), hu((size_t)max_cov + 1);
    AdmissionOptions ao;
    ao.min_mq = min_mq; ao.max_alns_per_start = 100; ao.secondary = true;
    int64_t i = 0;
    while (i < b.n && rc == NGSEP_OK) {
        const int seq = b.seq_idx[i];
        int64_t j = i;
        while (j < b.n && b.seq_idx[j] == seq) j++;
        if (seq >= 0 && seq < (int)genome.seqs.size()) {
            std::vector<HostRead> reads = admit_sequence(b, i, j, ao);
            const int seq_len = (int)genome.seqs[seq].size();
            std::vector<int32_t> first(reads.size()), last(reads.size());
            std::vector<uint8_t> unique(reads.size());
            int max_span = 1;
            for (size_t k = 0; k < reads.size(); k++) {
                first[k] = reads[k].st.first; last[k] = reads[k].st.last; unique[k] = reads[k].unique ? 1 : 0;
                max_span = std::max(max_span, last[k] - first[k] + 1);
            }
            // pileups exist from the first read's start to the last covered position of the sequence (APG:453-498)
            for (int64_t tf = 1; tf <= seq_len && rc == NGSEP_OK; tf += kAutoTileLen) {
                const int tl = (int)std::min<int64_t>(seq_len, tf + kAutoTileLen - 1);
                const size_t lo = std::lower_bound(first.begin(), first.end(), (int)tf - max_span + 1) - first.begin();
                const size_t hi = std::upper_bound(first.begin(), first.end(), tl) - first.begin();
                if (lo >= hi) continue;
                rc = ngsep_coverage_histogram(ctx, (int)tf, tl, (int64_t)(hi - lo), first.data() + lo, last.data() + lo, unique.data() + lo, max_cov, ha.data(), hu.data());
                if (rc != NGSEP_OK) { err = ngsep_last_error(ctx); break; }
                for (int d = 0; d <= max_cov; d++) { all[(size_t)d] += ha[(size_t)d]; uniq[(size_t)d] += hu[(size_t)d]; }
            }
        }
        i = j;
    }
    ngsep_ctx_destroy(ctx);
    if (rc != NGSEP_OK) return rc;
    char line[96];
    for (int d = 1; d < max_cov; d++) {   // printCoverageStats (:250-256)
        std::snprintf(line, sizeof line, "%d\t%llu\t%llu\n", d, (unsigned long long)all[(size_t)d], (unsigned long long)uniq[(size_t)d]);
        out += line;
    }
    std::snprintf(line, sizeof line, "More\t%llu\t%llu\n", (unsigned long long)all[(size_t)max_cov], (unsigned long long)uniq[(size_t)max_cov]);
    out += line;
    return NGSEP_OK;
}

}  // namespace host
}  // namespace ngsep_b200
