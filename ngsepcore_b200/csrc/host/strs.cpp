// -knownSTRs (SingleSampleVariantsDetector.java:906-913): the tandem-repeat loci the indel realigner takes as input variants.
// The first three columns of the file are sequence, first, last (genome/io/SimpleGenomicRegionFileHandler.java:57-80; the yeast
// tutorial's training/Saccharomyces_cerevisiae_STRs.txt is a TRF table read this way); neighbouring loci are merged and every
// locus is padded by one base on either side (makeNonRedundantSTRs / mergeSTRs / makeSTRVariant, SSVD:843-894).
#include <algorithm>
#include <cstdlib>
#include <fstream>

#include "pipeline.hpp"

namespace ngsep_b200 {
namespace host {

namespace {
struct Row { int seq, first, last; };

// AbstractLimitedSequence.getOverlapLength (sequences/AbstractLimitedSequence.java:376-390): the longest suffix of a that is a
// prefix of b (or that b, being shorter, fits into)
int suffix_prefix_overlap(const std::string& a, const std::string& b) {
    for (size_t i = 0; i < a.size(); i++) {
        size_t x = i, y = 0;
        while (x < a.size() && y < b.size() && a[x] == b[y]) { x++; y++; }
        if (x == a.size()) return (int)(a.size() - i);
    }
    return 0;
}
}  // namespace

bool load_known_strs(const std::string& path, const Genome& genome, std::vector<std::vector<StrLocus>>& out, std::string& err) {
    std::ifstream in(path);
    if (!in) { err = "cannot open " + path; return false; }
    std::vector<Row> rows;
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::string item[3];
        size_t s = 0;
        int k = 0;
        for (; k < 3; k++) {
            size_t e = line.find_first_of(" \t", s);
            item[k] = line.substr(s, e == std::string::npos ? std::string::npos : e - s);
            if (e == std::string::npos) { k++; break; }
            s = e + 1;
        }
        if (k < 3) continue;
        char* end = nullptr;
        const long f = std::strtol(item[1].c_str(), &end, 10);
        if (item[1].empty() || *end) continue;       // not a number: the line is skipped with a warning in the reference
        const long l = std::strtol(item[2].c_str(), &end, 10);
        if (item[2].empty() || *end) continue;
        int seq = -1;
        for (size_t i = 0; i < genome.names.size(); i++) if (genome.names[i] == item[0]) { seq = (int)i; break; }
        if (seq < 0) continue;                       // a sequence the genome does not have yields no variant
        rows.push_back({seq, (int)f, (int)l});
    }
    std::stable_sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) { return a.seq != b.seq ? a.seq < b.seq : (a.first != b.first ? a.first < b.first : a.last < b.last); });
    out.assign(genome.seqs.size(), {});
    size_t i = 0;
    while (i < rows.size()) {
        const int seq = rows[i].seq;
        const int seq_len = (int)genome.seqs[(size_t)seq].size();
        auto emit = [&](int first, int last) {
            std::string ref;
            if (!genome.get(seq, first, last, ref)) return;          // outside the sequence: no variant (SSVD:885-888)
            out[(size_t)seq].push_back({first, last});
        };
        int first = 0, last = 0;
        for (; i < rows.size() && rows[i].seq == seq; i++) {
            const Row& r = rows[i];
            bool merge = false;
            if (last != 0) {
                if (r.first - last > 5) merge = false;
                else if (r.first - last <= 2) merge = true;
                else {
                    std::string a, b;
                    merge = genome.get(seq, std::max(first, last - 10), last, a) && genome.get(seq, r.first, r.last, b) && suffix_prefix_overlap(upper(a), upper(b)) > 5;
                }
            }
            if (last == 0 || !merge) {
                if (last > 0) emit(std::max(1, first - 1), std::min(last + 1, seq_len));
                first = r.first;
            }
            last = r.last;
        }
        if (last > 0) emit(std::max(1, first - 1), std::min(last + 1, seq_len));
    }
    return true;
}

}  // namespace host
}  // namespace ngsep_b200
