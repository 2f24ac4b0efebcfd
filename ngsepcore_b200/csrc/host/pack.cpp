// Tile packer: host reads (with their realigner versions) -> the header / payload layout of include/ngsep_pileup.h.
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "pipeline.hpp"

namespace ngsep_b200 {
namespace host {

namespace {

struct CodeLut {
    uint8_t v[256];
    CodeLut() {
        std::memset(v, 4, sizeof(v));
        v['A'] = v['a'] = 0; v['C'] = v['c'] = 1; v['G'] = v['g'] = 2; v['T'] = v['t'] = 3;
    }
};
const CodeLut& lut() { static const CodeLut l; return l; }

inline int ref_last_of(const ReadState& st) {
    int ref = st.first;
    for (uint32_t c : st.cigar) if (cig_ref(c)) ref += cig_len(c);
    return ref - 1;
}

void emit_record(const HostRead& r, const ReadState& st, bool windowed, int from, int to, uint32_t qoff, PackedTile& out) {
    const CodeLut& L = lut();
    const int len = r.L;
    const int nc = (int)st.cigar.size();
    if (len > 0xffff || nc > 0xff) return;  // rejected earlier by the detector; kept as a guard
    // one pass over the read: 2-bit bases, "not ACGT" bits and "quality <= 3" bits (never counted, CH:214)
    const int nb = (len + 15) / 16, nm_words = (len + 31) / 32;
    uint32_t small[3 * 64];
    std::vector<uint32_t> big;
    uint32_t* bw = small;
    if (nb + 2 * nm_words > (int)(sizeof(small) / sizeof(small[0]))) { big.assign((size_t)nb + 2 * (size_t)nm_words, 0u); bw = big.data(); }
    else std::memset(small, 0, sizeof(uint32_t) * (size_t)(nb + 2 * nm_words));
    uint32_t* nmk = bw + nb;
    uint32_t* lq = nmk + nm_words;
    uint32_t any_n = 0u;
    const uint8_t* bases = r.bases;
    const uint8_t* quals = r.quals;
    for (int i = 0; i < len; i++) {
        const uint32_t c = L.v[bases[i]];
        bw[i >> 4] |= (c & 3u) << ((i & 15) * 2);
        const uint32_t isn = c >> 2;                  // 1 for anything outside ACGT
        nmk[i >> 5] |= isn << (i & 31);
        any_n |= isn;
    }
    if (quals) for (int i = 0; i < len; i++) lq[i >> 5] |= (uint32_t)((int)quals[i] - 33 <= 3) << (i & 31);
    const bool has_n = any_n != 0u;
    ngsep_read_hdr h;
    h.first = r.admit_first;
    h.meta = (uint32_t)len | ((uint32_t)nc << 16) | (r.neg ? NGSEP_M_NEG : 0u) | (has_n ? NGSEP_M_NMASK : 0u) | (windowed ? NGSEP_M_WINDOW : 0u);
    if (nc == 1 && !has_n && !windowed && st.cigar[0] == (((uint32_t)len << 3) | 3u)) h.meta |= NGSEP_M_SIMPLE;
    h.ign_start = (int16_t)st.ign_s;
    h.ign_end = (int16_t)st.ign_e;
    h.off = (uint32_t)out.payload.size();
    out.hdr.push_back(h);
    auto& p = out.payload;
    if (windowed) { p.push_back((uint32_t)from); p.push_back((uint32_t)to); p.push_back((uint32_t)st.first); }
    p.insert(p.end(), st.cigar.begin(), st.cigar.end());
    p.insert(p.end(), bw, bw + nb);
    if (has_n) p.insert(p.end(), nmk, nmk + nm_words);
    p.insert(p.end(), lq, lq + nm_words);
    p.push_back(qoff);
    out.max_span = std::max(out.max_span, ref_last_of(st) - r.admit_first + 1);
    out.max_back = std::max(out.max_back, r.admit_first - st.first);
    out.aligned_bases += (uint64_t)len;
}

}  // namespace

static void pack_range(const std::vector<HostRead>& reads, size_t lo, size_t hi, PackedTile& out) {
    out.hdr.clear();
    out.payload.clear();
    out.qual.clear();
    out.max_span = 1;
    out.max_back = 0;
    out.aligned_bases = 0;
    out.hdr.reserve(hi - lo);
    {   // one reservation per array instead of geometric regrowth
        size_t bases = 0;
        for (size_t i = lo; i < hi; i++) bases += (size_t)reads[i].L;
        out.payload.reserve((hi - lo) * 8 + bases / 16 + bases / 16 + 64);
        out.qual.reserve(bases + 4 * (hi - lo) + 64);
    }
    for (size_t i = lo; i < hi; i++) {
        const HostRead& r = reads[i];
        // the read's qualities go once into the quality array (every version of the read points at them)
        const uint32_t qoff = (uint32_t)(out.qual.size() / 4);
        if (r.quals) out.qual.insert(out.qual.end(), r.quals, r.quals + r.L);
        else out.qual.insert(out.qual.end(), (size_t)r.L, (uint8_t)'+');   // absent QUAL reads as '+' (RA:1025)
        out.qual.resize((out.qual.size() + 3) & ~(size_t)3, (uint8_t)'!');
        if (r.closed.empty() && r.st.first == r.admit_first) { emit_record(r, r.st, false, 0, 0, qoff, out); continue; }
        for (const Version& v : r.closed) emit_record(r, v.st, true, v.from, v.to, qoff, out);
        emit_record(r, r.st, true, r.cur_from, INT_MAX, qoff, out);
    }
}

static int host_threads() {
    if (const char* e = std::getenv("NGSEP_HOST_THREADS")) { int v = std::atoi(e); if (v >= 1 && v <= 256) return v; }
    unsigned hw = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(8u, hw ? hw : 1u));
}

void pack_reads(const std::vector<HostRead>& reads, size_t lo, size_t hi, PackedTile& out, bool with_groups) {
    out.group.clear();
    const size_t n = hi > lo ? hi - lo : 0;
    const int nt = (int)std::min<size_t>((size_t)host_threads(), n / 20000 + 1);   // small tiles are not worth a thread each
    if (nt <= 1) pack_range(reads, lo, hi, out);
    else {
        // contiguous ranges of reads are packed independently, then concatenated: payload offsets move by the words in front
        // of the range, quality offsets (the last payload word of every record) by the quality words in front of it
        std::vector<PackedTile> part((size_t)nt);
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) {
            const size_t a = lo + n * (size_t)t / (size_t)nt, b = lo + n * (size_t)(t + 1) / (size_t)nt;
            th.emplace_back([&reads, &part, a, b, t]() { pack_range(reads, a, b, part[(size_t)t]); });
        }
        for (auto& x : th) x.join();
        size_t nh = 0, np = 0, nq = 0;
        for (auto& p : part) { nh += p.hdr.size(); np += p.payload.size(); nq += p.qual.size(); }
        out.hdr.resize(nh); out.payload.resize(np); out.qual.resize(nq);
        out.max_span = 1; out.max_back = 0; out.aligned_bases = 0;
        std::vector<size_t> bh((size_t)nt), bp((size_t)nt), bq((size_t)nt);
        nh = np = nq = 0;
        for (int t = 0; t < nt; t++) { bh[t] = nh; bp[t] = np; bq[t] = nq; nh += part[t].hdr.size(); np += part[t].payload.size(); nq += part[t].qual.size(); }
        th.clear();
        for (int t = 0; t < nt; t++) {
            th.emplace_back([&out, &part, &bh, &bp, &bq, t]() {
                PackedTile& p = part[(size_t)t];
                const uint32_t dp = (uint32_t)bp[t], dq = (uint32_t)(bq[t] / 4);
                for (size_t i = 0; i < p.hdr.size(); i++) {
                    const size_t end = (i + 1 < p.hdr.size()) ? p.hdr[i + 1].off : p.payload.size();
                    p.payload[end - 1] += dq;          // qoff
                    p.hdr[i].off += dp;
                }
                std::copy(p.hdr.begin(), p.hdr.end(), out.hdr.begin() + (ptrdiff_t)bh[t]);
                std::copy(p.payload.begin(), p.payload.end(), out.payload.begin() + (ptrdiff_t)bp[t]);
                std::copy(p.qual.begin(), p.qual.end(), out.qual.begin() + (ptrdiff_t)bq[t]);
            });
        }
        for (auto& x : th) x.join();
        for (auto& p : part) { out.max_span = std::max(out.max_span, p.max_span); out.max_back = std::max(out.max_back, p.max_back); out.aligned_bases += p.aligned_bases; }
    }
    if (with_groups) {   // one id per record (every version of a read carries the read's group)
        out.group.reserve(out.hdr.size());
        size_t k = 0;
        for (size_t i = lo; i < hi; i++) {
            const HostRead& r = reads[i];
            const size_t n_rec = (r.closed.empty() && r.st.first == r.admit_first) ? 1 : r.closed.size() + 1;
            const uint16_t g = (r.rg >= 0 && r.rg < 0xffff) ? (uint16_t)r.rg : (uint16_t)0xffff;
            for (size_t v = 0; v < n_rec && k < out.hdr.size(); v++, k++) out.group.push_back(g);
        }
        out.group.resize(out.hdr.size(), (uint16_t)0xffff);
    }
}

ngsep_tile PackedTile::view(int seq_idx, int first, int last) const {
    ngsep_tile t;
    t.seq_idx = seq_idx; t.tile_first = first; t.tile_last = last;
    t.n_reads = (int)hdr.size();
    t.max_span = max_span; t.max_back = max_back;
    t.n_words = payload.size();
    t.hdr = hdr.data();
    t.payload = payload.data();
    t.group = group.size() == hdr.size() && !group.empty() ? group.data() : nullptr;
    t.qual = qual.data();
    t.n_qual = qual.size();
    return t;
}

}  // namespace host
}  // namespace ngsep_b200
