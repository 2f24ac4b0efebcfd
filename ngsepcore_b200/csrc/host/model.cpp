// Host read model: block-table queries, realigner mutations and admission.  See model.hpp.
#include "model.hpp"

#include <algorithm>
#include <climits>
#include <cstring>
#include <atomic>
#include <queue>
#include <thread>
#include <unordered_map>

namespace ngsep_b200 {
namespace host {

bool Genome::get(int seq, int first, int last, std::string& out) const {  // genome/ReferenceGenome.java:238-243
    if (seq < 0 || seq >= (int)seqs.size()) return false;
    const std::string& s = seqs[seq];
    if (first < 1 || last > (int)s.size() || last < first - 1) return false;
    out.assign(s, first - 1, last - first + 1);
    for (auto& c : out) if (!std::strchr("AaCcNngGtT", c) || c == 0) c = 'N';  // sequences/DNAMaskedSequence.java:33,57-59
    return true;
}

namespace {

// Walks the blocks of a read state.  f(block_index, ref0, read0, len, op) -> true to stop.
template <typename F>
inline void walk(const ReadState& st, F&& f) {
    int ref = st.first, rd = 0;
    for (size_t i = 0; i < st.cigar.size(); i++) {
        uint32_t v = st.cigar[i];
        int len = cig_len(v);
        if (f((int)i, ref, rd, len, cig_op(v))) return;
        if (cig_ref(v)) ref += len;
        if (cig_read(v)) rd += len;
    }
}

// Skip decision of RA:782-795 for base j of a match block.  anchor_ok: out, whether the anchor (j == len-1 before an
// indel) survives.  Returns true when the base is skipped.
inline bool skipped(const HostRead& r, int i, int rd, int j, int len) {
    const auto& cg = r.st.cigar;
    bool prev_indel = i > 0 && op_is_indel(cig_op(cg[i - 1]));
    bool next_indel = false;
    int next_cons = 0;
    if (i + 1 < (int)cg.size()) {
        next_indel = op_is_indel(cig_op(cg[i + 1]));
        next_cons = cig_read(cg[i + 1]) ? cig_len(cg[i + 1]) : 0;
    }
    int rp = rd + j;
    if (rp < r.st.ign_s) return true;
    if (r.L - rp <= r.st.ign_e) return true;
    if (prev_indel && j < 2) return true;
    if (next_indel && j < len - 1 && j >= len - 2) return true;
    if (next_indel && j == len - 1) {
        if (rp < 2 || r.L - rp - next_cons < 2) return true;
        if (r.L - (rp + next_cons + 1) < r.st.ign_e) return true;
    }
    return false;
}

}  // namespace

void HostRead::derive() {
    anchors.clear();
    const auto& cg = st.cigar;
    walk(st, [&](int i, int ref, int rd, int len, int op) {
        (void)op;
        if (!cig_both(cg[i]) || len == 0) return false;
        if (i + 1 >= (int)cg.size()) return false;
        int nop = cig_op(cg[i + 1]);
        if (!op_is_indel(nop)) return false;
        int j = len - 1;
        if (rd + j >= L) return false;
        if (skipped(*this, i, rd, j, len)) return false;
        int nlen = cig_len(cg[i + 1]);
        Anchor a;
        a.pos = ref + j;
        a.ins = nop == kOpI;
        a.len = nlen;
        a.last = a.pos + 1 + (a.ins ? 0 : nlen);
        a.read_off = rd + j;
        // TreeMap.put: a later anchor at the same reference position replaces the earlier one
        for (auto& e : anchors) if (e.pos == a.pos) { e = a; return false; }
        anchors.push_back(a);
        return false;
    });
    std::sort(anchors.begin(), anchors.end(), [](const Anchor& x, const Anchor& y) { return x.pos < y.pos; });
    derived = true;
}

int HostRead::read_offset(int pos) const {
    if (pos < st.first || pos > st.last) return -1;
    int ans = -1;
    bool done = false;
    walk(st, [&](int i, int ref, int rd, int len, int op) {
        (void)op;
        if (!cig_both(st.cigar[i])) return false;
        if (pos < ref) { done = true; return true; }
        if (ref + len > pos) {
            int a = rd + pos - ref;
            ans = (a < 0 || a >= L) ? -1 : a;
            done = true;
            return true;
        }
        return false;
    });
    (void)done;
    return ans;
}

int HostRead::call_length(int pos, int* off_out) const {
    if (pos < st.first || pos > st.last) return 0;
    int result = 0;
    walk(st, [&](int i, int ref, int rd, int len, int op) {
        (void)op;
        if (!cig_both(st.cigar[i])) return false;
        if (pos < ref) return true;
        if (ref + len > pos) {
            int j = pos - ref;
            int rp = rd + j;
            if (rp < 0 || rp >= L) return true;
            if (off_out) *off_out = rp;
            if (skipped(*this, i, rd, j, len)) return true;
            bool next_indel = i + 1 < (int)st.cigar.size() && op_is_indel(cig_op(st.cigar[i + 1]));
            if (j == len - 1 && next_indel) {
                uint32_t nv = st.cigar[i + 1];
                result = cig_op(nv) == kOpI ? cig_len(nv) + 2 : 2;
            } else {
                result = 1;
            }
            return true;
        }
        return false;
    });
    return result;
}

const Anchor* HostRead::anchor_at(int pos) const {
    for (auto& a : anchors) if (a.pos == pos) return &a;
    return nullptr;
}

bool HostRead::has_anchor_in(int lo, int hi) const {
    for (auto& a : anchors) if (a.pos >= lo && a.pos <= hi) return true;
    return false;
}

bool HostRead::span_call(int pos, int last_pos, std::string& al, std::string& qs) const {
    if (call_length(pos) == 0) return false;                 // PR:135-137
    int rf = read_offset(pos), rl = read_offset(last_pos);
    if (rf < 0 || rl < 0 || rl < rf) return false;
    if (rf < st.ign_s || L - rl <= st.ign_e) return false;    // RA:1042-1044
    al.assign((const char*)bases + rf, rl - rf + 1);
    qs.assign((const char*)quals + rf, rl - rf + 1);
    return true;
}

bool HostRead::move_indel(int anchor_pos, int new_pos) {
    int disp = new_pos - anchor_pos;
    if (disp == 0) return true;
    auto& cg = st.cigar;
    int n = (int)cg.size();
    std::vector<uint32_t> out(n, 0u);
    int next_idx = -1;
    int ref = st.first;
    for (int i = 0; i < n; i++) {
        int len = cig_len(cg[i]), op = cig_op(cg[i]);
        if (op_is_indel(op) && ref == anchor_pos + 1) {
            if (i == 0 || i == n - 1) return false;
            uint32_t b = cg[i - 1], a = cg[i + 1];
            if (!cig_both(b) || cig_len(b) <= -disp) return false;
            if (!cig_both(a) || cig_len(a) <= disp) return false;
            out[i - 1] = cig_make(cig_len(b) + disp, cig_op(b));
            out[i] = cg[i];
            out[i + 1] = cig_make(cig_len(a) - disp, cig_op(a));
            next_idx = i + 1;
        } else if (i != next_idx) {
            out[i] = cg[i];
        }
        if (cig_ref(cg[i])) ref += len;
    }
    if (next_idx < 0) return false;
    cg.swap(out);
    derived = false;
    return true;
}

static void push_gap(std::vector<uint32_t>& v, int unk_read, int unk_ref) {
    int diff = unk_read - unk_ref;
    if (diff == 0) { if (unk_read > 0) v.push_back(cig_make(unk_read, kOpM)); }
    else if (diff > 0) { v.push_back(cig_make(diff, kOpI)); if (unk_ref > 0) v.push_back(cig_make(unk_ref, kOpM)); }
    else { v.push_back(cig_make(-diff, kOpD)); if (unk_read > 0) v.push_back(cig_make(unk_read, kOpM)); }
}

void HostRead::realign_start(int new_first, int first_match, int ref_after, int read_after) {
    std::vector<uint32_t> v;
    v.push_back(cig_make(first_match, kOpM));
    int unk_read = read_after - first_match;
    int unk_ref = ref_after - (new_first + first_match);
    if (unk_ref < 0 || unk_read < 0) return;
    push_gap(v, unk_read, unk_ref);
    int rd = 0;
    bool copy = false;
    for (uint32_t c : st.cigar) {
        int len = cig_len(c);
        if (copy) v.push_back(c);
        if (cig_read(c)) {
            if (!copy && read_after < rd + len) { copy = true; v.push_back(cig_make(rd + len - read_after, cig_op(c))); }
            rd += len;
        }
    }
    st.first = new_first;
    st.cigar.swap(v);
    derived = false;
}

void HostRead::realign_end(int ref_before, int read_before, int final_ref_start, int final_match) {
    int bp_end = L - read_before - 1;
    std::vector<uint32_t> v;
    int rd = 0;
    bool copy = true;
    for (uint32_t c : st.cigar) {
        int len = cig_len(c);
        if (cig_read(c)) {
            if (copy && read_before < rd + len) {
                int d = rd + len - read_before - 1;
                v.push_back(cig_make(len - d, cig_op(c)));
                copy = false;
            }
            rd += len;
        }
        if (copy) v.push_back(c);
    }
    int unk_ref = final_ref_start - ref_before - 1;
    int unk_read = bp_end - final_match;
    if (unk_ref < 0 || unk_read < 0) return;
    push_gap(v, unk_read, unk_ref);
    v.push_back(cig_make(final_match, kOpM));
    st.cigar.swap(v);
    st.last = final_ref_start + final_match - 1;
    derived = false;
}

// ------------------------------------------------------------------------------------------------------------
// admission
// ------------------------------------------------------------------------------------------------------------
namespace {
std::vector<HostRead> admit_range(const BatchView& b, int64_t lo, int64_t hi, const AdmissionOptions& opt);
}

// Everything admission does is local to the records of one start position — the duplicate-record rule compares equal starts
// (RAFR:292-306), the cap counts the alignments of one start (APG:407-433) — so a sequence cut where the start changes is
// admitted piece by piece, on `threads` host threads, with the pieces' reads concatenated in order.
std::vector<HostRead> admit_sequence(const BatchView& b, int64_t lo, int64_t hi, const AdmissionOptions& opt, int threads) {
    const int64_t n = hi - lo;
    if (threads < 4 || n < 200000) return admit_range(b, lo, hi, opt);   // (two threads lose to one: the pieces' reads are moved once more)
    const int nt = (int)std::min<int64_t>(threads, n / 50000);
    std::vector<int64_t> cut((size_t)nt + 1);
    cut[0] = lo; cut[(size_t)nt] = hi;
    for (int t = 1; t < nt; t++) {
        int64_t c = std::max(cut[(size_t)t - 1], lo + n * t / nt);
        while (c < hi && c > lo && b.first[c] == b.first[c - 1]) c++;   // never inside the records of one start position
        cut[(size_t)t] = c;
    }
    std::vector<std::vector<HostRead>> parts((size_t)nt);
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back([&, t]() { parts[(size_t)t] = admit_range(b, cut[(size_t)t], cut[(size_t)t + 1], opt); });
    parts[0] = admit_range(b, cut[0], cut[1], opt);
    for (auto& x : th) x.join();
    size_t total = 0;
    std::vector<size_t> base((size_t)nt);
    for (int t = 0; t < nt; t++) { base[(size_t)t] = total; total += parts[(size_t)t].size(); }
    std::vector<HostRead> out(total);
    th.clear();
    auto place = [&](int t) { std::move(parts[(size_t)t].begin(), parts[(size_t)t].end(), out.begin() + (ptrdiff_t)base[(size_t)t]); std::vector<HostRead>().swap(parts[(size_t)t]); };
    for (int t = 1; t < nt; t++) th.emplace_back(place, t);
    place(0);
    for (auto& x : th) x.join();
    return out;
}

namespace {
std::vector<HostRead> admit_range(const BatchView& b, int64_t lo, int64_t hi, const AdmissionOptions& opt) {
    std::vector<HostRead> out;
    out.reserve((size_t)(hi - lo));
    int filter = 0x4;
    if (!opt.secondary) { filter |= 0x100; if (!opt.non_unique_primary) filter |= 0x1000; }
    int64_t last_seen = -1;
    std::vector<HostRead> prim, sec;
    int group_start = INT_MIN;
    std::vector<std::pair<int, int>> per_rg;   // (read group, alignments admitted at this start): a handful of groups, searched linearly
    auto flush = [&]() {
        per_rg.clear();
        for (auto* lst : {&prim, &sec}) {
            for (auto& r : *lst) {
                auto it = per_rg.begin();
                while (it != per_rg.end() && it->first != r.rg) ++it;
                if (it == per_rg.end()) per_rg.emplace_back(r.rg, 1);
                else if (opt.max_alns_per_start <= 0 || it->second < opt.max_alns_per_start) it->second++;
                else continue;
                // 5' end is the alignment start for + strand reads, the end for - strand reads (RA:613-644)
                if (r.neg) { r.st.ign_e = opt.ignore5; r.st.ign_s = opt.ignore3; } else { r.st.ign_s = opt.ignore5; r.st.ign_e = opt.ignore3; }
                out.push_back(std::move(r));
            }
        }
        prim.clear();
        sec.clear();
    };
    for (int64_t i = lo; i < hi; i++) {
        int flags = b.flags[i];
        // duplicate record (RAFR:292-306): compared with the previous record of the file, whatever its fate
        if (last_seen >= 0 && b.name_hash) {
            int pf = b.flags[last_seen];
            bool same = b.first[last_seen] == b.first[i] && ((pf & 1) == (flags & 1)) && (!(pf & 1) || ((pf & 0x40) == (flags & 0x40))) &&
                        b.name_hash[last_seen] == b.name_hash[i];
            if (same) continue;
        }
        last_seen = i;
        bool multiple;
        int nh = b.nh ? b.nh[i] : -1;
        if (flags & 0x100) multiple = true;
        else if (nh > 1) multiple = true;
        else if (nh == 1) multiple = false;
        else multiple = b.mapq[i] < opt.min_mq;
        if (multiple) flags |= 0x1000;
        if (flags & filter) continue;
        int64_t c0 = b.cigar_off[i], c1 = b.cigar_off[i + 1];
        if (c1 <= c0) continue;  // mapped record without CIGAR: the reference's conversion throws and the record is skipped
        HostRead r;
        r.rec = i;
        r.L = (int)(b.base_off[i + 1] - b.base_off[i]);
        r.neg = (flags & 0x10) != 0;
        r.unique = !multiple;
        r.rg = b.rg ? b.rg[i] : -1;
        r.bases = b.bases + b.base_off[i];
        r.quals = b.quals + b.base_off[i];
        r.admit_first = r.st.first = b.first[i];
        // collapse equal neighbouring operations (RA:1296-1312) and derive read/reference extents
        int read_len = 0, ref_len = 0;
        for (int64_t k = c0; k < c1; k++) {
            uint32_t v = b.cigar[k];
            if (!r.st.cigar.empty() && cig_op(r.st.cigar.back()) == cig_op(v)) r.st.cigar.back() = cig_make(cig_len(r.st.cigar.back()) + cig_len(v), cig_op(v));
            else if (cig_len(v) > 0 || true) r.st.cigar.push_back(v);
            if (cig_read(v)) read_len += cig_len(v);
            if (cig_ref(v)) ref_len += cig_len(v);
        }
        if (read_len != r.L || r.L == 0) continue;  // SEQ '*' or inconsistent with the CIGAR: outside the supported input
        r.st.last = r.st.first + ref_len - 1;
        if (opt.query_first > r.st.last) continue;  // APG:316
        if (r.st.first != group_start) { flush(); group_start = r.st.first; }
        if (flags & 0x100) sec.push_back(std::move(r)); else prim.push_back(std::move(r));
    }
    flush();
    return out;
}
}  // namespace

// ------------------------------------------------------------------------------------------------------------
// admission without materialising the reads
// ------------------------------------------------------------------------------------------------------------
namespace {
inline bool flagged_multiple(const BatchView& b, int64_t i, int min_mq) {   // RAFR:229-236, 284-291
    const int flags = b.flags[i];
    const int nh = b.nh ? b.nh[i] : -1;
    if (flags & 0x100) return true;
    if (nh > 1) return true;
    if (nh == 1) return false;
    return b.mapq[i] < min_mq;
}

std::vector<LiteRead> admit_range_lite(const BatchView& b, int64_t base, int64_t lo, int64_t hi, const AdmissionOptions& opt) {
    std::vector<LiteRead> out;
    out.reserve((size_t)(hi - lo));
    int filter = 0x4;
    if (!opt.secondary) { filter |= 0x100; if (!opt.non_unique_primary) filter |= 0x1000; }
    int64_t last_seen = -1;
    struct Pending { LiteRead l; int rg; };
    std::vector<Pending> prim, sec;
    int group_start = INT_MIN;
    std::vector<std::pair<int, int>> per_rg;
    auto flush = [&]() {
        per_rg.clear();
        for (auto* lst : {&prim, &sec}) {
            for (auto& r : *lst) {
                auto it = per_rg.begin();
                while (it != per_rg.end() && it->first != r.rg) ++it;
                if (it == per_rg.end()) per_rg.emplace_back(r.rg, 1);
                else if (opt.max_alns_per_start <= 0 || it->second < opt.max_alns_per_start) it->second++;
                else continue;
                out.push_back(r.l);
            }
        }
        prim.clear();
        sec.clear();
    };
    for (int64_t i = lo; i < hi; i++) {
        int flags = b.flags[i];
        if (last_seen >= 0 && b.name_hash) {   // duplicate record (RAFR:292-306)
            const int pf = b.flags[last_seen];
            const bool same = b.first[last_seen] == b.first[i] && ((pf & 1) == (flags & 1)) && (!(pf & 1) || ((pf & 0x40) == (flags & 0x40))) &&
                              b.name_hash[last_seen] == b.name_hash[i];
            if (same) continue;
        }
        last_seen = i;
        if (flagged_multiple(b, i, opt.min_mq)) flags |= 0x1000;
        if (flags & filter) continue;
        const int64_t c0 = b.cigar_off[i], c1 = b.cigar_off[i + 1];
        if (c1 <= c0) continue;
        const int L = (int)(b.base_off[i + 1] - b.base_off[i]);
        int read_len = 0, ref_len = 0;
        for (int64_t k = c0; k < c1; k++) {
            const uint32_t v = b.cigar[k];
            if (cig_read(v)) read_len += cig_len(v);
            if (cig_ref(v)) ref_len += cig_len(v);
        }
        if (read_len != L || L == 0) continue;
        Pending p;
        p.l.rec = (int32_t)(i - base);
        p.l.first = b.first[i];
        p.l.last = p.l.first + ref_len - 1;
        p.l.heavy = (c1 - c0 == 1 && cig_op(b.cigar[c0]) == kOpM) ? kLiteSimple : kLiteComplex;
        p.rg = b.rg ? b.rg[i] : -1;
        if (opt.query_first > p.l.last) continue;  // APG:316
        if (p.l.first != group_start) { flush(); group_start = p.l.first; }
        if (flags & 0x100) sec.push_back(p); else prim.push_back(p);
    }
    flush();
    return out;
}
}  // namespace

HostRead make_read(const BatchView& b, int64_t i, const AdmissionOptions& opt) {
    HostRead r;
    const int flags = b.flags[i];
    r.rec = i;
    r.L = (int)(b.base_off[i + 1] - b.base_off[i]);
    r.neg = (flags & 0x10) != 0;
    r.unique = !flagged_multiple(b, i, opt.min_mq);
    r.rg = b.rg ? b.rg[i] : -1;
    r.bases = b.bases + b.base_off[i];
    r.quals = b.quals + b.base_off[i];
    r.admit_first = r.st.first = b.first[i];
    int ref_len = 0;
    for (int64_t k = b.cigar_off[i]; k < b.cigar_off[i + 1]; k++) {   // equal neighbouring operations collapse (RA:1296-1312)
        const uint32_t v = b.cigar[k];
        if (!r.st.cigar.empty() && cig_op(r.st.cigar.back()) == cig_op(v)) r.st.cigar.back() = cig_make(cig_len(r.st.cigar.back()) + cig_len(v), cig_op(v));
        else r.st.cigar.push_back(v);
        if (cig_ref(v)) ref_len += cig_len(v);
    }
    r.st.last = r.st.first + ref_len - 1;
    if (r.neg) { r.st.ign_e = opt.ignore5; r.st.ign_s = opt.ignore3; } else { r.st.ign_s = opt.ignore5; r.st.ign_e = opt.ignore3; }   // RA:613-644
    return r;
}

std::vector<LiteRead> admit_sequence_lite(const BatchView& b, int64_t lo, int64_t hi, const AdmissionOptions& opt, int threads) {
    const int64_t n = hi - lo;
    if (threads < 2 || n < 400000) return admit_range_lite(b, lo, lo, hi, opt);
    const int nt = (int)std::min<int64_t>(threads, n / 100000);
    std::vector<int64_t> cut((size_t)nt + 1);
    cut[0] = lo; cut[(size_t)nt] = hi;
    for (int t = 1; t < nt; t++) {
        int64_t c = std::max(cut[(size_t)t - 1], lo + n * t / nt);
        while (c < hi && c > lo && b.first[c] == b.first[c - 1]) c++;   // never inside the records of one start position
        cut[(size_t)t] = c;
    }
    std::vector<std::vector<LiteRead>> parts((size_t)nt);
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back([&, t]() { parts[(size_t)t] = admit_range_lite(b, lo, cut[(size_t)t], cut[(size_t)t + 1], opt); });
    parts[0] = admit_range_lite(b, lo, cut[0], cut[1], opt);
    for (auto& x : th) x.join();
    size_t total = 0;
    for (auto& p : parts) total += p.size();
    std::vector<LiteRead> out;
    out.reserve(total);
    for (auto& p : parts) out.insert(out.end(), p.begin(), p.end());
    return out;
}

// ------------------------------------------------------------------------------------------------------------
// several inputs: APG.processFiles
// ------------------------------------------------------------------------------------------------------------
namespace {
void reader_survivors(const BatchView& b, const AdmissionOptions& opt, std::vector<int64_t>& keep, std::vector<int>& last) {
    int filter = 0x4;
    if (!opt.secondary) { filter |= 0x100; if (!opt.non_unique_primary) filter |= 0x1000; }
    int64_t prev = -1;
    for (int64_t i = 0; i < b.n; i++) {
        int flags = b.flags[i];
        if (prev >= 0 && b.name_hash) {   // duplicate record, RAFR:292-306
            int pf = b.flags[prev];
            bool same = b.first[prev] == b.first[i] && ((pf & 1) == (flags & 1)) && (!(pf & 1) || ((pf & 0x40) == (flags & 0x40))) && b.name_hash[prev] == b.name_hash[i];
            if (same) continue;
        }
        prev = i;
        bool multiple;
        int nh = b.nh ? b.nh[i] : -1;
        if (flags & 0x100) multiple = true; else if (nh > 1) multiple = true; else if (nh == 1) multiple = false; else multiple = b.mapq[i] < opt.min_mq;
        if (multiple) flags |= 0x1000;
        if (flags & filter) continue;
        if (b.cigar_off[i + 1] <= b.cigar_off[i]) continue;
        int ref_len = 0;
        for (int64_t k = b.cigar_off[i]; k < b.cigar_off[i + 1]; k++) if (cig_ref(b.cigar[k])) ref_len += cig_len(b.cigar[k]);
        keep.push_back(i);
        last.push_back(b.first[i] + ref_len - 1);
    }
}
}  // namespace

void merge_order(const std::vector<BatchView>& inputs, const AdmissionOptions& opt, std::vector<std::pair<int, int64_t>>& out) {
    const size_t n = inputs.size();
    std::vector<std::vector<int64_t>> keep(n);
    std::vector<std::vector<int>> last(n);
    {   // the inputs are independent up to here: one thread each (a population run has a file per sample)
        const unsigned hw = std::thread::hardware_concurrency();
        const size_t nt = std::min<size_t>(n, std::max(1u, std::min(32u, hw ? hw : 1u)));
        std::atomic<size_t> next{0};
        auto work = [&]() { for (;;) { const size_t f = next.fetch_add(1); if (f >= n) return; reader_survivors(inputs[f], opt, keep[f], last[f]); } };
        std::vector<std::thread> th;
        for (size_t t = 1; t < nt; t++) th.emplace_back(work);
        work();
        for (auto& x : th) x.join();
    }
    // chooseNextAln (APG:262-289) looks at the head of every input and keeps the smallest by (sequence, first, last), the lowest input
    // on ties: a heap over the heads with the input index as the last key gives the same order in O(log n) per record
    struct Head { int seq, first, last, file; };
    auto after = [](const Head& a, const Head& b) {   // a comes after b
        if (a.seq != b.seq) return a.seq > b.seq;
        if (a.first != b.first) return a.first > b.first;
        if (a.last != b.last) return a.last > b.last;
        return a.file > b.file;
    };
    std::priority_queue<Head, std::vector<Head>, decltype(after)> heap(after);
    std::vector<size_t> head(n, 0);
    auto push = [&](size_t f) {
        if (head[f] >= keep[f].size()) return;
        const int64_t i = keep[f][head[f]];
        heap.push(Head{inputs[f].seq_idx[i], inputs[f].first[i], last[f][head[f]], (int)f});
    };
    size_t total = 0;
    for (size_t f = 0; f < n; f++) { total += keep[f].size(); push(f); }
    out.clear();
    out.reserve(total);
    while (!heap.empty()) {
        const size_t f = (size_t)heap.top().file;
        heap.pop();
        out.push_back({(int)f, keep[f][head[f]]});
        head[f]++;
        push(f);
    }
}

}  // namespace host
}  // namespace ngsep_b200
