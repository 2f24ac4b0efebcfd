// Indel-realigner logic applied at ONE event position (a position where some covering read anchors an indel).
// Restates IndelRealignerPileupListener.onPileup; the input variants of -knownVariants / -knownSTRs reach it as loci
// (IRPL = /root/reference/src/ngsep/discovery/IndelRealignerPileupListener.java).  At every other position the
// listener only sets referenceSpan = 1 (IRPL:87-124), which is why the host can skip them.
#include <algorithm>
#include <map>
#include <set>

#include "pipeline.hpp"

namespace ngsep_b200 {
namespace host {

namespace {

int hamming(const std::string& a, const char* b) {
    int d = 0;
    for (size_t i = 0; i < a.size(); i++) d += a[i] != b[i];
    return d;
}

// sequences/HammingSequenceDistanceMeasure.java:88-101 with CountsRankHelper's TreeMap + stable sort: most frequent
// character per column, ties to the smallest character.
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

// Iteration order of java.util.HashMap<Integer,?> for small distinct keys inserted in the given order.
std::vector<int> hashmap_order(const std::vector<int>& keys) {
    size_t cap = 16;
    while (keys.size() > cap * 3 / 4) cap *= 2;
    std::vector<std::pair<unsigned, int>> v;
    for (int k : keys) { unsigned h = (unsigned)k; h ^= h >> 16; v.push_back({h & (unsigned)(cap - 1), k}); }
    std::stable_sort(v.begin(), v.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    std::vector<int> out;
    for (auto& p : v) out.push_back(p.second);
    return out;
}

int mononucleotide_run(const std::string& s) {  // IRPL:365-391
    int cnt[4] = {0, 0, 0, 0};
    const int min_len = 5;
    int n = (int)s.size(), i = 0;
    for (; i < n && i < min_len; i++) { int j = dna_index(s[i]); if (j >= 0 && s[i] < 'a') cnt[j]++; }
    int base = -1;
    for (int j = 0; j < 4; j++) if (cnt[j] >= i - 1) { base = j; break; }
    if (base < 0 || i < min_len) return 0;
    while (i < n && cnt[base] >= i - 1) { int j = dna_index(s[i]); if (j >= 0 && s[i] < 'a') cnt[j]++; i++; }
    i--;
    int prev = (s[i - 1] < 'a') ? dna_index(s[i - 1]) : -1;
    return prev != base ? i - 1 : i;
}

}  // namespace

EventOutcome Realigner::on_event(int pos, std::vector<HostRead*>& pileup, const StrLocus* locus) {
    EventOutcome out;
    int event_end;
    if (locus) {   // IRPL:90-95: the pileup is an input STR, the event is its locus
        out.fired = true;
        out.str = locus->str;
        out.span = locus->last - locus->first + 1;
        event_end = locus->last;
        if (event_end <= pos) return out;   // IRPL:118: nothing to conciliate
    } else {
        int max_len = 0, max_span = 0;
        for (HostRead* r : pileup) {
            const Anchor* a = r->anchor_at(pos);
            if (!a) continue;
            max_len = std::max(max_len, a->len);
            max_span = std::max(max_span, a->last - a->pos + 1);
        }
        if (max_len == 0) return out;  // not an event after all (stale queue entry)
        out.fired = true;
        event_end = pos + std::max(max_len, max_span) + 1;  // IRPL:116
    }
    const int locus_span = out.span;

    // ---- conciliateIndels, IRPL:165-216 -------------------------------------------------------------------
    std::set<int> lengths;
    std::vector<HostRead*> indel_reads;
    std::vector<int> votes(event_end - pos + 1, 0);
    int longest = 0;
    for (HostRead* r : pileup) {  // analyzeIndels, IRPL:229-265
        for (const Anchor& a : r->anchors) {
            if (a.last >= pos && a.pos <= event_end) {
                lengths.insert(a.len);
                longest = std::max(longest, a.len);
                int i = a.pos - pos;
                if (i >= 0 && i < (int)votes.size()) votes[i]++;
                indel_reads.push_back(r);
                break;
            }
        }
    }
    if (lengths.empty()) return out;
    int best_vote = 0;
    for (int i = 1; i < (int)votes.size(); i++) if (votes[best_vote] < votes[i]) best_vote = i;  // first strict maximum
    int answer = 0;
    bool fixed = locus != nullptr;
    if (fixed) best_vote = 0;   // IRPL:183-184: no vote, no search for a new tandem repeat
    if (!fixed && lengths.size() > 1) {  // lookForNewSTR, IRPL:315-349
        int new_span = 0;
        std::string ref;
        int run = 0;
        if (genome->get(seq, pos + 1, indel_reads.back()->st.last, ref)) run = mononucleotide_run(upper(ref));
        if (run > 0) new_span = run + 2;
        else {
            for (HostRead* r : indel_reads) {
                const Anchor* a = r->anchor_at(pos);
                if (!a || a->len != longest) continue;
                int off = r->read_offset(pos);
                int tr = 0;
                if (off >= 0) tr = mononucleotide_run(upper(std::string((const char*)r->bases + off + 1, r->L - off - 1)));
                if (tr > 0) {
                    int sp = a->last - a->pos + 1;
                    new_span = (a->len >= sp) ? a->len + 2 : tr + sp;
                    break;
                }
            }
        }
        if (new_span > 1) {
            best_vote = 0;
            answer = new_span;
            event_end = pos + answer - 1;
            fixed = true;
            out.str = out.new_str = true;
        }
    }
    // moveIndelStarts, IRPL:274-313
    int new_end = pos + 1;
    for (HostRead* r : indel_reads) {
        for (const Anchor& a : r->anchors) {
            if (a.last >= pos && a.pos <= event_end) {
                int from = a.pos;
                r->move_indel(from, pos + best_vote);
                break;
            }
        }
        if (!r->derived) r->derive();
        int ref_last = pos;
        for (const Anchor& a : r->anchors) if (a.pos >= pos && a.pos <= ref_last + longest) ref_last = a.last;
        new_end = std::max(new_end, ref_last);
    }
    if (best_vote > 0) { out.span = answer > 0 ? answer : locus_span; return out; }
    if (!fixed && new_end != event_end) {
        event_end = new_end;
        answer = event_end - pos + 1;
    }
    process_ends(pileup, pos, event_end);
    out.span = answer > 0 ? answer : locus_span;
    return out;
}

// processEndsOfAlignments, IRPL:400-526
void Realigner::process_ends(std::vector<HostRead*>& pileup, int ev_first, int ev_last) {
    std::string before, after, within;
    bool has_before = genome->get(seq, ev_first - 100, ev_first, before);
    bool has_after = genome->get(seq, ev_last, ev_last + 100, after);
    bool has_within = (ev_first != ev_last - 1) && genome->get(seq, ev_first + 1, ev_last - 1, within);
    before = upper(before); after = upper(after); within = upper(within);
    std::string ref_before = before, ref_after = after;
    if (has_before && has_within) ref_before += within;
    if (has_after && has_within) ref_after = within + ref_after;

    // inserted consensus (IRPL:528-555) or consensus deletion length (IRPL:557-578)
    std::string ins;
    bool has_ins = false;
    {
        std::map<int, std::vector<std::string>> by_len;
        std::vector<int> order;
        for (HostRead* r : pileup) {
            if (!r->anchor_at(ev_first)) continue;
            int off = -1;
            int cl = r->call_length(ev_first, &off);
            if (cl == 0) continue;
            if (cl <= 2) continue;  // allele without inner bases
            std::string inner((const char*)r->bases + off + 1, cl - 2);
            if (!by_len.count((int)inner.size())) order.push_back((int)inner.size());
            by_len[(int)inner.size()].push_back(inner);
        }
        size_t best = 0;
        const std::vector<std::string>* pick = nullptr;
        for (int l : hashmap_order(order)) if (by_len[l].size() > best) { best = by_len[l].size(); pick = &by_len[l]; }
        if (pick) { ins = column_consensus(*pick); has_ins = true; }
    }
    std::string alt_before = before, alt_after = after;
    int offset = 0;
    if (has_ins) {
        offset = (int)ins.size();
        if (has_before) alt_before += ins;
        if (has_before && has_within) alt_before += within;
        if (has_after && has_within) alt_after = within + alt_after;
        if (has_after) alt_after = ins + alt_after;
    } else {
        std::map<int, int> counts;
        std::vector<int> order;
        for (HostRead* r : pileup) {
            const Anchor* a = r->anchor_at(ev_first);
            if (!a) continue;
            int inner = a->last - a->pos - 1;
            if (!counts.count(inner)) order.push_back(inner);
            counts[inner]++;
        }
        int del = 0, mx = 0;
        for (int l : hashmap_order(order)) if (mx < counts[l]) { del = l; mx = counts[l]; }
        if (del > ev_last - ev_first - 1) del = ev_last - ev_first - 1;
        offset = -del;
        if (has_within) {
            std::string rest = (del == 0) ? within : (del < (int)within.size() ? within.substr(del) : std::string());
            if (has_before) alt_before += rest;
            if (has_after) alt_after = rest + alt_after;
        }
    }
    const int min_good = 5, max_realign = 50;  // IRPL:46-47
    for (HostRead* r : pileup) {
        const int aln_first = r->st.first, aln_last = r->st.last;
        if (!r->derived) r->derive();
        bool indels_before = r->has_anchor_in(aln_first, ev_first - 1);
        bool indels_after = r->has_anchor_in(ev_last + 1, aln_last);
        int bp_good = std::max(offset, min_good);
        bool trim_start = ev_first - aln_first < bp_good && !indels_before;
        int pos_after = r->read_offset(ev_last);
        if (!indels_before && has_before && pos_after >= bp_good && pos_after - offset <= max_realign && pos_after < (int)ref_before.size() &&
            pos_after < (int)alt_before.size() && !r->anchor_at(ev_first)) {
            int d_ref = hamming(ref_before.substr(ref_before.size() - pos_after), (const char*)r->bases);
            int d_alt = hamming(alt_before.substr(alt_before.size() - pos_after), (const char*)r->bases);
            int new_first = ev_last - pos_after + 1 + offset;
            int first_match = ev_first - new_first + 1;
            if (d_alt < d_ref && d_alt < 3 && first_match >= min_good) {
                r->realign_start(new_first, first_match, ev_last, pos_after);
                trim_start = false;
            }
        }
        if (trim_start) {
            int ignore = ev_last - aln_first + 1 + r->soft_clip_start();
            r->st.ign_s = (int8_t)std::max(r->st.ign_s, ignore);  // the reference narrows to byte here (IRPL:478)
            r->derived = false;
        }
        if (!r->derived) r->derive();
        bool trim_end = aln_last - ev_last < bp_good && !indels_after;
        int pos_before = r->read_offset(ev_first);
        int suffix = pos_before >= 0 ? r->L - pos_before - 1 : 0;
        if (!indels_after && has_after && suffix >= bp_good && suffix - offset <= max_realign && suffix < (int)ref_after.size() &&
            suffix < (int)alt_after.size() && (!r->anchor_at(ev_first) || pos_after < 0)) {
            const char* tail = (const char*)r->bases + pos_before + 1;
            int d_ref = hamming(ref_after.substr(0, suffix), tail);
            int d_alt = hamming(alt_after.substr(0, suffix), tail);
            int final_match = suffix - (offset > 0 ? offset : 0);
            int new_ev_last = ev_first + 1 + (offset < 0 ? -offset : 0);
            if (d_alt < d_ref && d_alt < 3 && final_match >= min_good) {
                r->realign_end(ev_first, pos_before, new_ev_last, final_match);
                trim_end = false;
            }
        }
        if (trim_end) {
            int ignore = aln_last - ev_first + 1 + r->soft_clip_end();
            r->st.ign_e = (int8_t)std::max(r->st.ign_e, ignore);  // IRPL:520
            r->derived = false;
        }
        if (!r->derived) r->derive();
    }
}

}  // namespace host
}  // namespace ngsep_b200
