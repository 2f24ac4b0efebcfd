// TEST INFRASTRUCTURE ONLY — see ngsep_oracle.hpp.  Part 3: IndelRealignerPileupListener,
// SingleSampleVariantPileupListener, AlignmentsPileupGenerator, VCF text, SingleSampleVariantsDetector.
#include "ngsep_oracle.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>

namespace ngsep_oracle {

static inline int dna_index(char c) {
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; }
    return -1;
}
static std::string to_upper(const std::string& s) {
    std::string o = s;
    for (auto& c : o) if (c >= 'a' && c <= 'z') c = (char)(c - 32);
    return o;
}
static double hamming(const std::string& a, const std::string& b) {  // HammingSequenceDistanceMeasure.java:43-51
    int d = 0;
    for (size_t i = 0; i < a.size(); i++) if (a[i] != b[i]) d++;
    return d;
}

// ---------------------------------------------------------------------------------------------
// IndelRealignerPileupListener (no input variants: -knownVariants / -knownSTRs are "next" rows)
// ---------------------------------------------------------------------------------------------
const InputVariant* IndelRealignerPileupListener::intersect_with_variants(int pos) {  // IRPL:141-157
    if (!seq_variants) return nullptr;
    while (idx_next_variant < seq_variants->size()) {
        const InputVariant& v = (*seq_variants)[idx_next_variant];
        if (pos < v.first) break;
        if (v.first <= pos && pos <= v.last) return &v;
        idx_next_variant++;
    }
    return nullptr;
}

void IndelRealignerPileupListener::on_pileup(PileupRecord& p) {  // IRPL:84-126
    int cur = p.pos;
    int span = 1;
    int predicted_end = cur;
    const InputVariant* var = intersect_with_variants(cur);
    if (var) {
        if (var->first == cur) {
            if (var->str) p.str = true;
            span = var->last - var->first + 1;
            predicted_end = var->last;
        } else {
            p.embedded = true;
        }
    }
    std::vector<ReadAlignment*>& alns = p.alns;
    if (!var) {
        int max_len = 0, max_span = 0;
        for (ReadAlignment* a : alns) {
            const IndelCall* indel = a->get_indel_call(cur);
            if (indel) {
                int isp = indel->last - indel->first + 1;
                int il = indel->length;
                if (il > max_len) max_len = il;
                if (isp > max_span) max_span = isp;
            }
        }
        if (max_len > 0) predicted_end = cur + std::max(max_len, max_span) + 1;
    }
    if (predicted_end > cur) {
        n_events++;
        int cs = conciliate(p, alns, predicted_end, var != nullptr);
        if (cs > 0) span = cs;
    }
    p.span = span;
}

int IndelRealignerPileupListener::analyze(std::vector<ReadAlignment*>& alns, int ev_start, int ev_end, std::set<int>& lengths,
                                          std::vector<ReadAlignment*>& indel_alns, std::vector<int>& votes) {  // IRPL:229-265
    int max_length = 0;
    lengths.clear();
    indel_alns.clear();
    std::fill(votes.begin(), votes.end(), 0);
    for (ReadAlignment* a : alns) {
        bool found = false;
        auto* indels = a->get_indel_calls();
        if (indels) {
            for (auto& kv : *indels) {
                int start = kv.first;
                const IndelCall& indel = kv.second;
                if (indel.last >= ev_start && start <= ev_end) {
                    found = true;
                    int length = indel.length;
                    lengths.insert(length);
                    if (max_length < length) max_length = length;
                    int i = start - ev_start;
                    if (i >= 0 && i < (int)votes.size()) votes[i]++;
                    break;
                }
            }
        }
        if (found) indel_alns.push_back(a);
    }
    return max_length;
}

int IndelRealignerPileupListener::move_starts(std::vector<ReadAlignment*>& alns, int first, int last, int max_len, int offset) {  // IRPL:274-313
    int answer = first + 1;
    for (ReadAlignment* a : alns) {
        auto* indels = a->get_indel_calls();
        if (indels) {
            for (auto& kv : *indels) {
                int start = kv.first;
                if (kv.second.last >= first && start <= last) {
                    a->move_indel_start(start, first + offset);
                    break;
                }
            }
        }
        indels = a->get_indel_calls();
        if (indels) {
            int aln_ref_last = first;
            for (auto& kv : *indels) {
                int start = kv.first;
                if (start >= first && start <= aln_ref_last + max_len) aln_ref_last = kv.second.last;
            }
            if (aln_ref_last > answer) answer = aln_ref_last;
        }
    }
    return answer;
}

int check_mononucleotide(const std::string& seq) {  // IRPL:365-391
    int counts[4] = {0, 0, 0, 0};
    int min_length = 5;
    int i = 0;
    int n = (int)seq.size();
    while (i < n && i < min_length) {
        int j = dna_index(seq[i]);
        if (j >= 0) counts[j]++;
        i++;
    }
    int base_idx = -1;
    for (int j = 0; j < 4; j++) if (counts[j] >= i - 1) { base_idx = j; break; }
    if (base_idx == -1 || i < min_length) return 0;
    while (i < n && counts[base_idx] >= i - 1) {
        int j = dna_index(seq[i]);
        if (j >= 0) counts[j]++;
        i++;
    }
    i--;
    if (dna_index(seq[i - 1]) != base_idx) return i - 1;
    return i;
}

int IndelRealignerPileupListener::look_for_new_str(PileupRecord& p, std::vector<ReadAlignment*>& alns, int max_len) {  // IRPL:315-349
    int new_span = 0;
    int cur = p.pos;
    if (alns.empty()) return 0;
    std::string seq;
    int length_ref = 0;
    if (genome->get(p.seq, cur + 1, alns.back()->last, seq)) length_ref = check_mononucleotide(to_upper(seq));
    if (length_ref > 0) return length_ref + 2;
    for (ReadAlignment* a : alns) {
        const IndelCall* indel = a->get_indel_call(cur);
        if (indel && indel->length == max_len) {
            int isp = indel->last - indel->first + 1;
            int il = indel->length;
            // checkTandemRepeat(aln,currentPos) IRPL:351-357
            int length_tr = 0;
            int rf = a->aligned_read_position(cur);
            if (rf >= 0) length_tr = check_mononucleotide(to_upper(a->bases.substr(rf + 1)));
            if (length_tr > 0) {
                if (il >= isp) new_span = il + 2;
                else new_span = length_tr + isp;
                break;
            }
        }
    }
    return new_span;
}

int IndelRealignerPileupListener::conciliate(PileupRecord& p, std::vector<ReadAlignment*>& alns, int event_end, bool fixed_event) {  // IRPL:165-216
    int answer = 0;
    int cur = p.pos;
    bool fixed = fixed_event;
    std::set<int> lengths;
    std::vector<ReadAlignment*> indel_alns;
    std::vector<int> votes(event_end - cur + 1, 0);
    int max_length = analyze(alns, cur, event_end, lengths, indel_alns, votes);
    if (lengths.empty()) return answer;
    if (min_prop_alns_realign > 0 && indel_alns.size() < min_prop_alns_realign * alns.size()) return answer;
    int max_i = 0;
    if (!fixed) {
        // NumberArrays.getIndexMaximum, math/NumberArrays.java:87-95: first strict maximum
        int idx = -1;
        for (int i = 0; i < (int)votes.size(); i++) if (idx == -1 || votes[idx] < votes[i]) idx = i;
        max_i = idx;
        if (lengths.size() > 1) {
            int new_span = look_for_new_str(p, indel_alns, max_length);
            if (new_span > 1) {
                max_i = 0;
                answer = new_span;
                event_end = cur + answer - 1;
                fixed = true;
                p.str = true;
                p.new_str = true;
            }
        }
    }
    int new_end = move_starts(indel_alns, cur, event_end, max_length, max_i);
    if (max_i > 0) return answer;
    if (!fixed && new_end != event_end) {
        event_end = new_end;
        answer = event_end - cur + 1;
    }
    process_ends(alns, p.seq, cur, event_end);
    return answer;
}

static bool inserted_consensus(std::vector<ReadAlignment*>& alns, int ev_first, std::string& out) {  // IRPL:528-555
    std::map<int, std::vector<std::string>> by_len;
    std::vector<int> insertion_order;
    for (ReadAlignment* a : alns) {
        if (!a->get_indel_call(ev_first)) continue;
        std::string allele;
        if (!a->allele_call(ev_first, allele)) continue;
        allele = allele.substr(1, allele.size() - 2);
        if (allele.empty()) continue;
        int l = (int)allele.size();
        if (!by_len.count(l)) insertion_order.push_back(l);
        by_len[l].push_back(allele);
    }
    size_t best_n = 0;
    const std::vector<std::string>* best = nullptr;
    for (int l : java_int_hashmap_order(insertion_order)) {  // HashMap.values() iteration
        auto& v = by_len[l];
        if (v.size() > best_n) { best_n = v.size(); best = &v; }
    }
    if (!best) return false;
    out = hamming_consensus(*best);
    return true;
}

static int deletion_consensus_length(std::vector<ReadAlignment*>& alns, int ev_first) {  // IRPL:557-578
    std::map<int, int> counts;
    std::vector<int> insertion_order;
    for (ReadAlignment* a : alns) {
        const IndelCall* call = a->get_indel_call(ev_first);
        if (!call) continue;
        int inner = call->last - call->first - 1;
        if (!counts.count(inner)) insertion_order.push_back(inner);
        counts[inner]++;
    }
    int mx = 0, answer = 0;
    for (int l : java_int_hashmap_order(insertion_order)) {
        int c = counts[l];
        if (mx < c) { answer = l; mx = c; }
    }
    return answer;
}

void IndelRealignerPileupListener::process_ends(std::vector<ReadAlignment*>& alns, const std::string& seq_name, int ev_first, int ev_last) {  // IRPL:400-526
    std::string seq_before, seq_after, seq_within;
    bool has_before = genome->get(seq_name, ev_first - 100, ev_first, seq_before);
    bool has_after = genome->get(seq_name, ev_last, ev_last + 100, seq_after);
    bool has_within = false;
    if (ev_first != ev_last - 1) has_within = genome->get(seq_name, ev_first + 1, ev_last - 1, seq_within);
    if (has_within) seq_within = to_upper(seq_within);
    std::string ref_before = has_before ? to_upper(seq_before) : std::string();
    std::string ref_after = has_after ? to_upper(seq_after) : std::string();
    if (has_before && has_within) ref_before += seq_within;
    if (has_after && has_within) ref_after = seq_within + ref_after;

    std::string ins_cons;
    bool has_ins = inserted_consensus(alns, ev_first, ins_cons);
    std::string alt_before = has_before ? to_upper(seq_before) : std::string();
    std::string alt_after = has_after ? to_upper(seq_after) : std::string();
    int offset = 0;
    if (has_ins) {
        offset = (int)ins_cons.size();
        if (has_before) alt_before += ins_cons;
        if (has_before && has_within) alt_before += seq_within;
        if (has_after && has_within) alt_after = seq_within + alt_after;
        if (has_after) alt_after = ins_cons + alt_after;
    } else {
        int del_len = deletion_consensus_length(alns, ev_first);
        if (del_len > ev_last - ev_first - 1) del_len = ev_last - ev_first - 1;
        offset = -del_len;
        if (has_before && has_within) {
            if (del_len == 0) alt_before += seq_within;
            else if (del_len < (int)seq_within.size()) alt_before += seq_within.substr(del_len);
        }
        if (has_after && has_within) {
            if (del_len == 0) alt_after = seq_within + alt_after;
            else if (del_len < (int)seq_within.size()) alt_after = seq_within.substr(del_len) + alt_after;
        }
    }
    for (ReadAlignment* a : alns) {
        int aln_first = a->first;
        int aln_last = a->last;
        bool indels_before = a->has_indel_calls_in(aln_first, ev_first - 1);
        bool indels_after = a->has_indel_calls_in(ev_last + 1, aln_last);
        int bp_good = std::max(offset, min_bp_good_ref_aln);
        bool trim_start = ev_first - aln_first < bp_good && !indels_before;
        int read_pos_after = a->aligned_read_position(ev_last);
        if (!indels_before && has_before && read_pos_after >= bp_good && read_pos_after - offset <= max_bp_realign_end &&
            read_pos_after < (int)ref_before.size() && read_pos_after < (int)alt_before.size() && a->get_indel_call(ev_first) == nullptr) {
            std::string read_prefix = a->bases.substr(0, read_pos_after);
            std::string ref_suffix = ref_before.substr(ref_before.size() - read_pos_after);
            double ref_dist = hamming(ref_suffix, read_prefix);
            std::string alt_suffix = alt_before.substr(alt_before.size() - read_pos_after);
            double alt_dist = hamming(alt_suffix, read_prefix);
            int new_first = ev_last - read_pos_after + 1 + offset;
            int first_match = ev_first - new_first + 1;
            if (alt_dist < ref_dist && alt_dist < 3 && first_match >= min_bp_good_ref_aln) {
                a->realign_start(new_first, first_match, ev_last, read_pos_after);
                trim_start = false;
            }
        }
        if (trim_start) {
            int ignore_bp = ev_last - aln_first + 1;
            ignore_bp += a->soft_clip_start();
            int8_t bp = (int8_t)std::max((int)a->ignore_start, ignore_bp);  // (byte) cast, IRPL:478
            a->set_ignore_start(bp);
        }
        bool trim_end = aln_last - ev_last < bp_good && !indels_after;
        int read_pos_before = a->aligned_read_position(ev_first);
        int suffix_len = 0;
        if (read_pos_before >= 0) suffix_len = a->read_length - read_pos_before - 1;
        if (!indels_after && has_after && suffix_len >= bp_good && suffix_len - offset <= max_bp_realign_end &&
            suffix_len < (int)ref_after.size() && suffix_len < (int)alt_after.size() && (a->get_indel_call(ev_first) == nullptr || read_pos_after < 0)) {
            std::string read_suffix = a->bases.substr(read_pos_before + 1, a->read_length - (read_pos_before + 1));
            std::string ref_prefix = ref_after.substr(0, suffix_len);
            double ref_dist = hamming(ref_prefix, read_suffix);
            std::string alt_prefix = alt_after.substr(0, suffix_len);
            double alt_dist = hamming(alt_prefix, read_suffix);
            int final_match = suffix_len;
            if (offset > 0) final_match -= offset;
            int new_ev_last = ev_first + 1;
            if (offset < 0) new_ev_last -= offset;
            if (alt_dist < ref_dist && alt_dist < 3 && final_match >= min_bp_good_ref_aln) {
                a->realign_end(ev_first, read_pos_before, new_ev_last, final_match);
                trim_end = false;
            }
        }
        if (trim_end) {
            int ignore_bp = aln_last - ev_first + 1;
            ignore_bp += a->soft_clip_end();
            int8_t bp = (int8_t)std::max((int)a->ignore_end, ignore_bp);  // (byte) cast, IRPL:520
            a->set_ignore_end(bp);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// SingleSampleVariantPileupListener (discovery mode; ploidy < 3)
// ---------------------------------------------------------------------------------------------
bool SingleSampleVariantPileupListener::reference_allele_discovery(PileupRecord& p, const ReferenceGenome& g, bool call_embedded, bool ignore_lower, std::string& out) {  // SSVPL:191-206
    if (!call_embedded && p.embedded) return false;
    int last = p.pos + p.span - 1;
    std::string seq;
    if (!g.get(p.seq, p.pos, last, seq)) return false;
    if (seq.empty()) throw std::out_of_range("empty reference allele");
    if (ignore_lower && seq[0] >= 'a' && seq[0] <= 'z') return false;
    out = to_upper(seq);
    if (p.embedded) { out = out.substr(0, 1); p.str = false; }
    return true;
}

void SingleSampleVariantPileupListener::on_pileup(PileupRecord& p) {  // SSVPL:146-177
    if (n_input_variants > 0) {   // -knownVariants: the variants that start here are genotyped, nothing is discovered (SSVPL:162-176)
        if (!seq_input_variants || next_siv_index >= seq_input_variants->size()) return;
        const GenomicVariant* iv = &(*seq_input_variants)[next_siv_index];
        while (iv->first <= p.pos) {
            if (iv->first == p.pos) called.push_back(genotype_variant_sample(*iv, p, nullptr, het_rate));
            next_siv_index++;
            if (next_siv_index >= seq_input_variants->size()) return;
            iv = &(*seq_input_variants)[next_siv_index];
        }
        return;
    }
    if (p.input_str() && p.pos >= last_indel_end) last_indel_end = p.pos + p.span - 1;
    else if (p.pos <= last_indel_end) p.embedded = true;
    std::string ref;
    if (!reference_allele_discovery(p, *genome, call_embedded, ignore_lower, ref)) return;
    CalledVariant cv;
    if (discover_variant(p, ref, cv)) {
        if (!cv.var.is_snv() && !cv.undecided() && !cv.hom_ref()) last_indel_end = cv.last();
        called.push_back(std::move(cv));
    }
}

bool SingleSampleVariantPileupListener::discover_variant(PileupRecord& p, const std::string& ref_allele, CalledVariant& out) {  // SSVPL:213-232
    bool have;
    if (ref_allele.size() > 1) have = discover_with_span(p, ref_allele, out);
    else have = discover_snv_at(p, ref_allele[0], out);
    if (have && (out.undecided() || out.hom_ref() || min_quality > out.gq)) have = false;
    if (have) {
        out.sample_id = sample.id;
        out.update_cn_from_counts(sample.ploidy);
        if (out.var.is_snv() && p.embedded) out.var.type = TYPE_EMBEDDED_SNV;
    }
    return have;
}

bool SingleSampleVariantPileupListener::discover_snv_at(PileupRecord& p, char ref, CalledVariant& out) {  // SSVPL:234-239
    auto calls = p.allele_calls(1, (const std::string*)nullptr);
    CountsHelper h = CountsHelper::snv(calls, max_base_qs, 0.5);
    if (counts_dump) counts_dump->push_back({p.pos, h.counts[0], h.counts[1], h.counts[2], h.counts[3], h.total});
    if (sample.ploidy < 3) return discover_snv(h, p.seq, p.pos, ref, het_rate, csb, out);
    // pools / polyploids, SSVPL:240-254
    GenomicVariant v;
    if (!create_snv_variant_pool(p, h, ref, 0.5 / (double)sample.ploidy, v)) return false;
    CalledVariant call = genotype_variant_pool(v, sample.ploidy, calls, het_rate);
    if (v.is_biallelic()) { out = call; return true; }
    if (call.undecided() || call.hom_ref()) return false;
    auto idx = call.called_indexes();
    if (idx.size() > 2) { out = call; return true; }
    if (idx.size() == 2 && idx[0] != 0) { out = call; return true; }
    std::vector<std::string> new_alleles = {v.alleles[0]};   // LinkedHashSet: the reference, then the called alleles
    for (int8_t i : idx) if (std::find(new_alleles.begin(), new_alleles.end(), v.alleles[(size_t)i]) == new_alleles.end()) new_alleles.push_back(v.alleles[(size_t)i]);
    GenomicVariant nv = make_new_variant(v, new_alleles);
    out = genotype_variant_pool(nv, sample.ploidy, calls, het_rate);
    return true;
}

bool SingleSampleVariantPileupListener::discover_indel(PileupRecord& p, const std::vector<std::string>& alleles, const std::vector<PileupAlleleCall>& calls, CalledVariant& out) {  // SSVPL:275-296
    CountsHelper h = CountsHelper::indel(alleles, calls, max_base_qs, 0.5);
    if (sample.ploidy < 3) return call_indel(p, h, nullptr, het_rate, csb, out);
    if (h.alleles.size() == 1 || h.total == 0) return false;                      // createIndelVariantPool, SSVPL:333-338
    GenomicVariant indel;
    indel.seq = p.seq; indel.first = p.pos; indel.alleles = h.alleles; indel.len = (int)h.alleles[0].size(); indel.last = p.pos + indel.len - 1;
    if (alleles_same_length(indel.alleles)) return false;
    CalledVariant call = genotype_variant_pool(indel, sample.ploidy, calls, het_rate);
    if (indel.is_biallelic()) { out = call; return true; }
    if (call.undecided() || call.hom_ref()) return false;
    auto idx = call.called_indexes();
    if (idx.size() > 2) { out = call; return true; }
    if (idx.size() == 2 && idx[0] != 0) { out = call; return true; }
    std::vector<std::string> new_alleles = {indel.alleles[0]};
    for (int8_t i : idx) if (std::find(new_alleles.begin(), new_alleles.end(), indel.alleles[(size_t)i]) == new_alleles.end()) new_alleles.push_back(indel.alleles[(size_t)i]);
    GenomicVariant nv = make_new_variant(indel, new_alleles);
    if (alleles_same_length(nv.alleles)) return false;
    out = genotype_variant_pool(nv, sample.ploidy, calls, het_rate);
    return true;
}

// Simple algorithm based on pool thresholds to genotype a variant in a pool, SSVPL:402-503
CalledVariant SingleSampleVariantPileupListener::genotype_variant_pool(const GenomicVariant& v, short haplotypes, const std::vector<PileupAlleleCall>& calls, double h) {
    double step = 1.0 / (double)haplotypes;
    const std::vector<std::string>& alleles = v.alleles;
    std::vector<double> freqs;
    std::vector<CountsHelper> helpers;
    for (double freq = step; freq < 0.51; freq += step) {
        freqs.push_back(freq);
        if (v.is_snv()) helpers.push_back(CountsHelper::snv_gt(alleles, calls, max_base_qs, freq));
        else helpers.push_back(CountsHelper::indel(alleles, calls, max_base_qs, freq));
    }
    const CountsHelper* helper = &helpers[0];
    const std::vector<int>& counts = helper->counts;
    int major = -1;
    for (size_t i = 0; i < counts.size(); i++) if (major == -1 || counts[(size_t)major] < counts[i]) major = (int)i;   // NumberArrays.getIndexMaximum
    if (counts[(size_t)major] < haplotypes) {
        CalledVariant u = CalledVariant::make_impl(v, {});
        u.update_cn_from_counts(haplotypes);
        return u;
    }
    double lp_het = std::log10(h), lp_hom = std::log10(1 - h);
    std::vector<double> terms(helpers.size() + 1);
    double term_hom = helper->logc[(size_t)major][(size_t)major] + lp_hom;
    double max_het_post = 0, min_hom_post = 1, max_freq = 0;
    int max_freq_idx = 0, max_alt = -1;
    for (int i = 0; i < (int)alleles.size(); i++) {
        if (i == major) continue;
        terms[0] = term_hom;
        for (size_t j = 0; j < freqs.size(); j++) terms[j + 1] = helpers[j].logc[(size_t)major][(size_t)i] + lp_het;
        CountsHelper::posterior_from_events(terms);
        int idx_max = -1;
        for (size_t k = 0; k < terms.size(); k++) if (idx_max == -1 || terms[(size_t)idx_max] < terms[k]) idx_max = (int)k;
        if (idx_max == 0) min_hom_post = std::min(min_hom_post, terms[0]);
        else if (max_alt == -1 || max_het_post < terms[(size_t)idx_max]) {
            max_het_post = terms[(size_t)idx_max];
            max_freq_idx = idx_max - 1;
            max_freq = freqs[(size_t)max_freq_idx];
            max_alt = i;
        }
    }
    std::vector<int8_t> called = {(int8_t)major};
    if (max_alt != -1) called.push_back((int8_t)max_alt);
    std::sort(called.begin(), called.end());
    CalledVariant call = CalledVariant::make_impl(v, called);
    call.depth = helper->total;
    std::vector<short> acn(alleles.size(), 0);
    if (max_alt == -1) {
        call.gq = phred_score(1 - min_hom_post);
        acn[(size_t)major] = haplotypes;
    } else {
        helper = &helpers[(size_t)max_freq_idx];
        auto post = helper->posteriors_major(h, major);
        call.gq = phred_score(1 - post[(size_t)max_alt]);
        short alt_cn = (short)java_round(max_freq * haplotypes);
        if (alt_cn == 0) alt_cn++;
        else if (alt_cn == haplotypes) alt_cn--;
        acn[(size_t)max_alt] = alt_cn;
        acn[(size_t)major] = (short)(haplotypes - alt_cn);
    }
    call.alleles_cn = acn;                                   // setAllelesCopyNumber: also the total copy number (Impl:204-226)
    { int t = 0; for (short c : acn) t += c; call.total_cn = (short)t; }
    call.has_report = true; call.rep_alleles = alleles; call.rep_counts = counts; call.rep_has_counts = true;   // counts of the FIRST helper (SSVPL:497)
    call.rep_logc = helper->logc; call.rep_has_logc = true;
    return call;
}

bool SingleSampleVariantPileupListener::discover_with_span(PileupRecord& p, const std::string& ref_allele, CalledVariant& out) {  // SSVPL:257-273
    auto calls = p.allele_calls((int)ref_allele.size(), (const std::string*)nullptr);
    auto alleles = cluster_allele_calls(calls, ref_allele, max_base_qs);
    bool have = discover_indel(p, alleles, calls, out);
    if (have && (out.undecided() || out.hom_ref() || min_quality > out.gq)) have = false;
    if (!p.input_str() && !have) {
        if (p.new_str) { p.str = false; p.new_str = false; }
        have = discover_snv_at(p, ref_allele[0], out);
    }
    return have;
}

CalledVariant SingleSampleVariantPileupListener::genotype_variant_sample(const GenomicVariant& v, PileupRecord& p, const Sample* s, double h) {  // SSVPL:361-391
    const std::string& ref_allele = v.alleles[0];
    short ploidy = s ? s->ploidy : sample.ploidy;
    auto calls = p.allele_calls((int)ref_allele.size(), s ? &s->read_groups : nullptr);
    CalledVariant cv;
    if (v.is_snv()) {
        CountsHelper hs = CountsHelper::snv(calls, max_base_qs, 0.5);
        if (ploidy >= 3) {   // SSVPL:367-370
            cv = genotype_variant_pool(v, ploidy, calls, h);
            cv.has_all_counts = true; cv.all_counts = hs.counts;
        } else {
            cv = genotype_snv(v, hs, h, false);
            cv.update_cn_from_counts(ploidy);
        }
    } else if (ploidy >= 3) {   // SSVPL:376-377
        cv = genotype_variant_pool(v, ploidy, calls, h);
    } else {
        CountsHelper hi = CountsHelper::indel(v.alleles, calls, max_base_qs, 0.5);
        call_indel(p, hi, &v, h, false, cv);
        cv.update_cn_from_counts(ploidy);
    }
    if (min_quality > cv.gq) cv.make_undecided();
    cv.sample_id = s ? s->id : sample.id;
    return cv;
}

// ---------------------------------------------------------------------------------------------
// AlignmentsPileupGenerator
// ---------------------------------------------------------------------------------------------
int AlignmentsPileupGenerator::filter_flags() const {  // APG:367-371
    int f = FLAG_READ_UNMAPPED;
    if (!secondary) {
        f += FLAG_SECONDARY;
        if (!non_unique_primary) f += FLAG_MULTIPLE_ALN;
    }
    return f;
}

int AlignmentsPileupGenerator::query_filter(const ReadAlignment& a) {  // APG:310-322
    if (!has_query) return 0;
    if (query_seq == a.seq_name) {
        query_found = true;
        if (a.first > query_last) return 2;
        if (query_first > a.last) return 1;
        return 0;
    } else if (query_found) return 2;
    return 1;
}

void AlignmentsPileupGenerator::process_alignment(const ReadPtr& a) {  // APG:377-403
    if (have_seq) {
        bool same = cur_seq == a->seq_name;
        if (!same || last_start != a->first) {
            same_start_alns();
            if (!same) {
                process_pileups(cur_last + 1);
                for (auto* l : listeners) l->on_sequence_end(cur_seq);
                have_seq = false;
            } else {
                process_pileups(a->first);
            }
        }
    }
    if (!have_seq) {  // startSequence APG:435-444
        cur_seq = a->seq_name;
        have_seq = true;
        cur_pos = a->first;
        cur_last = a->last;
        for (auto* l : listeners) l->on_sequence_start(cur_seq);
    }
    if (a->last > cur_last) cur_last = a->last;
    if (a->is_secondary()) same_secondary.push_back(a);
    else same_primary.push_back(a);
    last_start = a->first;
}

void AlignmentsPileupGenerator::same_start_alns() {  // APG:407-433
    int pos_start = 0;
    if (!same_primary.empty()) pos_start = same_primary[0]->first;
    else if (!same_secondary.empty()) pos_start = same_secondary[0]->first;
    if (pos_start == 0) return;
    std::map<std::string, int> per_rg;
    std::vector<ReadPtr> all(same_primary);
    all.insert(all.end(), same_secondary.begin(), same_secondary.end());
    same_primary.clear();
    same_secondary.clear();
    for (auto& a : all) {
        auto it = per_rg.find(a->read_group);
        if (it == per_rg.end()) per_rg[a->read_group] = 1;
        else if (max_alns_per_start <= 0 || it->second < max_alns_per_start) it->second++;
        else continue;
        a->set_ignore_5p(ignore5);
        a->set_ignore_3p(ignore3);
        pending.push_back(a);
    }
}

void AlignmentsPileupGenerator::notify_end() {  // APG:447-452
    same_start_alns();
    process_pileups(std::min(query_last, cur_last) + 1);
    if (have_seq) for (auto* l : listeners) l->on_sequence_end(cur_seq);
    have_seq = false;
}

void AlignmentsPileupGenerator::process_pileups(int start) {  // APG:453-462
    if (start == cur_pos) return;
    while (cur_pos < start) {
        if (!process_current()) {
            update_pending();
            if (pending.empty()) cur_pos = start;
        }
    }
    update_pending();
}

void AlignmentsPileupGenerator::update_pending() {  // APG:464-471
    std::vector<ReadPtr> keep;
    for (auto& a : pending) if (a->last >= cur_pos) keep.push_back(a);
    pending.swap(keep);
}

bool AlignmentsPileupGenerator::process_current() {  // APG:475-498
    if (pending.empty()) { cur_pos++; return false; }
    if (has_query && (cur_pos < query_first || cur_pos > query_last)) { cur_pos++; return false; }
    PileupRecord pileup;
    pileup.seq = cur_seq;
    pileup.pos = cur_pos;
    for (auto& a : pending) pileup.add(a.get());
    n_pileups++;
    if (pileup.n_alns > 0) n_nonempty++;
    for (auto* l : listeners) l->on_pileup(pileup);
    cur_pos++;
    return pileup.n_alns > 0;
}

// ---------------------------------------------------------------------------------------------
// VCF text
// ---------------------------------------------------------------------------------------------
void print_vcf_header(std::ostream& out, const std::vector<Sample>& samples, bool psp) {  // VCFFileHeader.java:45-71, 219-246
    out << "##fileformat=VCFv4.2\n";
    struct L { const char *t, *id, *num, *type, *desc; };
    static const L lines[] = {
        {"INFO", "CNV", "1", "Integer", "Number of samples with CNVs around this variant"},
        {"INFO", "TA", "1", "String", "Variant annotation based on a gene model"},
        {"INFO", "TID", "1", "String", "Id of the transcript related to the variant annotation"},
        {"INFO", "TGN", "1", "String", "Name of the gene related to the variant annotation"},
        {"INFO", "TCO", "1", "Float", "One based codon position of the start of the variant. The decimal is the codon position"},
        {"INFO", "TACH", "1", "String", "Description of the aminoacid change produced by a non-synonymous mutation. String encoded as reference aminoacid, position and mutated aminoacid"},
        {"INFO", "NS", "1", "Integer", "Number of samples genotyped"},
        {"INFO", "MAF", "1", "Float", "Minor allele frequency"},
        {"INFO", "OH", "1", "Float", "Observed heterozygosity"},
        {"INFO", "AN", "1", "Integer", "Number of alleles in called genotypes"},
        {"INFO", "AFS", "R", "Integer", "Allele counts over the population for all alleles, including the reference"},
        {"INFO", "TYPE", "1", "String", "Type of variant"},
        {"INFO", "FS", "1", "Float", "Phred-scaled p-value using Fisher's exact test to detect strand bias"},
        {"INFO", "END", "1", "Integer", "End position of the structural variant"},
        {"INFO", "SVTYPE", "1", "String", "Type of SV:DEL=Deletion, INS=Insertion, DUP=Duplication, INV=Inversion"},
        {"INFO", "SVLEN", "1", "Integer", "Difference in length between REF and ALT alleles"},
        {"FORMAT", "GT", "1", "String", "Genotype"},
        {"FORMAT", "PL", "G", "Integer", "Phred-scaled genotype likelihoods rounded to the closest integer"},
        {"FORMAT", "GQ", "1", "Integer", "Genotype quality"},
        {"FORMAT", "DP", "1", "Integer", "Read depth"},
        {"FORMAT", "ADP", "R", "Integer", "Counts for observed alleles, including the reference allele"},
        {"FORMAT", "BSDP", "4", "Integer", "Number of base calls (depth) for the 4 nucleotides in called SNVs sorted as A,C,G,T"},
        {"FORMAT", "ACN", "R", "Integer", "Predicted copy number of each allele taking into account the prediction of number of copies of the region surrounding the variant"},
    };
    for (auto& l : lines)  // attribute order ID,Number,Type,Description: VCFHeaderLine.java:44-50
        out << "##" << l.t << "=<ID=" << l.id << ",Number=" << l.num << ",Type=" << l.type << ",Description=\"" << l.desc << "\">\n";
    if (psp) {
        std::map<std::string, const Sample*> sorted;  // samplesWithHeaderLine is a TreeMap
        for (auto& s : samples) sorted[s.id] = &s;
        for (auto& kv : sorted) out << "##SAMPLE=<ID=" << kv.second->id << ",PL=" << kv.second->ploidy << ">\n";
    }
    out << "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO";
    if (!samples.empty()) out << "\tFORMAT";
    for (auto& s : samples) out << "\t" << s.id;
    out << "\n";
}

static const char* type_name(int8_t t) {  // GenomicVariantImpl.java:59-71
    switch (t) {
        case TYPE_BIALLELIC_SNV: return "SNV"; case TYPE_MULTIALLELIC_SNV: return "MULTISNV"; case TYPE_EMBEDDED_SNV: return "EMBEDDED";
        case TYPE_INDEL: return "INDEL"; case TYPE_STR: return "STR";
    }
    return nullptr;
}

void print_genotype_field(std::ostream& out, CalledVariant& call, bool snv_format, short ploidy) {  // VCFFileWriter.java:159-308
    out << "\t";
    const auto& alleles = call.var.alleles;
    auto idx = call.called_indexes();
    // GT
    if (idx.empty()) { out << "."; if (ploidy > 1) out << "/."; }
    else if (idx.size() == 1) { out << (int)idx[0]; if (ploidy > 1) out << "/" << (int)idx[0]; }
    else for (size_t i = 0; i < idx.size(); i++) { if (i > 0) out << "/"; out << (int)idx[i]; }
    // PL
    out << ":";
    for (size_t j = 0; j < alleles.size(); j++) for (size_t i = 0; i <= j; i++) {
        if (i > 0 || j > 0) out << ",";
        int cond = 0;
        bool present = call.called_snv ? ((float)(call.homo_ref + call.ref_alt + call.alt_ref + call.homo_alt) != 0.0f) : (call.has_report && call.rep_has_logc);
        if (present) cond = (int)java_round(-10 * call.report_logc(alleles[i], alleles[j]));
        out << cond;
    }
    out << ":" << call.gq;          // GQ
    out << ":" << call.depth;       // DP
    out << ":";
    if (snv_format) {               // BSDP
        if (call.called_snv) out << call.cnt[0] << "," << call.cnt[1] << "," << call.cnt[2] << "," << call.cnt[3];
        else if (!call.has_all_counts) out << "0,0,0,0";
        else for (size_t i = 0; i < call.all_counts.size(); i++) { if (i) out << ","; out << call.all_counts[i]; }
    } else {                        // ADP
        bool present = call.called_snv ? (call.cnt[0] + call.cnt[1] + call.cnt[2] + call.cnt[3] > 0) : (call.has_report && call.rep_has_counts);
        for (size_t i = 0; i < alleles.size(); i++) { if (i) out << ","; out << (present ? call.report_count(alleles[i]) : 0); }
    }
    out << ":";                     // ACN
    short total = call.total_cn;
    if (total == 0) out << ".";
    else {
        std::vector<short> cn;
        if (call.called_snv) { cn = {0, 0}; if (!call.undecided()) { cn[0] = call.ref_cn; cn[1] = (short)(call.total_cn - call.ref_cn); } }
        else cn = call.alleles_cn;
        if (call.undecided()) cn[0] = total;
        for (size_t j = 0; j < cn.size(); j++) { if (j) out << ","; out << cn[j]; }
    }
}

void print_vcf_record_single(std::ostream& out, CalledVariant& call, short ploidy) {  // SSVD:944-963 + VCFFileWriter.java:44-104
    bool snv_format = call.called_snv || call.has_all_counts;
    const GenomicVariant& v = call.var;
    out << v.seq << "\t" << v.first << "\t" << (v.id.empty() ? "." : v.id) << "\t" << v.alleles[0] << "\t";
    if (v.alleles.size() == 1) out << ".";
    else for (size_t i = 1; i < v.alleles.size(); i++) { if (i > 1) out << ","; out << v.alleles[i]; }
    out << "\t" << v.qs << "\t.\t";
    bool printed = false;
    if (call.called_snv && call.strand_bias != INVALID_STRAND_BIAS_SCORE) { out << "FS=" << (int)call.strand_bias; printed = true; }
    const char* tn = type_name(v.type);
    if (v.type != TYPE_UNDETERMINED && v.type != TYPE_BIALLELIC_SNV && tn) { if (printed) out << ";"; out << "TYPE=" << tn; printed = true; }
    if (!printed) out << ".";
    out << "\t" << (snv_format ? "GT:PL:GQ:DP:BSDP:ACN" : "GT:PL:GQ:DP:ADP:ACN");
    print_genotype_field(out, call, snv_format, ploidy);
    out << "\n";
}

// ---------------------------------------------------------------------------------------------
// SingleSampleVariantsDetector (findSNVs with default repeats/RD/RP switches off)
// ---------------------------------------------------------------------------------------------
void SingleSampleVariantsDetector::begin() {
    if (!opt.het_modified && opt.ploidy == 1) opt.het_rate = 0.000001;  // SSVD:591-593
    gen.query_seq = opt.query_seq; gen.has_query = opt.has_query; gen.query_first = opt.query_first; gen.query_last = opt.query_last;
    gen.max_alns_per_start = opt.max_alns_per_start; gen.non_unique_primary = opt.non_unique_primary; gen.secondary = opt.secondary;
    gen.ignore5 = (int8_t)opt.ignore5; gen.ignore3 = (int8_t)opt.ignore3; gen.min_mq = opt.min_mq;
    realigner.genome = genome;
    if (!opt.known_variants.empty()) {   // SSVD:897-905: the same collection goes to the realigner and to the variant listener
        auto by_seq = sort_variants_by_sequence(load_variants(opt.known_variants));
        for (auto& kv : by_seq) {
            var_listener.n_input_variants += kv.second.size();
            for (auto& v : kv.second) realigner.input_variants[kv.first].push_back({v.first, v.last, v.type == TYPE_STR});
        }
        var_listener.input_variants = std::move(by_seq);
    } else if (!opt.known_strs.empty()) realigner.input_variants = make_non_redundant_strs(*genome, load_regions(opt.known_strs));   // SSVD:906-913
    var_listener.genome = genome; var_listener.het_rate = opt.het_rate; var_listener.csb = opt.csb;
    var_listener.max_base_qs = (int8_t)opt.max_base_qs; var_listener.ignore_lower = opt.ignore_lower;
    var_listener.call_embedded = opt.call_embedded; var_listener.min_quality = opt.min_quality;
    var_listener.sample = Sample{opt.sample_id, opt.ploidy, {}};
    gen.listeners = {&realigner, &var_listener, this};  // SSVD:919-925
    reader.filter_flags = gen.filter_flags();
    reader.min_mq = gen.min_mq;
    reader.genome = genome;
    if (out) print_vcf_header(*out, {var_listener.sample}, opt.psp);
}

void SingleSampleVariantsDetector::add_record(const SamRecord& rec) {
    if (stopped) return;
    ReadPtr a = reader.load(rec);
    if (!a) return;
    int q = gen.query_filter(*a);
    if (q == 2) { stopped = true; return; }
    if (q == 1) return;
    gen.process_alignment(a);
}

void SingleSampleVariantsDetector::finish() { gen.notify_end(); }

void SingleSampleVariantsDetector::on_sequence_end(const std::string&) {  // SSVD:933-968
    auto& calls = var_listener.called;
    std::stable_sort(calls.begin(), calls.end(), [](const CalledVariant& a, const CalledVariant& b) {
        if (a.var.first != b.var.first) return a.var.first < b.var.first;
        return a.var.last < b.var.last;
    });
    for (auto& c : calls) {
        c.update_cn_from_counts(opt.ploidy);  // intersectVariantsCNVs SSVD:975
        if (out) print_vcf_record_single(*out, c, opt.ploidy);
    }
    if (out) out->flush();
    calls.clear();
}

}  // namespace ngsep_oracle


// ---- -knownSTRs: genome/io/SimpleGenomicRegionFileHandler.java:57-80 and SingleSampleVariantsDetector.makeNonRedundantSTRs (SSVD:843-894) ----
namespace ngsep_oracle {
std::vector<GenomicRegionRow> load_regions(const std::string& path) {
    std::vector<GenomicRegionRow> out;
    std::ifstream in(path);
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::vector<std::string> items;   // line.split(" |\t")
        size_t s = 0;
        while (true) {
            size_t e = line.find_first_of(" \t", s);
            items.push_back(line.substr(s, e == std::string::npos ? std::string::npos : e - s));
            if (e == std::string::npos) break;
            s = e + 1;
        }
        if (items.size() < 3) continue;
        char* end = nullptr;
        long f = std::strtol(items[1].c_str(), &end, 10);
        if (items[1].empty() || *end) continue;     // NumberFormatException: the line is skipped
        long l = std::strtol(items[2].c_str(), &end, 10);
        if (items[2].empty() || *end) continue;
        out.push_back({items[0], (int)f, (int)l});
    }
    return out;
}

static int overlap_length(const std::string& s1, const std::string& s2) {   // sequences/AbstractLimitedSequence.java:376-390
    for (size_t i = 0; i < s1.size(); i++) {
        size_t a = i, b = 0;
        bool ok = true;
        for (; a < s1.size() && b < s2.size(); a++, b++) if (s1[a] != s2[b]) { ok = false; break; }
        if (ok && a == s1.size()) return (int)(s1.size() - i);
    }
    return 0;
}
static std::string upper_copy(std::string s) { for (auto& c : s) if (c >= 'a' && c <= 'z') c = (char)(c - 32); return s; }

std::map<std::string, std::vector<InputVariant>> make_non_redundant_strs(const ReferenceGenome& genome, const std::vector<GenomicRegionRow>& strs) {
    std::map<std::string, std::vector<InputVariant>> answer;
    for (size_t si = 0; si < genome.names.size(); si++) {
        const std::string& name = genome.names[si];
        const int seq_len = (int)genome.seqs[si].size();
        std::vector<GenomicRegionRow> rs;
        for (auto& r : strs) if (r.seq == name) rs.push_back(r);
        std::stable_sort(rs.begin(), rs.end(), [](const GenomicRegionRow& a, const GenomicRegionRow& b) { return a.first != b.first ? a.first < b.first : a.last < b.last; });
        auto merge = [&](int first, int last, const GenomicRegionRow& r) {   // mergeSTRs, SSVD:872-879
            if (r.first - last > 5) return false;
            if (r.first - last <= 2) return true;
            std::string ref1, ref2;
            if (!genome.get(name, std::max(first, last - 10), last, ref1) || !genome.get(name, r.first, r.last, ref2)) return false;
            return overlap_length(upper_copy(ref1), upper_copy(ref2)) > 5;
        };
        auto emit = [&](int first, int last) {   // makeSTRVariant, SSVD:882-894: the variant spans its reference allele
            std::string ref;
            if (!genome.get(name, first, last, ref)) return;
            answer[name].push_back({first, first + (int)ref.size() - 1, true});
        };
        int first = 0, last = 0;
        for (auto& r : rs) {
            if (last == 0 || !merge(first, last, r)) {
                if (last > 0) emit(std::max(1, first - 1), std::min(last + 1, seq_len));
                first = r.first;
            }
            last = r.last;
        }
        if (last > 0) emit(std::max(1, first - 1), std::min(last + 1, seq_len));
    }
    return answer;
}
}  // namespace ngsep_oracle


// ---- -knownVariants: vcf/VCFFileReader.loadVariants (VCFFileReader.java:585-600) over loadVCFRecord / loadGenomicVariant /
//      loadInfoField (:152-306), load mode MINIMAL: only the variant of each record matters on this path ----
namespace ngsep_oracle {
static std::vector<std::string> parse_string(const std::string& s, char delim) {   // main/io/ParseUtils.java:38-53
    std::vector<std::string> out;
    std::string cur;
    for (char c : s) { if (c == delim) { out.push_back(cur); cur.clear(); } else cur.push_back(c); }
    out.push_back(cur);
    return out;
}
static bool java_parse_int(const std::string& s, long& v) {   // Integer.parseInt
    if (s.empty()) return false;
    size_t i = (s[0] == '-' || s[0] == '+') ? 1 : 0;
    if (i == s.size()) return false;
    for (size_t k = i; k < s.size(); k++) if (!std::isdigit((unsigned char)s[k])) return false;
    errno = 0;
    v = std::strtol(s.c_str(), nullptr, 10);
    return errno == 0 && v >= INT32_MIN && v <= INT32_MAX;
}
static int8_t variant_type_id(const std::string& name) {   // GenomicVariantImpl.java:44-58, 90-96
    static const std::map<std::string, int8_t> ids = {
        {"SNV", TYPE_BIALLELIC_SNV}, {"MULTISNV", TYPE_MULTIALLELIC_SNV}, {"EMBEDDED", TYPE_EMBEDDED_SNV}, {"INDEL", TYPE_INDEL}, {"STR", TYPE_STR},
        {"CNV", 10}, {"REPEAT", 11}, {"DEL", 12}, {"INS", 13}, {"INV", 14}, {"DUP", 15}, {"Deletion", 12}, {"Insertion", 13}, {"Inversion", 14}, {"Duplication", 15}};
    auto it = ids.find(name);
    return it == ids.end() ? TYPE_UNDETERMINED : it->second;
}

std::vector<GenomicVariant> load_variants(const std::string& path) {
    std::vector<GenomicVariant> answer;
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open " + path);
    std::string line;
    size_t n_samples = 0;
    bool have_samples_line = false;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (!have_samples_line) {   // loadHeader (:142-149): "##" lines, then the line with the sample ids (VCFFileHeader.java:155-167)
            if (line.rfind("##", 0) == 0) continue;
            if (line.empty() || line[0] != '#') throw std::runtime_error("VCF file does not have line with sample ids");
            auto items = parse_string(line, '\t');
            n_samples = items.size() > 9 ? items.size() - 9 : 0;
            have_samples_line = true;
            continue;
        }
        auto items = parse_string(line, '\t');
        if (items.size() < 8) continue;                                    // :154-157
        // loadGenomicVariant, :192-255
        long position;
        if (!java_parse_int(items[1], position) || position <= 0) continue;
        if (items[3].empty() || items[4].empty()) continue;
        std::vector<std::string> alleles = {items[3]};
        if (items[4][0] != '.') for (auto& a : parse_string(items[4], ',')) alleles.push_back(a);
        if ((int)alleles.size() > MAX_NUM_ALLELES) continue;
        short qs = 0;
        if (!items[5].empty() && items[5][0] != '.') {
            char* end = nullptr;
            double d = std::strtod(items[5].c_str(), &end);
            if (!end || *end) continue;
            if (d > 32767) d = 255;
            qs = (short)java_round(d);
            if (qs < 0) continue;
        }
        GenomicVariant v;
        v.seq = items[0]; v.first = (int)position; v.qs = qs;
        if (items[2] != ".") v.id = items[2];
        auto in_alphabet = [](char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; };
        if (alleles.size() == 2 && items[3].size() == 1 && items[4].size() == 1 && in_alphabet(items[3][0]) && in_alphabet(items[4][0])) {
            v.snv_class = true; v.type = TYPE_BIALLELIC_SNV; v.last = v.first; v.len = 1;
        } else {
            v.len = (int)alleles[0].size(); v.last = v.first + v.len - 1;      // GenomicVariantImpl.java:99-106
        }
        v.alleles = alleles;
        // loadInfoField, :261-306 (what changes the variant itself)
        if (items[7] != ".") {
            bool sv_len_present = false;
            std::vector<std::string> info;   // parseStringWithText(';', '"')
            { std::string cur; bool text = false; for (char c : items[7]) { if (c == '"') text = !text; if (c == ';' && !text) { info.push_back(cur); cur.clear(); } else cur.push_back(c); } info.push_back(cur); }
            for (auto& it : info) {
                size_t eq = it.find('=');
                if (eq == std::string::npos) continue;
                std::string attr = it.substr(0, eq), value = it.substr(eq + 1);
                if (attr == "TYPE") { int8_t t = variant_type_id(value); if (t > 0 && t <= 14) v.type = t; }
                if (attr == "END" && !v.snv_class) { long l; if (!java_parse_int(value, l)) throw std::runtime_error("NumberFormatException: " + value); v.last = (int)l; if (!sv_len_present) v.len = v.last - v.first + 1; }
                if (attr == "SVLEN" && !v.snv_class) { long l; if (!java_parse_int(value, l)) throw std::runtime_error("NumberFormatException: " + value); v.len = (int)l; sv_len_present = true; }
                if (attr == "SVTYPE") { int8_t t = variant_type_id(value); if (t >= 10) v.type = t; }
            }
        }
        // :166-181: the genotype columns must agree with the header
        if (items.size() == 8) { if (n_samples > 0) continue; }
        else if (items.size() - 9 != n_samples) continue;
        if (v.alleles.size() < 2) continue;     // filterReferenceSitesGVCF, :593
        if (v.type >= 10) continue;             // filterSVs, :594
        answer.push_back(std::move(v));
    }
    if (!have_samples_line) throw std::runtime_error("VCF file does not have line with sample ids");
    return answer;
}

std::map<std::string, std::vector<GenomicVariant>> sort_variants_by_sequence(const std::vector<GenomicVariant>& vars) {
    std::map<std::string, std::vector<GenomicVariant>> out;
    for (auto& v : vars) out[v.seq].push_back(v);
    for (auto& kv : out)
        std::stable_sort(kv.second.begin(), kv.second.end(), [](const GenomicVariant& a, const GenomicVariant& b) { return a.first != b.first ? a.first < b.first : a.last < b.last; });
    return out;
}
}  // namespace ngsep_oracle
