// TEST INFRASTRUCTURE ONLY — see ngsep_oracle.hpp.  Part 1: Java-semantics helpers, reference genome,
// ReadAlignment (alignments/ReadAlignment.java) and the htsjdk->ReadAlignment conversion
// (alignments/io/ReadAlignmentFileReader.java).
#include "ngsep_oracle.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>

namespace ngsep_oracle {

// ---------------------------------------------------------------------------------------------
// Java semantics
// ---------------------------------------------------------------------------------------------
long java_round(double x) {
    // java.lang.Math.round(double): floor(x + 1/2) evaluated without the double-rounding of x+0.5
    if (std::isnan(x)) return 0;
    double f = std::floor(x);
    double diff = x - f;  // exact
    return (long)(diff >= 0.5 ? f + 1.0 : f);
}

short phred_score(double p) {  // math/PhredScoreHelper.java:31-40
    if (p == 0) return 255;
    double score = -10 * std::log10(p);
    if (score > 255) return 255;
    return (short)java_round(score);
}

double phred_probability(short q) {  // math/PhredScoreHelper.java:46-51
    if (q >= 255) return 0;
    return std::pow(10.0, -0.1 * q);
}

double log_sum10(double l1, double l2) {  // math/LogMath.java:38-44 (non-null arguments)
    if (l1 - l2 > 20) return l1;
    if (l2 - l1 > 20) return l2;
    return l1 + std::log10(1 + std::pow(10.0, l2 - l1));
}

static std::vector<double> g_log_fact;
static double fisher_exact_value(int a, int b, int c, int d) {  // math/FisherExactTest.java:120-136
    int n = a + b + c + d;
    if ((int)g_log_fact.size() <= n) {  // FisherExactTest.java:100-108
        int m = std::max(n, 10000);
        g_log_fact.assign(m + 1, 0.0);
        for (int i = 2; i <= m; i++) g_log_fact[i] = g_log_fact[i - 1] + std::log10((double)i);
    }
    double ans = g_log_fact[a + b];
    ans += g_log_fact[c + d];
    ans += g_log_fact[a + c];
    ans += g_log_fact[b + d];
    ans -= g_log_fact[a];
    ans -= g_log_fact[b];
    ans -= g_log_fact[c];
    ans -= g_log_fact[d];
    ans -= g_log_fact[n];
    return std::pow(10.0, ans);
}

double fisher_pvalue(int a, int b, int c, int d) {  // math/FisherExactTest.java:65-99 (quick = true, :31)
    if (a > b) { std::swap(a, b); std::swap(c, d); }
    if (a > c) { std::swap(a, c); std::swap(b, d); }
    int e = std::min(a, d);
    double answer = 0;
    while (a >= 0 && d >= 0) {
        double p = fisher_exact_value(a, b, c, d);
        if (e >= 10 && answer > 100.0 * e * p) break;
        answer += p;
        a--; b++; c++; d--; e++;
    }
    return answer;
}

// math/CountsRankHelper.java:36-63 for Character keys + sequences/HammingSequenceDistanceMeasure.java:88-101
std::string hamming_consensus(const std::vector<std::string>& seqs) {
    size_t l = seqs[0].size();
    std::string out;
    for (size_t i = 0; i < l; i++) {
        std::map<char, int> counts;  // TreeMap<Character,...>
        for (const auto& s : seqs) counts[s[i]]++;
        // stable sort by count descending over TreeMap value order => first maximum in key order wins
        char best = 0; int bc = -1;
        for (auto& kv : counts) if (kv.second > bc) { bc = kv.second; best = kv.first; }
        out.push_back(best);
    }
    return out;
}

std::vector<int> java_int_hashmap_order(const std::vector<int>& keys) {
    // java.util.HashMap: table 16, load factor .75, doubles when size exceeds threshold; bucket =
    // (h ^ h>>>16) & (cap-1) with h = Integer.hashCode = value; bins keep insertion order (resize splits
    // preserve relative order).  Iteration = bucket order, then bin order.
    size_t cap = 16;
    while (keys.size() > cap * 3 / 4) cap *= 2;
    std::vector<std::pair<unsigned, size_t>> tagged;
    for (size_t i = 0; i < keys.size(); i++) {
        unsigned h = (unsigned)keys[i];
        h ^= (h >> 16);
        tagged.push_back({h & (unsigned)(cap - 1), i});
    }
    std::stable_sort(tagged.begin(), tagged.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
    std::vector<int> out;
    for (auto& t : tagged) out.push_back(keys[t.second]);
    return out;
}

// ---------------------------------------------------------------------------------------------
// Reference genome
// ---------------------------------------------------------------------------------------------
void ReferenceGenome::add(const std::string& name, const std::string& chars) {
    std::string s = chars;
    static const char* ALPHA = "AaCcNngGtT";  // sequences/DNAMaskedSequence.java:33
    for (auto& c : s) if (!std::strchr(ALPHA, c) || c == 0) c = 'N';  // default index 4 = 'N' (:57-59)
    index[name] = (int)names.size();
    names.push_back(name);
    seqs.push_back(std::move(s));
}

ReferenceGenome ReferenceGenome::load_fasta(const std::string& path) {
    ReferenceGenome g;
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open fasta " + path);
    std::string line, name, seq;
    bool have = false;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (!line.empty() && line[0] == '>') {
            if (have) g.add(name, seq);
            // sequences/io/FastaFileReader.java:176-185: id up to first space or tab
            size_t e = line.find_first_of(" \t");
            name = line.substr(1, e == std::string::npos ? std::string::npos : e - 1);
            seq.clear();
            have = true;
        } else if (have) {
            seq += line;
        }
    }
    if (have) g.add(name, seq);
    return g;
}

int ReferenceGenome::find(const std::string& name) const {
    auto it = index.find(name);
    return it == index.end() ? -1 : it->second;
}

bool ReferenceGenome::get(const std::string& name, int first, int last, std::string& out) const {
    int i = find(name);
    if (i < 0) return false;
    const std::string& s = seqs[i];
    if (first >= 1 && last <= (int)s.size()) {  // ReferenceGenome.java:238-243
        if (last < first - 1) throw std::runtime_error("negative substring");
        out = s.substr(first - 1, last - first + 1);
        return true;
    }
    return false;
}

long ReferenceGenome::total_length() const {
    long t = 0;
    for (auto& s : seqs) t += (long)s.size();
    return t;
}

bool GenomicVariant::is_snv() const {
    if (snv_class) return true;
    if (alleles.size() < 2) return false;
    for (auto& a : alleles) {
        if (a.size() != 1) return false;
        if (!std::strchr("ACGT", a[0]) || a[0] == 0) return false;  // DNASequence.isInAlphabeth
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// ReadAlignment
// ---------------------------------------------------------------------------------------------
static inline int op_len(int v) { return v / 8; }
static inline int op_code(int v) { return v & 7; }
static inline bool c_read(int v) { return (v & 2) != 0; }
static inline bool c_ref(int v) { return (v & 1) != 0; }
static inline bool is_indel(int op) { return op == OP_D || op == OP_I; }
static inline int aln_value(int len, int op) { return 8 * len + op; }

void ReadAlignment::set_cigar_string(const std::string& cigar) {  // RA:1222-1266
    static const std::string CODES = "HDIMPNSX";
    std::vector<int> list;
    size_t next_start = 0;
    int exp_read = 0, exp_end = first;
    for (size_t i = 0; i < cigar.size(); i++) {
        char c = cigar[i];
        if (!std::isdigit((unsigned char)c)) {
            std::string ls = cigar.substr(next_start, i - next_start);
            if (ls.empty()) throw std::invalid_argument("Malformed CIGAR:" + cigar);
            int length = std::atoi(ls.c_str());
            int op = (int)CODES.find(c);
            if (CODES.find(c) == std::string::npos) op = (c == '=') ? OP_M : -1;
            if (op < 0) throw std::invalid_argument("Malformed CIGAR operator");  // Java would later index out of range
            list.push_back(aln_value(length, op));
            if (c_read(op)) exp_read += length;
            if (c_ref(op)) exp_end += length;
            next_start = i + 1;
        }
    }
    // collapseEqualEvents RA:1296-1312
    aln.clear();
    int last_op = -1, total = 0;
    for (int v : list) {
        int l = op_len(v), op = op_code(v);
        if (op != last_op) {
            if (total > 0) aln.push_back(aln_value(total, last_op));
            total = 0;
            last_op = op;
        }
        total += l;
    }
    if (total > 0) aln.push_back(aln_value(total, last_op));
    if (exp_read != read_length) read_length = exp_read;   // RA:1256-1259
    if (exp_end - 1 != last) last = exp_end - 1;           // RA:1260-1263
    calls_updated = false;
}

void ReadAlignment::set_quality_scores(const std::string& qs) {  // RA:581-595
    quals.assign(read_length, (uint8_t)38);
    for (size_t i = 0; i < qs.size() && (int)i < read_length; i++) {
        int sig = (unsigned char)qs[i];
        if (sig > 127) sig = 127;
        quals[i] = (uint8_t)sig;
    }
    has_quals = true;
}

void ReadAlignment::set_ignore_5p(short v) {  // RA:613-626
    if (is_negative()) { if (ignore_end != v) { ignore_end = v; calls_updated = false; } }
    else if (ignore_start != v) { ignore_start = v; calls_updated = false; }
}
void ReadAlignment::set_ignore_3p(short v) {  // RA:632-644
    if (is_negative()) { if (ignore_start != v) { ignore_start = v; calls_updated = false; } }
    else if (ignore_end != v) { ignore_end = v; calls_updated = false; }
}
void ReadAlignment::set_ignore_start(short v) { if (ignore_start != v) { ignore_start = v; calls_updated = false; } }
void ReadAlignment::set_ignore_end(short v) { if (ignore_end != v) { ignore_end = v; calls_updated = false; } }

void ReadAlignment::update_allele_calls_info() {  // RA:747-834
    if (calls_updated) return;
    int ref_pos = first, read_pos = 0;
    allele_call_length.assign(read_length, 0);
    indel_calls.clear();
    has_indel_calls = false;
    bool prev_indel = false;
    int n = (int)aln.size();
    for (int i = 0; i < n; i++) {
        int length = op_len(aln[i]);
        int op = op_code(aln[i]);
        bool cr = c_ref(aln[i]), cd = c_read(aln[i]);
        int next_op = -1, next_len = 0, next_read_cons = 0;
        bool next_indel = false;
        if (i < n - 1) {
            next_op = op_code(aln[i + 1]);
            next_len = op_len(aln[i + 1]);
            next_indel = is_indel(next_op);
            next_read_cons = c_read(next_op) ? next_len : 0;
        }
        if (cr) {
            if (cd) {
                for (int j = 0; j < length; j++) {
                    bool skip = read_pos < ignore_start;
                    skip = skip || (read_length - read_pos) <= ignore_end;
                    skip = skip || (prev_indel && j < ignore_close_to_indel);
                    skip = skip || (next_indel && j < length - 1 && j >= length - ignore_close_to_indel);
                    skip = skip || (next_indel && j == length - 1 && (read_pos < ignore_close_to_indel || read_length - read_pos - next_read_cons < ignore_close_to_indel));
                    int read_pos_after = read_pos + next_read_cons + 1;
                    skip = skip || (next_indel && j == length - 1 && (read_length - read_pos_after < ignore_end));
                    if (!skip) {
                        // Java would throw ArrayIndexOutOfBounds on inconsistent CIGARs; RAFR drops such records (RAFR:341-346)
                        if (read_pos >= read_length) throw std::out_of_range("allele call beyond read length");
                        if (j == length - 1 && next_indel) {
                            int ref_last = ref_pos + 1;
                            if (next_op == OP_I) {
                                allele_call_length[read_pos] = (short)(next_len + 2);
                            } else {
                                allele_call_length[read_pos] = 2;
                                ref_last += next_len;
                            }
                            has_indel_calls = true;
                            indel_calls[ref_pos] = IndelCall{ref_pos, ref_last, next_len};
                        } else {
                            allele_call_length[read_pos] = 1;
                        }
                    }
                    ref_pos++;
                    read_pos++;
                }
            } else {
                ref_pos += length;
            }
        } else if (cd) {
            read_pos += length;
        }
        prev_indel = is_indel(op);
    }
    calls_updated = true;
}

int ReadAlignment::aligned_read_position(int ref_pos) const {  // RA:842-871
    int cur_ref = first, cur_read = 0;
    if (ref_pos < first || ref_pos > last) return -1;
    for (int v : aln) {
        int length = op_len(v);
        bool cr = c_ref(v), cd = c_read(v);
        if (cr && cd) {
            if (ref_pos < cur_ref) return -1;
            else if (cur_ref + length > ref_pos) {
                int ans = cur_read + ref_pos - cur_ref;
                if (ans < 0 || ans >= read_length) return -1;
                return ans;
            }
        }
        if (cr) cur_ref += length;
        if (cd) cur_read += length;
    }
    return -1;
}

bool ReadAlignment::allele_call(int ref_pos, std::string& out) {  // RA:989-1000
    if (!has_bases) return false;
    int rp = aligned_read_position(ref_pos);
    if (rp < 0) return false;
    update_allele_calls_info();
    int length = allele_call_length[rp];
    if (length == 0) return false;
    if (rp + length > (int)bases.size()) throw std::out_of_range("allele call subsequence");
    out = bases.substr(rp, length);
    return true;
}

bool ReadAlignment::allele_call(int ref_first, int ref_last, std::string& out) {  // RA:1008-1016
    if (!has_bases) return false;
    update_allele_calls_info();
    int rf = aligned_read_position(ref_first);
    int rl = aligned_read_position(ref_last);
    if (rf < 0 || rl < 0 || rl < rf) return false;
    if (rf < ignore_start || read_length - rl <= ignore_end) return false;  // withinIgnoreRegions RA:1042-1044
    out = bases.substr(rf, rl - rf + 1);
    return true;
}

char ReadAlignment::base_quality(int ref_pos) const {  // RA:1022-1027
    int rp = aligned_read_position(ref_pos);
    if (rp < 0) return 33;
    if (!has_quals) return '+';
    return (char)quals[rp];
}

bool ReadAlignment::base_qualities(int ref_first, int ref_last, std::string& out) const {  // RA:1034-1041
    int rf = aligned_read_position(ref_first);
    int rl = aligned_read_position(ref_last);
    if (rf < 0 || rl < 0 || rl < rf) return false;
    if (rf < ignore_start || read_length - rl <= ignore_end) return false;
    if (!has_quals) { out.assign(rl - rf + 1, '+'); return true; }
    out.clear();
    for (int i = rf; i <= rl; i++) out.push_back((char)quals[i]);
    return true;
}

const std::map<int, IndelCall>* ReadAlignment::get_indel_calls() {  // RA:1050-1054
    update_allele_calls_info();
    return has_indel_calls ? &indel_calls : nullptr;
}

const IndelCall* ReadAlignment::get_indel_call(int ref_pos) {  // RA:1101-1106
    update_allele_calls_info();
    if (!has_indel_calls) return nullptr;
    auto it = indel_calls.find(ref_pos);
    return it == indel_calls.end() ? nullptr : &it->second;
}

bool ReadAlignment::has_indel_calls_in(int rf, int rl) {  // RA:1471-1478
    auto* calls = get_indel_calls();
    if (!calls) return false;
    for (auto& kv : *calls) if (kv.first >= rf && kv.first <= rl) return true;
    return false;
}

bool ReadAlignment::move_indel_start(int indel_ref_pos, int new_ref_pos) {  // RA:1114-1153
    int displacement = new_ref_pos - indel_ref_pos;
    if (displacement == 0) return true;
    int n = (int)aln.size();
    std::vector<int> na(n, 0);
    int indel_next = -1;
    int cur_ref = first;
    for (int i = 0; i < n; i++) {
        int length = op_len(aln[i]);
        int op = op_code(aln[i]);
        if (is_indel(op) && cur_ref == indel_ref_pos + 1) {
            if (i == 0 || i == n - 1) return false;
            int opb = op_code(aln[i - 1]), lb = op_len(aln[i - 1]);
            if (!c_read(opb) || !c_ref(opb) || lb <= -displacement) return false;
            int opa = op_code(aln[i + 1]), la = op_len(aln[i + 1]);
            if (!c_read(opa) || !c_ref(opa) || la <= displacement) return false;
            na[i - 1] = aln_value(lb + displacement, opb);
            na[i] = aln[i];
            na[i + 1] = aln_value(la - displacement, opa);
            indel_next = i + 1;
        } else if (i != indel_next) {
            na[i] = aln[i];
        }
        if (c_ref(aln[i])) cur_ref += length;
    }
    if (indel_next < 0) return false;
    aln = na;
    calls_updated = false;
    return true;
}

void ReadAlignment::realign_start(int new_first, int first_match_len, int ref_pos_after, int aln_read_pos_after) {  // RA:1372-1418
    std::vector<int> list;
    list.push_back(aln_value(first_match_len, OP_M));
    int next_ref = new_first + first_match_len;
    int unk_read = aln_read_pos_after - first_match_len;
    int unk_ref = ref_pos_after - next_ref;
    if (unk_ref < 0 || unk_read < 0) return;
    int diff = unk_read - unk_ref;
    if (diff == 0) {
        if (unk_read > 0) list.push_back(aln_value(unk_read, OP_M));
    } else if (diff > 0) {
        list.push_back(aln_value(diff, OP_I));
        if (unk_ref > 0) list.push_back(aln_value(unk_ref, OP_M));
    } else {
        list.push_back(aln_value(-diff, OP_D));
        if (unk_read > 0) list.push_back(aln_value(unk_read, OP_M));
    }
    int cur_read = 0;
    bool copy = false;
    for (int v : aln) {
        int length = op_len(v), op = op_code(v);
        bool cd = c_read(v);
        if (copy) list.push_back(v);
        if (cd) {
            if (!copy && aln_read_pos_after < cur_read + length) {
                copy = true;
                int d = cur_read + length - aln_read_pos_after;
                list.push_back(aln_value(d, op));
            }
            cur_read += length;
        }
    }
    first = new_first;
    aln = list;
    calls_updated = false;
}

void ReadAlignment::realign_end(int ref_pos_before, int aln_pos_before, int final_match_ref_start, int final_match_len) {  // RA:1427-1469
    int bp_end_read = read_length - aln_pos_before - 1;
    std::vector<int> list;
    int cur_read = 0;
    bool copy = true;
    for (int v : aln) {
        int length = op_len(v), op = op_code(v);
        bool cd = c_read(v);
        if (cd) {
            if (copy && aln_pos_before < cur_read + length) {
                int d = cur_read + length - aln_pos_before - 1;
                list.push_back(aln_value(length - d, op));
                copy = false;
            }
            cur_read += length;
        }
        if (copy) list.push_back(v);
    }
    int unk_ref = final_match_ref_start - ref_pos_before - 1;
    int unk_read = bp_end_read - final_match_len;
    if (unk_ref < 0 || unk_read < 0) return;
    int diff = unk_read - unk_ref;
    if (diff == 0) {
        if (unk_read > 0) list.push_back(aln_value(unk_read, OP_M));
    } else if (diff > 0) {
        list.push_back(aln_value(diff, OP_I));
        if (unk_ref > 0) list.push_back(aln_value(unk_ref, OP_M));
    } else {
        list.push_back(aln_value(-diff, OP_D));
        if (unk_read > 0) list.push_back(aln_value(unk_read, OP_M));
    }
    list.push_back(aln_value(final_match_len, OP_M));
    aln = list;
    last = final_match_ref_start + final_match_len - 1;
    calls_updated = false;
}

int ReadAlignment::soft_clip_start() const { return op_code(aln[0]) == OP_S ? op_len(aln[0]) : 0; }
int ReadAlignment::soft_clip_end() const { return op_code(aln.back()) == OP_S ? op_len(aln.back()) : 0; }

std::string ReadAlignment::cigar_string() const {
    static const char* CODES = "HDIMPNSX";
    std::string s;
    for (int v : aln) { s += std::to_string(op_len(v)); s.push_back(CODES[op_code(v)]); }
    return s;
}

// ---------------------------------------------------------------------------------------------
// SAM text -> SamRecord -> ReadAlignment
// ---------------------------------------------------------------------------------------------
bool parse_sam_line(const std::string& line, SamRecord& r) {
    if (line.empty() || line[0] == '@') return false;
    std::vector<std::string> f;
    size_t s = 0;
    while (true) {
        size_t e = line.find('\t', s);
        f.push_back(line.substr(s, e == std::string::npos ? std::string::npos : e - s));
        if (e == std::string::npos) break;
        s = e + 1;
    }
    if (f.size() < 11) return false;
    r = SamRecord();
    r.qname = f[0]; r.flag = std::atoi(f[1].c_str()); r.rname = f[2]; r.pos = std::atoi(f[3].c_str());
    r.mapq = std::atoi(f[4].c_str()); r.cigar = f[5]; r.rnext = f[6]; r.pnext = std::atoi(f[7].c_str());
    r.tlen = std::atoi(f[8].c_str()); r.seq = f[9]; r.qual = f[10];
    for (size_t i = 11; i < f.size(); i++) {
        const std::string& t = f[i];
        if (t.size() < 5) continue;
        if (!t.compare(0, 5, "NH:i:")) { r.has_nh = true; r.nh = std::atoi(t.c_str() + 5); }
        else if (!t.compare(0, 5, "RG:Z:")) { r.has_rg = true; r.rg = t.substr(5); }
        else if (!t.compare(0, 5, "NM:i:")) { r.has_nm = true; r.nm = std::atoi(t.c_str() + 5); }
    }
    return true;
}

static int cigar_ref_length(const std::string& cigar) {  // htsjdk Cigar.getReferenceLength: M D N = X
    int n = 0, tot = 0;
    for (char c : cigar) {
        if (std::isdigit((unsigned char)c)) n = n * 10 + (c - '0');
        else { if (c == 'M' || c == 'D' || c == 'N' || c == '=' || c == 'X') tot += n; n = 0; }
    }
    return tot;
}

ReadPtr AlignmentReader::load(const SamRecord& rec) {
    // isSameAlignment, RAFR:292-306 (lastRecord is updated before the conversion/filters, RAFR:338-339)
    if (have_last) {
        bool same = last.pos == rec.pos;
        bool p1 = last.flag & FLAG_PAIRED, p2 = rec.flag & FLAG_PAIRED;
        if (same && p1 != p2) same = false;
        if (same && p1 && ((last.flag & FLAG_FIRST_OF_PAIR) != (rec.flag & FLAG_FIRST_OF_PAIR))) same = false;
        if (same && last.qname != rec.qname) same = false;
        if (same) return nullptr;
    }
    last = rec;
    have_last = true;
    ReadPtr a;
    try {
        // loadAlignment, RAFR:219-272
        int flags = rec.flag;
        bool unmapped_ref = rec.rname == "*";
        // isMultiple RAFR:284-291
        bool multiple;
        if (rec.flag & FLAG_SECONDARY) multiple = true;
        else if (rec.has_nh && rec.nh > 1) multiple = true;
        else if (rec.has_nh && rec.nh == 1) multiple = false;
        else multiple = rec.mapq < min_mq;
        if (multiple) flags += FLAG_MULTIPLE_ALN;
        bool mate_same = (rec.rnext == "=") || (rec.rnext == rec.rname);
        if (!mate_same) flags += FLAG_MATE_DIFFERENT_SEQUENCE;  // reference index != mate reference index
        a = std::make_shared<ReadAlignment>();
        a->seq_name = unmapped_ref ? std::string() : rec.rname;
        a->seq_idx = genome ? genome->find(rec.rname) : -1;
        a->first = rec.pos;
        bool no_cigar = rec.cigar == "*";
        a->last = no_cigar ? rec.pos : rec.pos + cigar_ref_length(rec.cigar) - 1;  // htsjdk getAlignmentEnd
        a->read_length = rec.seq == "*" ? 0 : (int)rec.seq.size();
        a->flags = flags;
        if (!no_cigar) a->set_cigar_string(rec.cigar);
        else if (!(flags & FLAG_READ_UNMAPPED)) throw std::invalid_argument("null CIGAR for aligned read");  // RA:1225
        a->mapq = std::min(rec.mapq, 127);
        if (rec.has_rg) a->read_group = rec.rg;
        if (rec.seq != "*") {
            a->bases = rec.seq;
            a->has_bases = true;
            if (rec.qual != "*") a->set_quality_scores(rec.qual);
        }
        a->name = rec.qname;
        if (!(flags & FLAG_READ_UNMAPPED)) a->update_allele_calls_info();  // surfaces inconsistent CIGAR/SEQ like the Java path would
    } catch (const std::exception&) {
        return nullptr;  // RAFR:341-346: conversion failure => record skipped
    }
    if ((a->flags & filter_flags) != 0) return nullptr;       // RAFR:350
    if ((a->flags & required_flags) != required_flags) return nullptr;
    return a;
}

}  // namespace ngsep_oracle
