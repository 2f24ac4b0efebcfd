// ngsep_b200 — command line with the option ids of the reference modules (main/CommandsDescriptor.xml:480-703):
//   ngsep_b200 SingleSampleVariantsDetector -i in.sam -r ref.fa -o prefix [options]
//   ngsep_b200 MultisampleVariantsDetector -r ref.fa -o out.vcf [options] a.sam b.sam ...      (MSVD.main, MSVD:412-419)
// Reads SAM text or BAM (the reference reads SAM/BAM/CRAM through htsjdk, RAFR:171-183; both are decoded natively here),
// writes <prefix>.vcf like SSVD:926.  Errors follow the reference: a missing -i / -r is reported as in SSVD:595-596,
// an unknown option prints usage and exits 1 (CommandsDescriptor.java:441-445).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <malloc.h>
#include <iostream>
#include <iterator>
#include <unordered_map>

#include <zlib.h>

#include "pipeline.hpp"

using namespace ngsep_b200::host;

namespace {

struct OwnedBatch {
    std::vector<int32_t> seq_idx, first;
    std::vector<uint16_t> flags;
    std::vector<uint8_t> mapq;
    std::vector<int16_t> rg, nh;
    std::vector<int64_t> cigar_off{0}, base_off{0};
    std::vector<uint32_t> cigar;
    std::vector<uint8_t> bases, quals;
    std::vector<uint64_t> name_hash;
    BatchView view() const {
        BatchView v;
        v.n = (int64_t)first.size(); v.seq_idx = seq_idx.data(); v.first = first.data(); v.flags = flags.data(); v.mapq = mapq.data();
        v.rg = rg.data(); v.nh = nh.data(); v.cigar_off = cigar_off.data(); v.cigar = cigar.data(); v.base_off = base_off.data();
        v.bases = bases.data(); v.quals = quals.data(); v.name_hash = name_hash.empty() ? nullptr : name_hash.data();
        return v;
    }
    void append_from(const OwnedBatch& o, size_t i, int16_t new_rg) {   // record i of o, without its name hash
        seq_idx.push_back(o.seq_idx[i]); first.push_back(o.first[i]); flags.push_back(o.flags[i]); mapq.push_back(o.mapq[i]);
        rg.push_back(new_rg); nh.push_back(o.nh[i]);
        cigar.insert(cigar.end(), o.cigar.begin() + o.cigar_off[i], o.cigar.begin() + o.cigar_off[i + 1]);
        cigar_off.push_back((int64_t)cigar.size());
        bases.insert(bases.end(), o.bases.begin() + o.base_off[i], o.bases.begin() + o.base_off[i + 1]);
        quals.insert(quals.end(), o.quals.begin() + o.base_off[i], o.quals.begin() + o.base_off[i + 1]);
        base_off.push_back((int64_t)bases.size());
    }
};

struct RgTable { std::vector<std::string> ids, samples; std::unordered_map<std::string, int> index; };

bool load_fasta(const std::string& path, Genome& g) {
    std::ifstream in(path);
    if (!in) return false;
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (!line.empty() && line[0] == '>') {
            size_t e = line.find_first_of(" \t");  // sequences/io/FastaFileReader.java:176-185
            g.names.push_back(line.substr(1, e == std::string::npos ? std::string::npos : e - 1));
            g.seqs.emplace_back();
        } else if (!g.seqs.empty()) g.seqs.back() += line;
    }
    return true;
}

// ---- alignment input: SAM text or BAM (BGZF) ----------------------------------------------------------------------------
// The reference reads SAM / BAM / CRAM through htsjdk (RAFR:171-183); here SAM text and BAM are decoded natively into the same
// struct-of-arrays batch.  rgs (optional): the @RG table shared by every input of a multi-sample run (ids in header order,
// RAFR:184-197); read groups that only appear on records get an index too (they belong to no sample).
struct Loader {
    const Genome& g;
    OwnedBatch& b;
    RgTable* rgs;
    std::unordered_map<std::string, int> seq_of, rg_local;
    std::string err;
    Loader(const Genome& genome, OwnedBatch& batch, RgTable* table) : g(genome), b(batch), rgs(table) {
        for (size_t i = 0; i < g.names.size(); i++) seq_of[g.names[i]] = (int)i;
    }
    std::unordered_map<std::string, int>& rg_of() { return rgs ? rgs->index : rg_local; }

    bool header_line(const std::string& line) {
        if (!line.compare(0, 3, "@SQ")) {  // header sequences must exist in the FASTA with the same length (RAFR:199-214)
            size_t sn = line.find("SN:"), ln = line.find("LN:");
            if (sn != std::string::npos && ln != std::string::npos) {
                std::string name = line.substr(sn + 3, line.find('\t', sn) - sn - 3);
                long len = std::atol(line.c_str() + ln + 3);
                auto it = seq_of.find(name);
                if (it == seq_of.end()) { err = "Inconsistent file header. Sequence " + name + " not present in the reference sequences"; return false; }
                if ((long)g.seqs[it->second].size() != len) { err = "Inconsistent length in file header. Sequence " + name; return false; }
            }
        }
        if (rgs && !line.compare(0, 3, "@RG")) {
            std::string id, sm;
            size_t s0 = 0;
            while (true) {
                size_t e = line.find('\t', s0);
                std::string tok = line.substr(s0, e == std::string::npos ? std::string::npos : e - s0);
                if (!tok.compare(0, 3, "ID:")) id = tok.substr(3); else if (!tok.compare(0, 3, "SM:")) sm = tok.substr(3);
                if (e == std::string::npos) break;
                s0 = e + 1;
            }
            if (sm.empty()) sm = id;                                                          // RAFR:188
            auto it = rgs->index.find(id);
            if (it == rgs->index.end()) { rgs->index[id] = (int)rgs->ids.size(); rgs->ids.push_back(id); rgs->samples.push_back(sm); }
            else if (rgs->samples[it->second] != sm) {                                        // RAFR:194-196, MSVD:500-502
                err = "The read group ID: " + id + " is associated to two different samples: " + sm + " and " + rgs->samples[it->second] + ". Read group ids should be unique across samples";
                return false;
            }
        }
        return true;
    }

    // one decoded record: cigar in the reference's op codes (H0 D1 I2 M3 P4 N5 S6 X7), quality chars phred+33 (empty = missing)
    void record(const std::string& qname, int flag, int seq, int pos, int mapq, const std::vector<uint32_t>& cg, const std::string& seq_bases,
                const std::string& qual, bool has_seq, const std::string* rg_id, int nh_tag) {
        b.seq_idx.push_back(seq);
        b.first.push_back(pos);
        b.flags.push_back((uint16_t)flag);
        b.mapq.push_back((uint8_t)std::min(255, mapq));
        int16_t rg = -1;
        if (rg_id) {
            auto r = rg_of().emplace(*rg_id, (int)rg_of().size());
            if (rgs && r.second) { rgs->ids.push_back(*rg_id); rgs->samples.push_back(std::string(1, '\0')); }   // on records only: no sample
            rg = (int16_t)r.first->second;
        }
        b.rg.push_back(rg);
        b.nh.push_back((int16_t)std::max(-1, std::min(32000, nh_tag)));
        b.cigar.insert(b.cigar.end(), cg.begin(), cg.end());
        b.cigar_off.push_back((int64_t)b.cigar.size());
        if (has_seq) {
            b.bases.insert(b.bases.end(), seq_bases.begin(), seq_bases.end());
            size_t L = seq_bases.size();
            if (qual.empty()) b.quals.insert(b.quals.end(), L, (uint8_t)'+');  // RA:1025
            else { for (size_t k = 0; k < L; k++) b.quals.push_back(k < qual.size() ? (uint8_t)std::min(127, (int)(unsigned char)qual[k]) : (uint8_t)38); }  // RA:581-595
        }
        b.base_off.push_back((int64_t)b.bases.size());
        b.name_hash.push_back(std::hash<std::string>()(qname));
    }
};

bool load_sam_text(std::istream& in, Loader& ld) {
    std::string line;
    std::vector<std::string> f;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        if (line[0] == '@') { if (!ld.header_line(line)) return false; continue; }
        f.clear();
        size_t s = 0;
        while (true) { size_t e = line.find('\t', s); f.push_back(line.substr(s, e == std::string::npos ? std::string::npos : e - s)); if (e == std::string::npos) break; s = e + 1; }
        if (f.size() < 11) continue;
        auto it = ld.seq_of.find(f[2]);
        if (it == ld.seq_of.end()) continue;  // unmapped / unknown reference
        std::vector<uint32_t> cg;
        if (f[5] != "*") {
            int n = 0; bool bad = false;
            for (char c : f[5]) {
                if (c >= '0' && c <= '9') n = n * 10 + (c - '0');
                else { const char* codes = "HDIMPNSX"; const char* p = std::strchr(codes, c); int op = p ? (int)(p - codes) : (c == '=' ? 3 : -1); if (op < 0) { bad = true; break; } cg.push_back(cig_make(n, op)); n = 0; }
            }
            if (bad) continue;
        }
        const std::string* rg = nullptr;
        std::string rg_id;
        int nh = -1;
        for (size_t k = 11; k < f.size(); k++) {
            if (!f[k].compare(0, 5, "RG:Z:")) { rg_id = f[k].substr(5); rg = &rg_id; }
            else if (!f[k].compare(0, 5, "NH:i:")) nh = std::atoi(f[k].c_str() + 5);
        }
        ld.record(f[0], std::atoi(f[1].c_str()), it->second, std::atoi(f[3].c_str()), std::atoi(f[4].c_str()), cg, f[9], f[10] == "*" ? std::string() : f[10],
                  f[9] != "*", rg, nh);
    }
    return true;
}

// BGZF: a series of gzip members with a "BC" extra field holding the block size (SAM specification section 4.1)
bool bgzf_inflate(const std::string& path, std::vector<unsigned char>& out, std::string& err) {
    std::ifstream in(path, std::ios::binary);
    if (!in) { err = "cannot open " + path; return false; }
    std::vector<unsigned char> raw((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    size_t p = 0;
    while (p + 18 <= raw.size()) {
        if (raw[p] != 0x1f || raw[p + 1] != 0x8b || raw[p + 2] != 8 || !(raw[p + 3] & 4)) { err = path + ": not a BGZF file"; return false; }
        const size_t xlen = raw[p + 10] | (raw[p + 11] << 8);
        if (p + 12 + xlen > raw.size()) { err = path + ": truncated BGZF block"; return false; }
        size_t x = p + 12, bsize = 0;
        while (x + 4 <= p + 12 + xlen) {
            const size_t slen = raw[x + 2] | (raw[x + 3] << 8);
            if (raw[x] == 'B' && raw[x + 1] == 'C' && slen == 2 && x + 6 <= p + 12 + xlen) bsize = (size_t)(raw[x + 4] | (raw[x + 5] << 8)) + 1;
            x += 4 + slen;
        }
        // header (12) + extra field + at least an empty deflate stream + CRC32 and ISIZE (8) must fit the block, the block the file
        if (!bsize || bsize < xlen + 20 || p + bsize > raw.size()) { err = path + ": truncated BGZF block"; return false; }
        const size_t cdata = p + 12 + xlen, clen = bsize - xlen - 20;
        const size_t isize = raw[p + bsize - 4] | (raw[p + bsize - 3] << 8) | (raw[p + bsize - 2] << 16) | ((size_t)raw[p + bsize - 1] << 24);
        if (isize > 65536) { err = path + ": corrupt BGZF block"; return false; }   // a BGZF block inflates to at most 64 KiB
        if (isize) {
            const size_t o = out.size();
            out.resize(o + isize);
            z_stream zs{};
            if (inflateInit2(&zs, -15) != Z_OK) { err = "zlib init failed"; return false; }
            zs.next_in = raw.data() + cdata; zs.avail_in = (uInt)clen;
            zs.next_out = out.data() + o; zs.avail_out = (uInt)isize;
            const int rc = inflate(&zs, Z_FINISH);
            inflateEnd(&zs);
            if (rc != Z_STREAM_END || zs.total_out != isize) { err = path + ": corrupt BGZF block"; return false; }
        }
        p += bsize;
    }
    return true;
}

bool load_bam(const std::string& path, Loader& ld) {
    std::vector<unsigned char> d;
    if (!bgzf_inflate(path, d, ld.err)) return false;
    auto u32 = [&](size_t o) { return (uint32_t)d[o] | ((uint32_t)d[o + 1] << 8) | ((uint32_t)d[o + 2] << 16) | ((uint32_t)d[o + 3] << 24); };
    auto u16 = [&](size_t o) { return (uint32_t)d[o] | ((uint32_t)d[o + 1] << 8); };
    if (d.size() < 12 || std::memcmp(d.data(), "BAM\1", 4)) { ld.err = path + ": not a BAM file"; return false; }
    size_t p = 4;
    const size_t l_text = u32(p); p += 4;
    if (l_text > d.size() - p || d.size() - p - l_text < 4) { ld.err = path + ": truncated BAM header"; return false; }
    {   // the SAM header text
        std::string text((const char*)d.data() + p, l_text);
        size_t s = 0;
        while (s < text.size()) {
            size_t e = text.find('\n', s);
            std::string line = text.substr(s, e == std::string::npos ? std::string::npos : e - s);
            while (!line.empty() && (line.back() == '\r' || line.back() == '\0')) line.pop_back();
            if (!line.empty() && line[0] == '@' && !ld.header_line(line)) return false;
            if (e == std::string::npos) break;
            s = e + 1;
        }
    }
    p += l_text;
    const size_t n_ref = u32(p); p += 4;
    if (n_ref > (d.size() - p) / 8) { ld.err = path + ": truncated BAM header"; return false; }   // every reference takes at least 8 bytes
    std::vector<int> ref_to_seq(n_ref, -1);
    for (size_t r = 0; r < n_ref; r++) {
        if (d.size() - p < 4) { ld.err = path + ": truncated BAM header"; return false; }
        const size_t l_name = u32(p); p += 4;
        if (l_name > d.size() - p || d.size() - p - l_name < 4) { ld.err = path + ": truncated BAM header"; return false; }
        std::string name((const char*)d.data() + p, l_name ? l_name - 1 : 0); p += l_name;
        const long l_ref = (long)u32(p); p += 4;
        auto it = ld.seq_of.find(name);
        if (it == ld.seq_of.end()) { ld.err = "Inconsistent file header. Sequence " + name + " not present in the reference sequences"; return false; }   // RAFR:199-214
        if ((long)ld.g.seqs[it->second].size() != l_ref) { ld.err = "Inconsistent length in file header. Sequence " + name; return false; }
        ref_to_seq[r] = it->second;
    }
    static const char kSeq[] = "=ACMGRSVTWYHKDBN";
    static const int kOp[9] = {kOpM, kOpI, kOpD, kOpN, kOpS, kOpH, kOpP, kOpM, kOpX};   // BAM MIDNSHP=X -> the reference's codes
    while (p + 4 <= d.size()) {
        const size_t block = u32(p); p += 4;
        if (p + block > d.size() || block < 32) { ld.err = path + ": truncated BAM record"; return false; }
        const size_t q = p;
        const int ref_id = (int)u32(q), pos = (int)u32(q + 4) + 1;
        const size_t l_name = d[q + 8];
        const int mapq = d[q + 9];
        const size_t n_cig = u16(q + 12);
        const int flag = (int)u16(q + 14);
        const size_t l_seq = u32(q + 16);
        // every length of the fixed part against the record block (block <= the bytes left in the file, checked above)
        if (l_seq > block || 32 + l_name + 4 * n_cig + (l_seq + 1) / 2 + l_seq > block) { ld.err = path + ": truncated BAM record"; return false; }
        const size_t rec_end = q + block;
        size_t o = q + 32;
        std::string qname((const char*)d.data() + o, l_name ? l_name - 1 : 0); o += l_name;
        std::vector<uint32_t> cg;
        bool bad = false;
        for (size_t k = 0; k < n_cig; k++, o += 4) { const uint32_t v = u32(o); if ((v & 15u) > 8u) bad = true; else cg.push_back(cig_make((int)(v >> 4), kOp[v & 15u])); }
        std::string seq(l_seq, 'N'), qual;
        for (size_t k = 0; k < l_seq; k++) { const unsigned char by = d[o + (k >> 1)]; seq[k] = kSeq[(k & 1) ? (by & 15) : (by >> 4)]; }
        o += (l_seq + 1) / 2;
        if (l_seq && d[o] != 0xff) { qual.resize(l_seq); for (size_t k = 0; k < l_seq; k++) qual[k] = (char)std::min(126, (int)d[o + k] + 33); }
        o += l_seq;
        const std::string* rg = nullptr;
        std::string rg_id;
        int nh = -1;
        bool aux_bad = false;
        while (o + 3 <= rec_end) {   // optional fields; every value is checked against the end of the record before it is read
            const char t0 = (char)d[o], t1 = (char)d[o + 1], ty = (char)d[o + 2];
            o += 3;
            long ival = 0; bool is_int = true;
            size_t need = 0;
            switch (ty) {
                case 'A': case 'c': case 'C': need = 1; break;
                case 's': case 'S': need = 2; break;
                case 'i': case 'I': case 'f': need = 4; break;
                case 'B': need = 5; break;
                default: break;
            }
            if (need > rec_end - o) { aux_bad = true; break; }
            switch (ty) {
                case 'A': case 'c': ival = (int8_t)d[o]; o += 1; break;
                case 'C': ival = d[o]; o += 1; break;
                case 's': ival = (int16_t)u16(o); o += 2; break;
                case 'S': ival = u16(o); o += 2; break;
                case 'i': ival = (int32_t)u32(o); o += 4; break;
                case 'I': ival = u32(o); o += 4; break;
                case 'f': is_int = false; o += 4; break;
                case 'Z': case 'H': {
                    size_t e = o;
                    while (e < rec_end && d[e]) e++;
                    if (e >= rec_end) { aux_bad = true; break; }   // no terminator inside the record
                    if (t0 == 'R' && t1 == 'G' && ty == 'Z') { rg_id.assign((const char*)d.data() + o, e - o); rg = &rg_id; }
                    o = e + 1; is_int = false; break;
                }
                case 'B': {
                    const char sub = (char)d[o];
                    const size_t cnt = u32(o + 1);
                    const size_t w = (sub == 'c' || sub == 'C') ? 1 : ((sub == 's' || sub == 'S') ? 2 : 4);
                    o += 5;
                    if (cnt > (rec_end - o) / w) { aux_bad = true; break; }
                    o += cnt * w; is_int = false; break;
                }
                default: o = rec_end; is_int = false; break;
            }
            if (aux_bad) break;
            if (is_int && t0 == 'N' && t1 == 'H' && ty != 'A') nh = (int)std::min(32000L, std::max(-1L, ival));
        }
        if (aux_bad) { ld.err = path + ": truncated BAM record (optional fields)"; return false; }
        p = q + block;
        if (bad || ref_id < 0 || (size_t)ref_id >= n_ref) continue;   // unplaced record / unknown operation: skipped like in the SAM reader
        ld.record(qname, flag, ref_to_seq[ref_id], pos, mapq, cg, l_seq ? seq : std::string(), qual, l_seq > 0, rg, nh);
    }
    return true;
}

bool load_alignments(const std::string& path, const Genome& g, OwnedBatch& b, std::string& err, RgTable* rgs = nullptr) {
    Loader ld(g, b, rgs);
    unsigned char magic[2] = {0, 0};
    { std::ifstream probe(path, std::ios::binary); if (!probe) { err = "cannot open " + path; return false; } probe.read((char*)magic, 2); }
    bool ok;
    if (magic[0] == 0x1f && magic[1] == 0x8b) ok = load_bam(path, ld);   // BGZF (htsjdk picks the reader from the file content too)
    else { std::ifstream in(path); ok = load_sam_text(in, ld); }
    if (!ok) err = ld.err;
    return ok;
}

// AlignmentsPileupGenerator.processFiles (APG:212-289): merge_order (model.cpp) gives the consumption order of the inputs' records.
void merge_inputs(const std::vector<OwnedBatch>& files, const AdmissionOptions& opt, OwnedBatch& out) {
    std::vector<BatchView> views;
    for (auto& f : files) views.push_back(f.view());
    std::vector<std::pair<int, int64_t>> order;
    merge_order(views, opt, order);
    for (auto& o : order) out.append_from(files[(size_t)o.first], (size_t)o.second, files[(size_t)o.first].rg[(size_t)o.second]);
}

void usage() {
    std::fprintf(stderr, "usage: ngsep_b200 RelativeAlleleCountsCalculator -i <alignments.sam|.bam> -r <genome.fa> [-o <out.txt>] [-m minRD] [-M maxRD] [-q minBaseQS] [-s]\n");
    std::fprintf(stderr, "usage: ngsep_b200 CoverageStats -i <alignments.sam|.bam> -r <genome.fa> [-o <out.txt>] [-minMQ N]\n");
    std::fprintf(stderr, "usage: ngsep_b200 MultisampleVariantsDetector -r <genome.fa> -o <out.vcf> [options] <alignments.sam|.bam>*\n"
                         "       ngsep_b200 SingleSampleVariantsDetector -i <alignments.sam|.bam> -r <genome.fa> -o <outputPrefix>\n"
                         "  [-sampleId S] [-ploidy N] [-psp] [-minMQ N] [-querySeq S -first N -last N] [-ignoreLowerCaseRef]\n"
                         "  [-maxAlnsPerStartPos N] [-p] [-s] [-ignore5 N] [-ignore3 N] [-h X] [-maxBaseQS N] [-minQuality N]\n"
                         "  [-embeddedSNVs] [-csb] [-knownSTRs file] [-knownVariants file.vcf] [--device N] [--gpus N] [--tile-len N]\n");
}

}  // namespace

int main(int argc, char** argv) {
    // The host stages build and drop arrays of hundreds of megabytes per sequence: keep the freed memory in the process instead of
    // returning it to the kernel after every sequence (measured: admission of a 1.2 M-read sequence 0.157 s -> 0.118 s on one thread,
    // 0.055 s on eight, where page faults on fresh mappings otherwise serialise the threads).  Own process: the library never does this.
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    if (argc >= 2 && (!std::strcmp(argv[1], "RelativeAlleleCountsCalculator") || !std::strcmp(argv[1], "RelativeAlleleCounts"))) {
        // RelativeAlleleCountsCalculator (main/CommandsDescriptor.xml:805-847: -i, -o, -r, -m, -M, -q, -s; -frs / -srs are not built)
        std::string in, ref, out;
        int min_rd = 10, max_rd = 1000, min_bq = 20, device = 0;
        bool secondary = false;
        for (int i = 2; i < argc; i++) {
            std::string a = argv[i];
            auto val = [&]() -> const char* { if (i + 1 >= argc) { usage(); std::exit(1); } return argv[++i]; };
            if (a == "-i") in = val(); else if (a == "-r") ref = val(); else if (a == "-o") out = val();
            else if (a == "-m") min_rd = std::atoi(val()); else if (a == "-M") max_rd = std::atoi(val()); else if (a == "-q") min_bq = std::atoi(val());
            else if (a == "-s") secondary = true; else if (a == "--device") device = std::atoi(val());
            else { std::fprintf(stderr, "Unrecognized option %s\n", a.c_str()); usage(); return 1; }
        }
        if (in.empty()) { std::fprintf(stderr, "The alignments input file is a required parameter\n"); return 2; }   // RelativeAlleleCountsCalculator.java:183
        if (ref.empty()) { std::fprintf(stderr, "The reference genome file is a required parameter\n"); return 2; }
        Genome g;
        if (!load_fasta(ref, g)) { std::fprintf(stderr, "cannot read %s\n", ref.c_str()); return 2; }
        OwnedBatch b;
        std::string err, text;
        if (!load_alignments(in, g, b, err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 2; }
        int rc = run_relative_allele_counts(g, b.view(), min_rd, max_rd, min_bq, secondary, device, text, err);
        if (rc != 0) { std::fprintf(stderr, "ngsep_b200: %s (code %d)\n", err.c_str(), rc); return 3; }
        if (out.empty()) std::fwrite(text.data(), 1, text.size(), stdout);
        else { std::ofstream os(out); os << text; }
        return 0;
    }
    if (argc >= 2 && !std::strcmp(argv[1], "CoverageStats")) {
        // CoverageStatisticsCalculator (main/CommandsDescriptor.xml:458-478: -i, -o, -minMQ; -r is needed here to name the sequences)
        std::string in, ref, out;
        int min_mq = 20, device = 0;
        for (int i = 2; i < argc; i++) {
            std::string a = argv[i];
            auto val = [&]() -> const char* { if (i + 1 >= argc) { usage(); std::exit(1); } return argv[++i]; };
            if (a == "-i") in = val(); else if (a == "-r") ref = val(); else if (a == "-o") out = val(); else if (a == "-minMQ") min_mq = std::atoi(val());
            else if (a == "--device") device = std::atoi(val());
            else { std::fprintf(stderr, "Unrecognized option %s\n", a.c_str()); usage(); return 1; }
        }
        if (in.empty()) { std::fprintf(stderr, "The alignments input file is a required parameter\n"); return 2; }   // CoverageStatisticsCalculator.java:115
        if (ref.empty()) { std::fprintf(stderr, "The reference genome file is a required parameter\n"); return 2; }
        Genome g;
        if (!load_fasta(ref, g)) { std::fprintf(stderr, "cannot read %s\n", ref.c_str()); return 2; }
        OwnedBatch b;
        std::string err, text;
        if (!load_alignments(in, g, b, err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 2; }
        int rc = run_coverage_stats(g, b.view(), min_mq, device, text, err);
        if (rc != 0) { std::fprintf(stderr, "ngsep_b200: %s (code %d)\n", err.c_str(), rc); return 3; }
        if (out.empty()) std::fwrite(text.data(), 1, text.size(), stdout);   // "Writes to standard output unless the -o option is used"
        else { std::ofstream os(out); os << text; }
        return 0;
    }
    if (argc < 2 || (std::strcmp(argv[1], "SingleSampleVariantsDetector") && std::strcmp(argv[1], "MultisampleVariantsDetector"))) { usage(); return 1; }
    const bool multi = !std::strcmp(argv[1], "MultisampleVariantsDetector");
    DetectorOptions o;
    std::string in, ref, out;
    std::vector<std::string> inputs;
    std::string dump_batch;
    if (multi) out = "variants.vcf";   // MSVD.DEF_OUTPUT_FILE
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        if (multi && (a.empty() || a[0] != '-')) { for (; i < argc; i++) inputs.push_back(argv[i]); break; }   // MSVD:415-417
        auto val = [&]() -> const char* { if (i + 1 >= argc) { usage(); std::exit(1); } return argv[++i]; };
        if (a == "-i") in = val(); else if (a == "-r") ref = val(); else if (a == "-o") out = val();
        else if (a == "-sampleId") o.sample_id = val(); else if (a == "-ploidy") o.ploidy = std::atoi(val());
        else if (a == "-psp") o.print_sample_ploidy = true; else if (a == "-minMQ") o.min_mq = std::atoi(val());
        else if (a == "-querySeq") { o.query_seq = val(); o.has_query = true; } else if (a == "-first") o.query_first = std::atoi(val());
        else if (a == "-last") o.query_last = std::atoi(val()); else if (a == "-ignoreLowerCaseRef") o.ignore_lower = true;
        else if (a == "-maxAlnsPerStartPos") o.max_alns_per_start = std::atoi(val()); else if (a == "-p") o.non_unique_primary = true;
        else if (a == "-s") o.secondary = true; else if (a == "-ignore5") o.ignore5 = std::atoi(val());
        else if (a == "-ignore3") o.ignore3 = std::atoi(val()); else if (a == "-h") { o.het_rate = std::atof(val()); o.het_modified = true; }
        else if (a == "-maxBaseQS") o.max_base_qs = std::atoi(val()); else if (a == "-minQuality") o.min_quality = std::atoi(val());
        else if (a == "-embeddedSNVs") o.call_embedded = true; else if (a == "-csb") o.csb = true;
        else if (a == "-knownSTRs" && !multi) o.known_strs = val();
        else if (a == "-knownVariants" && !multi) o.known_variants = val();
        else if (a == "--device") o.device = std::atoi(val()); else if (a == "--tile-len") o.tile_len = std::atoi(val());
        else if (a == "--gpus") o.n_devices = std::max(1, std::atoi(val()));   // region tiles dealt to this many GPUs (device .. device + n - 1)
        else if (a == "--host-only") o.host_only = true;   // test hook: admission + event pre-pass only, indel / STR records only (no device)
        else if (a == "--dump-batch") dump_batch = val();  // test hook: the decoded records, as loaded, in a flat binary file
        else { std::fprintf(stderr, "Unrecognized option %s\n", a.c_str()); usage(); return 1; }
    }
    if (multi) {
        if (ref.empty()) { std::fprintf(stderr, "The reference genome file is a required parameter\n"); return 2; }   // MSVD:423
        Genome g;
        if (!load_fasta(ref, g)) { std::fprintf(stderr, "cannot read %s\n", ref.c_str()); return 2; }
        RgTable rgs;
        std::vector<OwnedBatch> files(inputs.size());
        std::string err;
        for (size_t f = 0; f < inputs.size(); f++) if (!load_alignments(inputs[f], g, files[f], err, &rgs)) { std::fprintf(stderr, "%s\n", err.c_str()); return 2; }
        AdmissionOptions ao;
        ao.min_mq = o.min_mq; ao.secondary = o.secondary; ao.non_unique_primary = o.non_unique_primary;
        OwnedBatch merged;
        const OwnedBatch* b = files.empty() ? &merged : &files[0];
        if (files.size() > 1) { merge_inputs(files, ao, merged); b = &merged; }   // one input: processFile, no merge (APG:216-219)
        // samples: the read groups of the headers only (a group seen on records alone belongs to no sample)
        std::vector<std::string> hdr_ids, hdr_sms;
        std::vector<int> hdr_index;
        for (size_t k = 0; k < rgs.ids.size(); k++) if (rgs.samples[k] != std::string(1, '\0')) { hdr_ids.push_back(rgs.ids[k]); hdr_sms.push_back(rgs.samples[k]); hdr_index.push_back((int)k); }
        std::vector<SampleDef> samples = samples_from_read_groups(hdr_ids, hdr_sms, o.ploidy);
        for (auto& s : samples) for (int& r : s.rgs) r = hdr_index[r];
        GpuMultisampleVariantsDetector det(o, samples, (int)rgs.ids.size());
        det.set_genome(&g);
        std::string vcf;
        int rc = det.run(b->view(), vcf, true);
        if (rc != 0) { std::fprintf(stderr, "ngsep_b200: %s (code %d)\n", det.error().c_str(), rc); return 3; }
        std::ofstream os(out);   // MSVD:452
        os << vcf;
        const RunStats& s = det.stats();
        std::fprintf(stderr, "samples %zu records %llu admitted %llu events %llu snv %llu indel %llu covered %llu kernel_ms %.3f\n", samples.size(),
                     (unsigned long long)s.records, (unsigned long long)s.admitted, (unsigned long long)s.events, (unsigned long long)s.snv_calls,
                     (unsigned long long)s.indel_calls, (unsigned long long)s.covered_positions, s.kernel_ms);
        return 0;
    }
    if (in.empty()) { std::fprintf(stderr, "The input file with alignments is a required parameter\n"); return 2; }
    if (ref.empty()) { std::fprintf(stderr, "The reference genome file is a required parameter\n"); return 2; }
    Genome g;
    if (!load_fasta(ref, g)) { std::fprintf(stderr, "cannot read %s\n", ref.c_str()); return 2; }
    OwnedBatch b;
    std::string err;
    if (!load_alignments(in, g, b, err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 2; }
    if (!dump_batch.empty()) {   // seq_idx, first, flags, mapq, rg, nh, cigar_off, cigar, base_off, bases, quals, name_hash — sizes first
        std::ofstream df(dump_batch, std::ios::binary);
        auto put = [&](const void* p, size_t bytes) { uint64_t n = bytes; df.write((const char*)&n, 8); df.write((const char*)p, (std::streamsize)bytes); };
        put(b.seq_idx.data(), b.seq_idx.size() * 4); put(b.first.data(), b.first.size() * 4); put(b.flags.data(), b.flags.size() * 2);
        put(b.mapq.data(), b.mapq.size()); put(b.rg.data(), b.rg.size() * 2); put(b.nh.data(), b.nh.size() * 2);
        put(b.cigar_off.data(), b.cigar_off.size() * 8); put(b.cigar.data(), b.cigar.size() * 4); put(b.base_off.data(), b.base_off.size() * 8);
        put(b.bases.data(), b.bases.size()); put(b.quals.data(), b.quals.size()); put(b.name_hash.data(), b.name_hash.size() * 8);
    }
    GpuSingleSampleVariantsDetector det(o);
    det.set_genome(&g);
    std::string vcf;
    int rc = det.run(b.view(), vcf, true);
    if (rc != 0) { std::fprintf(stderr, "ngsep_b200: %s (code %d)\n", det.error().c_str(), rc); return 3; }
    std::ofstream os(out + ".vcf");
    os << vcf;
    const RunStats& s = det.stats();
    std::fprintf(stderr, "records %llu admitted %llu events %llu versions %llu snv %llu indel %llu covered %llu kernel_ms %.3f\n",
                 (unsigned long long)s.records, (unsigned long long)s.admitted, (unsigned long long)s.events, (unsigned long long)s.versions,
                 (unsigned long long)s.snv_calls, (unsigned long long)s.indel_calls, (unsigned long long)s.covered_positions, s.kernel_ms);
    return 0;
}
