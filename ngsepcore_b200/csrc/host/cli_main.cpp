// ngsep_b200 — command line with the option ids of the reference modules (main/CommandsDescriptor.xml:480-703):
//   ngsep_b200 SingleSampleVariantsDetector -i in.sam -r ref.fa -o prefix [options]
//   ngsep_b200 MultisampleVariantsDetector -r ref.fa -o out.vcf [options] a.sam b.sam ...      (MSVD.main, MSVD:412-419)
// Reads SAM text or BAM (the reference reads SAM/BAM/CRAM through htsjdk, RAFR:171-183; both are decoded natively here),
// writes <prefix>.vcf like SSVD:926.  Errors follow the reference: a missing -i / -r is reported as in SSVD:595-596,
// an unknown option prints usage and exits 1 (CommandsDescriptor.java:441-445).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <malloc.h>
#include <iostream>
#include <iterator>
#include <string_view>
#include <thread>
#include <unordered_map>

#include <zlib.h>

#include "pipeline.hpp"

using namespace ngsep_b200::host;

namespace {

// std::allocator whose value-less construct() leaves the storage as it is: resize() of a byte array then touches no page (the
// threads that fill the array do, each its own part)
template <typename T>
struct NoInitAlloc : std::allocator<T> {
    template <typename U> struct rebind { using other = NoInitAlloc<U>; };
    NoInitAlloc() = default;
    template <typename U> NoInitAlloc(const NoInitAlloc<U>&) {}
    template <typename U, typename... A> void construct(U* p, A&&... a) {
        if constexpr (sizeof...(A) == 0) ::new ((void*)p) U; else ::new ((void*)p) U(std::forward<A>(a)...);
    }
};
using ByteArray = std::vector<uint8_t, NoInitAlloc<uint8_t>>;

struct OwnedBatch {
    std::vector<int32_t> seq_idx, first;
    std::vector<uint16_t> flags;
    std::vector<uint8_t> mapq;
    std::vector<int16_t> rg, nh;
    std::vector<int64_t> cigar_off{0}, base_off{0};
    std::vector<uint32_t> cigar;
    ByteArray bases, quals;
    std::vector<uint64_t> name_hash;
    BatchView view() const {
        BatchView v;
        v.n = (int64_t)first.size(); v.seq_idx = seq_idx.data(); v.first = first.data(); v.flags = flags.data(); v.mapq = mapq.data();
        v.rg = rg.data(); v.nh = nh.data(); v.cigar_off = cigar_off.data(); v.cigar = cigar.data(); v.base_off = base_off.data();
        v.bases = bases.data(); v.quals = quals.data(); v.name_hash = name_hash.empty() ? nullptr : name_hash.data();
        return v;
    }
    void swap_storage(OwnedBatch& o) {   // (used to release a batch's arrays)
        seq_idx.swap(o.seq_idx); first.swap(o.first); flags.swap(o.flags); mapq.swap(o.mapq); rg.swap(o.rg); nh.swap(o.nh); cigar_off.swap(o.cigar_off);
        base_off.swap(o.base_off); cigar.swap(o.cigar); bases.swap(o.bases); quals.swap(o.quals); name_hash.swap(o.name_hash);
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

// FASTA: the whole file in one read, lines found with memchr (a 3 Gb genome is 50 M lines: no stream, no string per line)
bool load_fasta(const std::string& path, Genome& g) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    in.seekg(0, std::ios::end);
    const std::streamoff file_size = in.tellg();
    in.seekg(0);
    std::vector<char, NoInitAlloc<char>> raw((size_t)std::max<std::streamoff>(0, file_size));
    if (!raw.empty() && !in.read(raw.data(), (std::streamsize)raw.size())) return false;
    const char* p = raw.data();
    const char* const end = p + raw.size();
    while (p < end) {
        const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
        const char* e = nl ? nl : end;
        const char* next = nl ? nl + 1 : end;
        if (e > p && e[-1] == '\r') e--;
        if (e > p && *p == '>') {
            const char* s = p + 1;
            const char* t = s;
            while (t < e && *t != ' ' && *t != '\t') t++;   // sequences/io/FastaFileReader.java:176-185
            g.names.emplace_back(s, (size_t)(t - s));
            g.seqs.emplace_back();
            // the sequence's bases up to the next header: reserved once
            const char* h = next;
            while (h < end) { if (*h == '>') break; const char* n2 = (const char*)std::memchr(h, '\n', (size_t)(end - h)); if (!n2) { h = end; break; } h = n2 + 1; }
            g.seqs.back().reserve((size_t)(h - next));
        } else if (e > p && !g.seqs.empty()) g.seqs.back().append(p, (size_t)(e - p));
        p = next;
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

// f(t) for t in [0, n) on n threads (the calling one included)
template <typename F>
void on_threads(int n, F&& f) {
    std::vector<std::thread> th;
    for (int t = 1; t < n; t++) th.emplace_back(f, t);
    f(0);
    for (auto& x : th) x.join();
}

// Bytes that are not cleared when they are allocated (the inflated file: 1 GB and more, written once by the inflating threads)
struct RawBytes {
    unsigned char* p = nullptr;
    size_t n = 0;
    RawBytes() = default;
    RawBytes(const RawBytes&) = delete;
    RawBytes& operator=(const RawBytes&) = delete;
    ~RawBytes() { std::free(p); }
    bool alloc(size_t k) { std::free(p); p = (unsigned char*)std::malloc(k ? k : 1); n = p ? k : 0; return p != nullptr; }
    size_t size() const { return n; }
    const unsigned char* data() const { return p; }
    unsigned char operator[](size_t i) const { return p[i]; }
};

// BGZF: a series of gzip members with a "BC" extra field holding the block size (SAM specification section 4.1).  The members are
// independent deflate streams whose inflated sizes stand in their trailers: one pass finds the blocks and where their bytes go,
// then the blocks are inflated on all host threads.
bool bgzf_inflate(const std::string& path, RawBytes& out, std::string& err) {
    std::ifstream in(path, std::ios::binary);
    if (!in) { err = "cannot open " + path; return false; }
    in.seekg(0, std::ios::end);
    const std::streamoff file_size = in.tellg();
    in.seekg(0);
    std::vector<unsigned char, NoInitAlloc<unsigned char>> raw((size_t)std::max<std::streamoff>(0, file_size));
    if (!raw.empty() && !in.read((char*)raw.data(), (std::streamsize)raw.size())) { err = "cannot read " + path; return false; }
    struct Block { size_t cdata, clen, isize, out; };
    std::vector<Block> blocks;
    size_t p = 0, total = 0;
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
        const size_t isize = raw[p + bsize - 4] | (raw[p + bsize - 3] << 8) | (raw[p + bsize - 2] << 16) | ((size_t)raw[p + bsize - 1] << 24);
        if (isize > 65536) { err = path + ": corrupt BGZF block"; return false; }   // a BGZF block inflates to at most 64 KiB
        if (isize) blocks.push_back({p + 12 + xlen, bsize - xlen - 20, isize, total});
        total += isize;
        p += bsize;
    }
    if (!out.alloc(total)) { err = path + ": not enough memory for the inflated file"; return false; }
    const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)host_threads(), blocks.size() / 16 + 1));
    std::atomic<size_t> next{0};
    std::atomic<bool> failed{false};
    on_threads(nt, [&](int) {
        z_stream zs{};
        if (inflateInit2(&zs, -15) != Z_OK) { failed.store(true); return; }
        for (;;) {
            const size_t k0 = next.fetch_add(16);
            if (k0 >= blocks.size() || failed.load()) break;
            for (size_t k = k0; k < std::min(blocks.size(), k0 + 16); k++) {
                const Block& bl = blocks[k];
                inflateReset(&zs);
                zs.next_in = raw.data() + bl.cdata; zs.avail_in = (uInt)bl.clen;
                zs.next_out = out.p + bl.out; zs.avail_out = (uInt)bl.isize;
                const int rc = inflate(&zs, Z_FINISH);
                if (rc != Z_STREAM_END || zs.total_out != bl.isize) { failed.store(true); break; }
            }
        }
        inflateEnd(&zs);
    });
    if (failed.load()) { err = path + ": corrupt BGZF block"; return false; }
    return true;
}

// What a reader thread decodes: a contiguous range of records as a batch of its own, read-group ids numbered locally in the order of
// their first appearance in the range.
struct Part { OwnedBatch b; std::vector<std::string> rg_ids; std::unordered_map<std::string, int> rg_index; std::string err; };

// The parts of a file, in file order, appended to the loader's batch; read-group ids get the index the sequential reader gives them
// (Loader::record: the order of their first appearance in the file), the first error in file order is the file's error.
bool merge_parts(Loader& ld, std::vector<Part>& parts) {
    const int nt = (int)parts.size();
    for (auto& part : parts) if (!part.err.empty()) { ld.err = part.err; return false; }   // (the first one in file order)
    // ---- the parts side by side ------------------------------------------------------------------------------------------------
    OwnedBatch& out = ld.b;
    std::vector<size_t> rec0((size_t)nt + 1, out.first.size()), cig0((size_t)nt + 1, out.cigar.size()), base0((size_t)nt + 1, out.bases.size());
    std::vector<std::vector<int16_t>> rg_map((size_t)nt);
    for (int t = 0; t < nt; t++) {
        Part& part = parts[(size_t)t];
        rec0[(size_t)t + 1] = rec0[(size_t)t] + part.b.first.size();
        cig0[(size_t)t + 1] = cig0[(size_t)t] + part.b.cigar.size();
        base0[(size_t)t + 1] = base0[(size_t)t] + part.b.bases.size();
        for (const std::string& id : part.rg_ids) {   // Loader::record's table, in the order of first appearance
            auto r = ld.rg_of().emplace(id, (int)ld.rg_of().size());
            if (ld.rgs && r.second) { ld.rgs->ids.push_back(id); ld.rgs->samples.push_back(std::string(1, '\0')); }   // on records only: no sample
            rg_map[(size_t)t].push_back((int16_t)r.first->second);
        }
    }
    const size_t n_all = rec0[(size_t)nt];
    out.seq_idx.resize(n_all); out.first.resize(n_all); out.flags.resize(n_all); out.mapq.resize(n_all); out.rg.resize(n_all); out.nh.resize(n_all);
    out.name_hash.resize(n_all); out.cigar_off.resize(n_all + 1); out.base_off.resize(n_all + 1);
    out.cigar.resize(cig0[(size_t)nt]); out.bases.resize(base0[(size_t)nt]); out.quals.resize(base0[(size_t)nt]);
    on_threads(nt, [&](int t) {
        OwnedBatch& b = parts[(size_t)t].b;
        const size_t n = b.first.size(), at = rec0[(size_t)t];
        if (!n) return;
        std::memcpy(out.seq_idx.data() + at, b.seq_idx.data(), n * sizeof(int32_t));
        std::memcpy(out.first.data() + at, b.first.data(), n * sizeof(int32_t));
        std::memcpy(out.flags.data() + at, b.flags.data(), n * sizeof(uint16_t));
        std::memcpy(out.mapq.data() + at, b.mapq.data(), n);
        std::memcpy(out.nh.data() + at, b.nh.data(), n * sizeof(int16_t));
        std::memcpy(out.name_hash.data() + at, b.name_hash.data(), n * sizeof(uint64_t));
        const std::vector<int16_t>& m = rg_map[(size_t)t];
        for (size_t i = 0; i < n; i++) out.rg[at + i] = b.rg[i] < 0 ? (int16_t)-1 : m[(size_t)b.rg[i]];
        for (size_t i = 0; i < n; i++) {   // (a part's offset arrays start with their own 0)
            out.cigar_off[at + i + 1] = (int64_t)cig0[(size_t)t] + b.cigar_off[i + 1];
            out.base_off[at + i + 1] = (int64_t)base0[(size_t)t] + b.base_off[i + 1];
        }
        if (!b.cigar.empty()) std::memcpy(out.cigar.data() + cig0[(size_t)t], b.cigar.data(), b.cigar.size() * sizeof(uint32_t));
        if (!b.bases.empty()) {
            std::memcpy(out.bases.data() + base0[(size_t)t], b.bases.data(), b.bases.size());
            std::memcpy(out.quals.data() + base0[(size_t)t], b.quals.data(), b.quals.size());
        }
        OwnedBatch().swap_storage(b);
    });
    return true;
}

// SAM text on all host threads: the file in one read, the header lines at its top handled one by one (Loader::header_line), the
// alignment lines cut into ranges at line ends and decoded into parts like BAM records.  A file with an '@' line below its first
// alignment line (the sequential reader takes it as a header line wherever it stands), a small file, or one thread: the
// sequential reader.
bool load_sam_parallel(const std::string& path, Loader& ld, bool& handled) {
    handled = false;
    const int threads = host_threads();
    if (threads < 2) return true;
    std::ifstream in(path, std::ios::binary);
    if (!in) return true;
    in.seekg(0, std::ios::end);
    const std::streamoff file_size = in.tellg();
    in.seekg(0);
    size_t part_records = 20000;
    if (const char* e = std::getenv("NGSEP_BAM_PART_RECORDS")) { const long v = std::atol(e); if (v >= 1) part_records = (size_t)v; }
    if (file_size < (std::streamoff)(part_records * 200)) return true;   // (about one part: not worth it)
    std::vector<char, NoInitAlloc<char>> raw((size_t)file_size);
    if (!in.read(raw.data(), (std::streamsize)raw.size())) return true;
    const char* const base = raw.data();
    const char* const end = base + raw.size();
    // the header block
    const char* p = base;
    std::vector<std::pair<const char*, const char*>> header;
    while (p < end && (*p == '@' || *p == '\n')) {
        const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
        const char* e = nl ? nl : end;
        if (e > p) header.emplace_back(p, e);
        p = nl ? nl + 1 : end;
    }
    const char* const body = p;
    for (const char* q = body; q < end;) {   // an '@' line further down: leave the file to the sequential reader
        const char* nl = (const char*)std::memchr(q, '\n', (size_t)(end - q));
        if (!nl) break;
        if (nl + 1 < end && nl[1] == '@') return true;
        q = nl + 1;
    }
    handled = true;
    for (auto& h : header) if (!ld.header_line(std::string(h.first, h.second))) return false;
    const size_t body_bytes = (size_t)(end - body);
    const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads, body_bytes / (part_records * 200) + 1));
    std::vector<const char*> cut((size_t)nt + 1, end);
    cut[0] = body;
    for (int t = 1; t < nt; t++) {
        const char* q = body + body_bytes * (size_t)t / (size_t)nt;
        if (q < cut[(size_t)t - 1]) q = cut[(size_t)t - 1];
        const char* nl = (const char*)std::memchr(q, '\n', (size_t)(end - q));
        cut[(size_t)t] = nl ? nl + 1 : end;
    }
    std::vector<Part> parts((size_t)nt);
    on_threads(nt, [&](int t) {
        Part& part = parts[(size_t)t];
        OwnedBatch& b = part.b;
        const char* q = cut[(size_t)t];
        const char* const stop = cut[(size_t)t + 1];
        {
            const size_t bytes = (size_t)(stop - q), guess = bytes / 150 + 16;
            b.seq_idx.reserve(guess); b.first.reserve(guess); b.flags.reserve(guess); b.mapq.reserve(guess); b.rg.reserve(guess); b.nh.reserve(guess);
            b.cigar_off.reserve(guess + 1); b.base_off.reserve(guess + 1); b.name_hash.reserve(guess); b.cigar.reserve(guess * 2);
            b.bases.reserve(bytes / 2 + 16); b.quals.reserve(bytes / 2 + 16);
        }
        std::vector<std::string_view> f;
        std::vector<uint32_t> cg;
        std::string key;
        while (q < stop) {
            const char* nl = (const char*)std::memchr(q, '\n', (size_t)(stop - q));
            const char* e = nl ? nl : stop;
            const char* next = nl ? nl + 1 : stop;
            if (e == q) { q = next; continue; }
            f.clear();
            for (const char* s0 = q;;) {
                const char* tab = (const char*)std::memchr(s0, '\t', (size_t)(e - s0));
                f.emplace_back(s0, (size_t)((tab ? tab : e) - s0));
                if (!tab) break;
                s0 = tab + 1;
            }
            q = next;
            if (f.size() < 11) continue;
            key.assign(f[2].data(), f[2].size());
            auto it = ld.seq_of.find(key);
            if (it == ld.seq_of.end()) continue;  // unmapped / unknown reference
            cg.clear();
            if (f[5] != "*") {
                int n = 0; bool bad = false;
                for (char c : f[5]) {
                    if (c >= '0' && c <= '9') n = n * 10 + (c - '0');
                    else { const char* codes = "HDIMPNSX"; const char* pc = std::strchr(codes, c); int op = (pc && c) ? (int)(pc - codes) : (c == '=' ? 3 : -1); if (op < 0) { bad = true; break; } cg.push_back(cig_make(n, op)); n = 0; }
                }
                if (bad) continue;
            }
            bool has_rg = false;
            std::string_view rg_id;
            int nh = -1;
            for (size_t k = 11; k < f.size(); k++) {
                if (f[k].size() >= 5 && !std::memcmp(f[k].data(), "RG:Z:", 5)) { rg_id = f[k].substr(5); has_rg = true; }
                else if (f[k].size() >= 5 && !std::memcmp(f[k].data(), "NH:i:", 5)) nh = std::atoi(std::string(f[k].substr(5)).c_str());
            }
            auto to_int = [](std::string_view v) { return std::atoi(std::string(v).c_str()); };
            // ---- Loader::record, into the part ---------------------------------------------------------------------------------
            b.seq_idx.push_back(it->second);
            b.first.push_back(to_int(f[3]));
            b.flags.push_back((uint16_t)to_int(f[1]));
            b.mapq.push_back((uint8_t)std::min(255, to_int(f[4])));
            int16_t rg = -1;
            if (has_rg) {
                key.assign(rg_id.data(), rg_id.size());
                auto ins = part.rg_index.emplace(key, (int)part.rg_ids.size());
                if (ins.second) part.rg_ids.push_back(key);
                rg = (int16_t)ins.first->second;
            }
            b.rg.push_back(rg);
            b.nh.push_back((int16_t)std::max(-1, std::min(32000, nh)));
            b.cigar.insert(b.cigar.end(), cg.begin(), cg.end());
            b.cigar_off.push_back((int64_t)b.cigar.size());
            if (f[9] != "*") {
                const size_t L = f[9].size(), at = b.bases.size();
                b.bases.resize(at + L);
                b.quals.resize(at + L);
                std::memcpy(b.bases.data() + at, f[9].data(), L);
                uint8_t* qs = b.quals.data() + at;
                if (f[10] == "*" || f[10].empty()) std::memset(qs, '+', L);                           // RA:1025
                else for (size_t k = 0; k < L; k++) qs[k] = k < f[10].size() ? (uint8_t)std::min(127, (int)(unsigned char)f[10][k]) : (uint8_t)38;   // RA:581-595
            }
            b.base_off.push_back((int64_t)b.bases.size());
            b.name_hash.push_back(std::hash<std::string_view>()(f[0]));
        }
    });
    return merge_parts(ld, parts);
}

// The records of a BAM file, decoded on all host threads.  One pass finds the record boundaries; contiguous ranges of records are
// then decoded independently into parts of the batch; the parts are laid side by side.  Read-group ids get the index the
// sequential reader gives them — the order of their first appearance in the file (RAFR:184-197 for the header's) — because the
// parts are merged in file order.
bool load_bam(const std::string& path, Loader& ld) {
    RawBytes d;
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
    // ---- record boundaries (one pass over the block sizes) ------------------------------------------------------------------
    std::vector<size_t> rec_at;
    while (p + 4 <= d.size()) {
        const size_t block = u32(p);
        if (p + 4 + block > d.size() || block < 32) { ld.err = path + ": truncated BAM record"; return false; }
        rec_at.push_back(p + 4);
        p += 4 + block;
    }
    // ---- the records, range by range ------------------------------------------------------------------------------------------
    const size_t n_rec = rec_at.size();
    size_t part_records = 20000;   // a part is not worth a thread below this (NGSEP_BAM_PART_RECORDS: test hook)
    if (const char* e = std::getenv("NGSEP_BAM_PART_RECORDS")) { const long v = std::atol(e); if (v >= 1) part_records = (size_t)v; }
    const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)host_threads(), n_rec / part_records + 1));
    std::vector<Part> parts((size_t)nt);
    static const char kSeq[] = "=ACMGRSVTWYHKDBN";
    static const int kOp[9] = {kOpM, kOpI, kOpD, kOpN, kOpS, kOpH, kOpP, kOpM, kOpX};   // BAM MIDNSHP=X -> the reference's codes
    struct SeqPairs { char c[256][2]; SeqPairs() { for (int v = 0; v < 256; v++) { c[v][0] = kSeq[v >> 4]; c[v][1] = kSeq[v & 15]; } } };
    static const SeqPairs pairs;
    on_threads(nt, [&](int t) {
        Part& part = parts[(size_t)t];
        OwnedBatch& b = part.b;
        const size_t r0 = n_rec * (size_t)t / (size_t)nt, r1 = n_rec * (size_t)(t + 1) / (size_t)nt;
        if (r1 > r0) {   // (the arrays are sized from the bytes of the range: an upper bound of everything they hold)
            const size_t bytes = (r1 < n_rec ? rec_at[r1] : d.size()) - rec_at[r0];
            b.seq_idx.reserve(r1 - r0); b.first.reserve(r1 - r0); b.flags.reserve(r1 - r0); b.mapq.reserve(r1 - r0); b.rg.reserve(r1 - r0); b.nh.reserve(r1 - r0);
            b.cigar_off.reserve(r1 - r0 + 1); b.base_off.reserve(r1 - r0 + 1); b.name_hash.reserve(r1 - r0);
            b.bases.reserve(bytes); b.quals.reserve(bytes);
        }
        std::vector<uint32_t> cg;
        for (size_t r = r0; r < r1; r++) {
            const size_t q = rec_at[r];
            const size_t block = u32(q - 4);
            const int ref_id = (int)u32(q), pos = (int)u32(q + 4) + 1;
            const size_t l_name = d[q + 8];
            const int mapq = d[q + 9];
            const size_t n_cig = u16(q + 12);
            const int flag = (int)u16(q + 14);
            const size_t l_seq = u32(q + 16);
            // every length of the fixed part against the record block (block <= the bytes left in the file, checked above)
            if (l_seq > block || 32 + l_name + 4 * n_cig + (l_seq + 1) / 2 + l_seq > block) { part.err = path + ": truncated BAM record"; return; }
            const size_t rec_end = q + block;
            size_t o = q + 32;
            const size_t name_at = o, name_len = l_name ? l_name - 1 : 0;
            o += l_name;
            cg.clear();
            bool bad = false;
            for (size_t k = 0; k < n_cig; k++, o += 4) { const uint32_t v = u32(o); if ((v & 15u) > 8u) bad = true; else cg.push_back(cig_make((int)(v >> 4), kOp[v & 15u])); }
            const size_t seq_at = o;
            o += (l_seq + 1) / 2;
            const size_t qual_at = o;
            o += l_seq;
            size_t rg_at = 0, rg_len = 0;
            bool has_rg = false;
            int nh = -1;
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
                bool aux_bad = need > rec_end - o;
                if (!aux_bad) switch (ty) {
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
                        if (t0 == 'R' && t1 == 'G' && ty == 'Z') { rg_at = o; rg_len = e - o; has_rg = true; }
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
                if (aux_bad) { part.err = path + ": truncated BAM record (optional fields)"; return; }
                if (is_int && t0 == 'N' && t1 == 'H' && ty != 'A') nh = (int)std::min(32000L, std::max(-1L, ival));
            }
            if (bad || ref_id < 0 || (size_t)ref_id >= n_ref) continue;   // unplaced record / unknown operation: skipped like in the SAM reader
            // ---- Loader::record, into the part -------------------------------------------------------------------------------
            b.seq_idx.push_back(ref_to_seq[(size_t)ref_id]);
            b.first.push_back(pos);
            b.flags.push_back((uint16_t)flag);
            b.mapq.push_back((uint8_t)std::min(255, mapq));
            int16_t rg = -1;
            if (has_rg) {   // index local to the part; the merge below turns it into the file-wide one
                std::string id((const char*)d.data() + rg_at, rg_len);
                auto ins = part.rg_index.emplace(id, (int)part.rg_ids.size());
                if (ins.second) part.rg_ids.push_back(id);
                rg = (int16_t)ins.first->second;
            }
            b.rg.push_back(rg);
            b.nh.push_back((int16_t)std::max(-1, std::min(32000, nh)));
            b.cigar.insert(b.cigar.end(), cg.begin(), cg.end());
            b.cigar_off.push_back((int64_t)b.cigar.size());
            if (l_seq) {
                const size_t at = b.bases.size();
                b.bases.resize(at + l_seq);
                b.quals.resize(at + l_seq);
                uint8_t* bs = b.bases.data() + at;
                uint8_t* qs = b.quals.data() + at;
                const unsigned char* sp = d.data() + seq_at;
                size_t k = 0;
                for (; k + 2 <= l_seq; k += 2) { const unsigned char by = sp[k >> 1]; bs[k] = (uint8_t)pairs.c[by][0]; bs[k + 1] = (uint8_t)pairs.c[by][1]; }
                if (k < l_seq) bs[k] = (uint8_t)pairs.c[sp[k >> 1]][0];
                const unsigned char* qp = d.data() + qual_at;
                if (qp[0] == 0xff) std::memset(qs, '+', l_seq);                                           // no qualities: RA:1025
                else for (size_t j = 0; j < l_seq; j++) qs[j] = (uint8_t)std::min(126, (int)qp[j] + 33);
            }
            b.base_off.push_back((int64_t)b.bases.size());
            b.name_hash.push_back(std::hash<std::string_view>()(std::string_view((const char*)d.data() + name_at, name_len)));
        }
    });
    return merge_parts(ld, parts);
}

bool load_alignments(const std::string& path, const Genome& g, OwnedBatch& b, std::string& err, RgTable* rgs = nullptr) {
    Loader ld(g, b, rgs);
    unsigned char magic[2] = {0, 0};
    { std::ifstream probe(path, std::ios::binary); if (!probe) { err = "cannot open " + path; return false; } probe.read((char*)magic, 2); }
    bool ok;
    const auto t_load = std::chrono::steady_clock::now();
    if (magic[0] == 0x1f && magic[1] == 0x8b) ok = load_bam(path, ld);   // BGZF (htsjdk picks the reader from the file content too)
    else {
        bool handled = false;
        ok = load_sam_parallel(path, ld, handled);
        if (!handled) { std::ifstream in(path); ok = load_sam_text(in, ld); }
    }
    if (!ok) err = ld.err;
    if (std::getenv("NGSEP_TIMING"))
        std::fprintf(stderr, "load_alignments %s: %.3f s, %zu records\n", path.c_str(), std::chrono::duration<double>(std::chrono::steady_clock::now() - t_load).count(), b.first.size());
    return ok;
}

// AlignmentsPileupGenerator.processFiles (APG:212-289): merge_order (model.cpp) gives the consumption order of the inputs' records.
void merge_inputs(const std::vector<OwnedBatch>& files, const AdmissionOptions& opt, OwnedBatch& out) {
    std::vector<BatchView> views;
    for (auto& f : files) views.push_back(f.view());
    std::vector<std::pair<int, int64_t>> order;
    merge_order(views, opt, order);
    // every record's place in the merged arrays, then the records copied on all host threads (a population run merges tens of millions)
    const size_t n = order.size(), rec0 = out.first.size();
    std::vector<int64_t> co(n + 1), bo(n + 1);
    co[0] = (int64_t)out.cigar.size(); bo[0] = (int64_t)out.bases.size();
    for (size_t i = 0; i < n; i++) {
        const OwnedBatch& f = files[(size_t)order[i].first];
        const size_t r = (size_t)order[i].second;
        co[i + 1] = co[i] + (f.cigar_off[r + 1] - f.cigar_off[r]);
        bo[i + 1] = bo[i] + (f.base_off[r + 1] - f.base_off[r]);
    }
    out.seq_idx.resize(rec0 + n); out.first.resize(rec0 + n); out.flags.resize(rec0 + n); out.mapq.resize(rec0 + n); out.rg.resize(rec0 + n); out.nh.resize(rec0 + n);
    out.cigar_off.resize(rec0 + n + 1); out.base_off.resize(rec0 + n + 1);
    out.cigar.resize((size_t)co[n]); out.bases.resize((size_t)bo[n]); out.quals.resize((size_t)bo[n]);
    const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)host_threads(), n / 50000 + 1));
    on_threads(nt, [&](int t) {
        for (size_t i = n * (size_t)t / (size_t)nt; i < n * (size_t)(t + 1) / (size_t)nt; i++) {
            const OwnedBatch& f = files[(size_t)order[i].first];
            const size_t r = (size_t)order[i].second, at = rec0 + i;
            out.seq_idx[at] = f.seq_idx[r]; out.first[at] = f.first[r]; out.flags[at] = f.flags[r]; out.mapq[at] = f.mapq[r]; out.rg[at] = f.rg[r]; out.nh[at] = f.nh[r];
            out.cigar_off[at + 1] = co[i + 1]; out.base_off[at + 1] = bo[i + 1];
            const size_t nc = (size_t)(co[i + 1] - co[i]), nb = (size_t)(bo[i + 1] - bo[i]);
            if (nc) std::memcpy(out.cigar.data() + co[i], f.cigar.data() + f.cigar_off[r], nc * sizeof(uint32_t));
            if (nb) {
                std::memcpy(out.bases.data() + bo[i], f.bases.data() + f.base_off[r], nb);
                std::memcpy(out.quals.data() + bo[i], f.quals.data() + f.base_off[r], nb);
            }
        }
    });
}

void usage() {
    std::fprintf(stderr, "usage: ngsep_b200 RelativeAlleleCountsCalculator -i <alignments.sam|.bam> -r <genome.fa> [-o <out.txt>] [-m minRD] [-M maxRD] [-q minBaseQS] [-s] [-frs regionsToFilter] [-srs regionsToSelect]\n");
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
        // RelativeAlleleCountsCalculator (main/CommandsDescriptor.xml:805-847: -i, -o, -r, -m, -M, -q, -s, -frs, -srs)
        std::string in, ref, out, frs, srs;
        int min_rd = 10, max_rd = 1000, min_bq = 20, device = 0;
        bool secondary = false;
        for (int i = 2; i < argc; i++) {
            std::string a = argv[i];
            auto val = [&]() -> const char* { if (i + 1 >= argc) { usage(); std::exit(1); } return argv[++i]; };
            if (a == "-i") in = val(); else if (a == "-r") ref = val(); else if (a == "-o") out = val();
            else if (a == "-m") min_rd = std::atoi(val()); else if (a == "-M") max_rd = std::atoi(val()); else if (a == "-q") min_bq = std::atoi(val());
            else if (a == "-s") secondary = true; else if (a == "--device") device = std::atoi(val());
            else if (a == "-frs") frs = val(); else if (a == "-srs") srs = val();
            else { std::fprintf(stderr, "Unrecognized option %s\n", a.c_str()); usage(); return 1; }
        }
        if (in.empty()) { std::fprintf(stderr, "The alignments input file is a required parameter\n"); return 2; }   // RelativeAlleleCountsCalculator.java:183
        if (ref.empty()) { std::fprintf(stderr, "The reference genome file is a required parameter\n"); return 2; }
        Genome g;
        if (!load_fasta(ref, g)) { std::fprintf(stderr, "cannot read %s\n", ref.c_str()); return 2; }
        OwnedBatch b;
        std::string err, text;
        if (!load_alignments(in, g, b, err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 2; }
        int rc = run_relative_allele_counts(g, b.view(), min_rd, max_rd, min_bq, secondary, device, text, err, frs, srs);
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
        if (files.size() > 1) {   // one input: processFile, no merge (APG:216-219)
            const auto t_merge = std::chrono::steady_clock::now();
            merge_inputs(files, ao, merged);
            b = &merged;
            if (std::getenv("NGSEP_TIMING"))
                std::fprintf(stderr, "merge of %zu inputs: %.3f s, %zu records\n", files.size(), std::chrono::duration<double>(std::chrono::steady_clock::now() - t_merge).count(), merged.first.size());
        }
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
    GpuSingleSampleVariantsDetector det(o);
    det.warm_up();   // the device contexts come up while the inputs are read
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
