// ngsep_b200 — command line with the option ids of the reference modules (main/CommandsDescriptor.xml:480-703):
//   ngsep_b200 SingleSampleVariantsDetector -i in.sam -r ref.fa -o prefix [options]
//   ngsep_b200 MultisampleVariantsDetector -r ref.fa -o out.vcf [options] a.sam b.sam ...      (MSVD.main, MSVD:412-419)
// Reads SAM text (the reference reads SAM/BAM/CRAM through htsjdk, RAFR:171-183; BAM decoding is a "next" row),
// writes <prefix>.vcf like SSVD:926.  Errors follow the reference: a missing -i / -r is reported as in SSVD:595-596,
// an unknown option prints usage and exits 1 (CommandsDescriptor.java:441-445).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <unordered_map>

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

// rgs (optional): the @RG table shared by every input of a multi-sample run (ids in header order, RAFR:184-197); read
// groups that only appear on records get an index too (they belong to no sample).
bool load_sam(const std::string& path, const Genome& g, OwnedBatch& b, std::string& err, RgTable* rgs = nullptr) {
    std::ifstream in(path);
    if (!in) { err = "cannot open " + path; return false; }
    std::unordered_map<std::string, int> seq_of, rg_local;
    std::unordered_map<std::string, int>& rg_of = rgs ? rgs->index : rg_local;
    for (size_t i = 0; i < g.names.size(); i++) seq_of[g.names[i]] = (int)i;
    std::string line;
    std::vector<std::string> f;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        if (line[0] == '@') {
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
            continue;
        }
        f.clear();
        size_t s = 0;
        while (true) { size_t e = line.find('\t', s); f.push_back(line.substr(s, e == std::string::npos ? std::string::npos : e - s)); if (e == std::string::npos) break; s = e + 1; }
        if (f.size() < 11) continue;
        auto it = seq_of.find(f[2]);
        int flag = std::atoi(f[1].c_str());
        if (it == seq_of.end()) continue;  // unmapped / unknown reference
        std::vector<uint32_t> cg;
        if (f[5] != "*") {
            int n = 0; bool bad = false;
            for (char c : f[5]) {
                if (c >= '0' && c <= '9') n = n * 10 + (c - '0');
                else { const char* codes = "HDIMPNSX"; const char* p = std::strchr(codes, c); int op = p ? (int)(p - codes) : (c == '=' ? 3 : -1); if (op < 0) { bad = true; break; } cg.push_back(cig_make(n, op)); n = 0; }
            }
            if (bad) continue;
        }
        b.seq_idx.push_back(it->second);
        b.first.push_back(std::atoi(f[3].c_str()));
        b.flags.push_back((uint16_t)flag);
        b.mapq.push_back((uint8_t)std::min(255, std::atoi(f[4].c_str())));
        int16_t rg = -1, nh = -1;
        for (size_t k = 11; k < f.size(); k++) {
            if (!f[k].compare(0, 5, "RG:Z:")) {
                auto r = rg_of.emplace(f[k].substr(5), (int)rg_of.size());
                if (rgs && r.second) { rgs->ids.push_back(f[k].substr(5)); rgs->samples.push_back(std::string(1, '\0')); }   // on records only: no sample
                rg = (int16_t)r.first->second;
            }
            else if (!f[k].compare(0, 5, "NH:i:")) nh = (int16_t)std::min(32000, std::atoi(f[k].c_str() + 5));
        }
        b.rg.push_back(rg); b.nh.push_back(nh);
        b.cigar.insert(b.cigar.end(), cg.begin(), cg.end());
        b.cigar_off.push_back((int64_t)b.cigar.size());
        if (f[9] != "*") {
            b.bases.insert(b.bases.end(), f[9].begin(), f[9].end());
            size_t L = f[9].size();
            if (f[10] == "*") b.quals.insert(b.quals.end(), L, (uint8_t)'+');  // RA:1025
            else { for (size_t k = 0; k < L; k++) b.quals.push_back(k < f[10].size() ? (uint8_t)std::min(127, (int)(unsigned char)f[10][k]) : (uint8_t)38); }  // RA:581-595
        }
        b.base_off.push_back((int64_t)b.bases.size());
        b.name_hash.push_back(std::hash<std::string>()(f[0]));
    }
    return true;
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
    std::fprintf(stderr, "usage: ngsep_b200 MultisampleVariantsDetector -r <genome.fa> -o <out.vcf> [options] <alignments.sam>*\n"
                         "       ngsep_b200 SingleSampleVariantsDetector -i <alignments.sam> -r <genome.fa> -o <outputPrefix>\n"
                         "  [-sampleId S] [-ploidy N] [-psp] [-minMQ N] [-querySeq S -first N -last N] [-ignoreLowerCaseRef]\n"
                         "  [-maxAlnsPerStartPos N] [-p] [-s] [-ignore5 N] [-ignore3 N] [-h X] [-maxBaseQS N] [-minQuality N]\n"
                         "  [-embeddedSNVs] [-csb] [--device N] [--tile-len N]\n");
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2 || (std::strcmp(argv[1], "SingleSampleVariantsDetector") && std::strcmp(argv[1], "MultisampleVariantsDetector"))) { usage(); return 1; }
    const bool multi = !std::strcmp(argv[1], "MultisampleVariantsDetector");
    DetectorOptions o;
    std::string in, ref, out;
    std::vector<std::string> inputs;
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
        else if (a == "--device") o.device = std::atoi(val()); else if (a == "--tile-len") o.tile_len = std::atoi(val());
        else { std::fprintf(stderr, "Unrecognized option %s\n", a.c_str()); usage(); return 1; }
    }
    if (multi) {
        if (ref.empty()) { std::fprintf(stderr, "The reference genome file is a required parameter\n"); return 2; }   // MSVD:423
        Genome g;
        if (!load_fasta(ref, g)) { std::fprintf(stderr, "cannot read %s\n", ref.c_str()); return 2; }
        RgTable rgs;
        std::vector<OwnedBatch> files(inputs.size());
        std::string err;
        for (size_t f = 0; f < inputs.size(); f++) if (!load_sam(inputs[f], g, files[f], err, &rgs)) { std::fprintf(stderr, "%s\n", err.c_str()); return 2; }
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
    if (!load_sam(in, g, b, err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 2; }
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
