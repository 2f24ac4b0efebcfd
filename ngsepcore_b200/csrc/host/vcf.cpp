// VCF text of the detector, byte-compatible with the reference's writer:
//   vcf/VCFFileHeader.java:45-71,219-246   header
//   vcf/VCFFileWriter.java:44-308          record, INFO and genotype fields
//   discovery/SingleSampleVariantsDetector.java:944-963   format choice (BSDP for SNV-like calls, ADP otherwise), FS / TYPE
//   variants/CalledSNV.java:134-160,259-266 and variants/CalledGenomicVariantImpl.java:228-286   float PL, copy numbers
#include <cmath>
#include <cstdio>
#include <sstream>

#include "pipeline.hpp"

namespace ngsep_b200 {
namespace host {

void host_tables(double het_rate, double* t) {
    for (int i = 0; i < NGSEP_N_TABLES; i++) t[i] = 0;
    for (int q = 3; q <= 30; q++) {
        double e = std::pow(10.0, -0.1 * q), s = 1 - e;
        t[3 * q] = std::log10(s);
        t[3 * q + 1] = -0.1 * q - std::log10(3.0);
        t[3 * q + 2] = std::log10(0.5 * s + (1 - 0.5) * e / 3);
    }
    t[93] = std::log10((1 - het_rate) / 4);
    t[94] = std::log10(het_rate / 12);
}

static std::string header_lines() {
    static const char* lines[][5] = {
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
    std::ostringstream o;
    o << "##fileformat=VCFv4.2\n";
    for (auto& l : lines) o << "##" << l[0] << "=<ID=" << l[1] << ",Number=" << l[2] << ",Type=" << l[3] << ",Description=\"" << l[4] << "\">\n";
    return o.str();
}

std::string GpuSingleSampleVariantsDetector::vcf_header(const std::string& sample_id, int ploidy, bool psp) {
    std::ostringstream o;
    o << header_lines();
    if (psp) o << "##SAMPLE=<ID=" << sample_id << ",PL=" << ploidy << ">\n";
    o << "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" << sample_id << "\n";
    return o.str();
}

// MSVD:453-455: every sample is added to the header (samplesWithHeaderLine is a TreeMap; the samples already come in id order)
std::string GpuMultisampleVariantsDetector::vcf_header(const std::vector<SampleDef>& samples, bool psp) {
    std::ostringstream o;
    o << header_lines();
    if (psp) for (auto& s : samples) o << "##SAMPLE=<ID=" << s.id << ",PL=" << s.ploidy << ">\n";
    o << "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO";
    if (!samples.empty()) o << "\tFORMAT";
    for (auto& s : samples) o << "\t" << s.id;
    o << "\n";
    return o.str();
}

void append_gt(std::string& s, const std::vector<int>& called, int ploidy) {  // VCFFileWriter.java:171-199
    if (called.empty()) { s += "."; if (ploidy > 1) s += "/."; }
    else if (called.size() == 1) { s += std::to_string(called[0]); if (ploidy > 1) { s += "/"; s += std::to_string(called[0]); } }
    else for (size_t i = 0; i < called.size(); i++) { if (i) s += "/"; s += std::to_string(called[i]); }
}

std::string snv_vcf_line(const ngsep_snv_call& c, const std::string& chrom, int ploidy, bool csb, bool embedded) {
    static const char B[] = "ACGT";
    const int ref = c.ref;
    auto L = [&](int i, int j) { return c.logcond[i * 4 + j]; };
    std::string s = chrom;
    s += "\t"; s += std::to_string(c.pos); s += "\t.\t"; s += B[ref]; s += "\t";
    std::vector<int> called;
    std::vector<long> pl;
    std::string acn, info;
    if (c.kind == NGSEP_CALL_TRIALLELIC) {
        int al[3] = {ref, c.a0, c.a1};
        s += B[c.a0]; s += ","; s += B[c.a1];
        for (int j = 0; j < 3; j++) for (int i = 0; i <= j; i++) pl.push_back(jround(-10 * L(al[i], al[j])));
        called = {1, 2};
        acn = "0,1,1";                      // CalledGenomicVariantImpl.java:238-241 (ploidy 2, two called alleles)
        if (ploidy == 1) acn = "0,1,1";
        info = "TYPE=MULTISNV";
    } else {
        int alt = c.a1;
        s += B[alt];
        auto f = [](double v) { return (double)(float)v; };  // CalledSNV keeps floats, CalledSNV.java:259-266
        pl = {jround(-10 * f(L(ref, ref))), jround(-10 * f(L(ref, alt))), jround(-10 * f(L(alt, alt)))};
        if (c.kind == NGSEP_CALL_HET) { called = {0, 1}; acn = "1,1"; }  // CalledSNV.java:143-147
        else { called = {1}; acn = "0," + std::to_string(ploidy); }       // CalledSNV.java:140-142
        if (csb) {  // CountsHelper.getScoreStrandBiasFisher(indexRef, indexAlt), CH:563-576
            double p = fisher_pvalue(c.strand[ref * 2], c.strand[alt * 2], c.strand[ref * 2 + 1], c.strand[alt * 2 + 1]);
            info = "FS=" + std::to_string(std::min(100, phred(p)));
        }
        if (embedded) { if (!info.empty()) info += ";"; info += "TYPE=EMBEDDED"; }
    }
    s += "\t"; s += std::to_string(c.qual); s += "\t.\t"; s += info.empty() ? "." : info;
    s += "\tGT:PL:GQ:DP:BSDP:ACN\t";
    append_gt(s, called, ploidy);
    s += ":";
    for (size_t i = 0; i < pl.size(); i++) { if (i) s += ","; s += std::to_string(pl[i]); }
    s += ":"; s += std::to_string(c.gq);
    s += ":"; s += std::to_string(c.depth);
    s += ":";
    for (int i = 0; i < 4; i++) { if (i) s += ","; s += std::to_string(c.count[i]); }
    s += ":"; s += acn;
    s += "\n";
    return s;
}

std::string indel_vcf_line(const IndelRecord& r, const std::string& chrom, int ploidy) {
    std::string s = chrom;
    s += "\t"; s += std::to_string(r.pos); s += "\t.\t"; s += r.alleles[0]; s += "\t";
    for (size_t i = 1; i < r.alleles.size(); i++) { if (i > 1) s += ","; s += r.alleles[i]; }
    s += "\t"; s += std::to_string(r.qual); s += "\t.\t"; s += r.pool ? "." : (r.str ? "TYPE=STR" : "TYPE=INDEL");   // (a pool variant is a plain GenomicVariantImpl, SSVPL:336)
    s += "\tGT:PL:GQ:DP:ADP:ACN\t";
    append_gt(s, r.called, ploidy);
    // VariantCallReport resolves alleles through a TreeMap name->index: with duplicate names the last index wins
    auto index_of = [&](const std::string& a) { int k = -1; for (size_t i = 0; i < r.alleles.size(); i++) if (r.alleles[i] == a) k = (int)i; return k; };
    s += ":";
    size_t n = r.alleles.size();
    bool firstv = true;
    for (size_t j = 0; j < n; j++) for (size_t i = 0; i <= j; i++) {
        if (!firstv) s += ",";
        firstv = false;
        s += std::to_string(jround(-10 * r.logc[index_of(r.alleles[i])][index_of(r.alleles[j])]));
    }
    s += ":"; s += std::to_string(r.gq);
    s += ":"; s += std::to_string(r.depth);
    s += ":";
    for (size_t i = 0; i < n; i++) { if (i) s += ","; s += std::to_string(r.counts[index_of(r.alleles[i])]); }
    // copy numbers, CalledGenomicVariantImpl.updateAllelesCopyNumberFromCounts (:228-286)
    std::vector<int> cn(n, 0);
    int nc = (int)r.called.size();
    if (nc == 1 && r.called[0] == 0) cn[0] = ploidy;
    else if (ploidy <= nc) for (int k : r.called) cn[k] = 1;
    else {
        std::vector<int> rc(nc);
        int tot = 0;
        for (int k = 0; k < nc; k++) { rc[k] = r.counts[index_of(r.alleles[r.called[k]])]; if (rc[k] == 0) rc[k] = 1; tot += rc[k]; }
        int sum = 0;
        for (int k = 0; k < nc; k++) { cn[r.called[k]] = (int)std::max(1L, jround((double)ploidy * rc[k] / tot)); sum += cn[r.called[k]]; }
        if (sum < ploidy) cn[r.called[0]] += ploidy - sum;
        else { int ex = sum - ploidy; for (int k = nc - 1; ex > 0 && k >= 0; k--) { int rm = std::min(ex, cn[r.called[k]] - 1); cn[r.called[k]] -= rm; ex -= rm; } }
    }
    s += ":";
    if (ploidy == 0) s += ".";
    else for (size_t i = 0; i < n; i++) { if (i) s += ","; s += std::to_string(cn[i]); }
    s += "\n";
    return s;
}

}  // namespace host
}  // namespace ngsep_b200
