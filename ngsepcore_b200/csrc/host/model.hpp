// Host side of the B200 pileup path (C++ above the C ABI; the reference's host language, Java, is not available
// in this image — see DESIGN.md).  This header: the read model used by admission, the indel-realigner pre-pass,
// the host indel caller and the tile packer.
//
// It is NOT the oracle and shares no code with oracle/: reads are kept as block tables over the caller's SoA
// buffers, events are processed sparsely (only positions where some read anchors an indel), and every mutation of
// a read is recorded as a VERSION with a validity window so the device can evaluate any position against the read
// state the reference would have seen there (hard part 1).
//
// Reference citations: RA = src/ngsep/alignments/ReadAlignment.java, APG = src/ngsep/discovery/
// AlignmentsPileupGenerator.java, IRPL = src/ngsep/discovery/IndelRealignerPileupListener.java.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace ngsep_b200 {
namespace host {

constexpr int kOpH = 0, kOpD = 1, kOpI = 2, kOpM = 3, kOpP = 4, kOpN = 5, kOpS = 6, kOpX = 7;
inline int cig_len(uint32_t v) { return (int)(v >> 3); }
inline int cig_op(uint32_t v) { return (int)(v & 7u); }
inline bool cig_ref(uint32_t v) { return (v & 1u) != 0; }
inline bool cig_read(uint32_t v) { return (v & 2u) != 0; }
inline bool cig_both(uint32_t v) { return (v & 3u) == 3u; }
inline bool op_is_indel(int op) { return op == kOpD || op == kOpI; }
inline uint32_t cig_make(int len, int op) { return (uint32_t)(len * 8 + op); }

// Decoded alignment records, struct-of-arrays (what htsjdk / the SAM reader hands over).
struct BatchView {
    int64_t n = 0;
    const int32_t* seq_idx = nullptr;
    const int32_t* first = nullptr;
    const uint16_t* flags = nullptr;
    const uint8_t* mapq = nullptr;
    const int16_t* rg = nullptr;        // -1 = no RG tag
    const int16_t* nh = nullptr;        // -1 = no NH tag
    const int64_t* cigar_off = nullptr; // [n+1]
    const uint32_t* cigar = nullptr;
    const int64_t* base_off = nullptr;  // [n+1]
    const uint8_t* bases = nullptr;     // ASCII; nullptr-equivalent "*" is a zero-length read
    const uint8_t* quals = nullptr;     // phred+33; may be nullptr (no qualities: '+' is used, RA:1025)
    const uint64_t* name_hash = nullptr;// optional, for the duplicate-record rule (RAFR:292-306)
};

struct Genome {
    std::vector<std::string> names;
    std::vector<std::string> seqs;      // as in the FASTA (case kept); chars outside AaCcNngGtT read as 'N'
    bool get(int seq, int first, int last, std::string& out) const;   // 1-based inclusive, false when out of range
};

// An indel anchored at reference position `pos` (last reference base before the event), RA:798-815.
struct Anchor { int pos, last, len; bool ins; int read_off; };

struct ReadState {                 // everything the realigner may change (IRPL:274-313, 400-526)
    int first = 0, last = 0;
    int ign_s = 0, ign_e = 0;
    std::vector<uint32_t> cigar;
    bool operator==(const ReadState& o) const { return first == o.first && last == o.last && ign_s == o.ign_s && ign_e == o.ign_e && cigar == o.cigar; }
};

struct Version { int from, to; ReadState st; };   // st describes the read for reference positions from <= p < to

struct HostRead {
    int64_t rec = -1;              // index into the batch
    int admit_first = 0;           // start at admission: the ordering key of the pileup (APG:416-431)
    int L = 0;
    bool neg = false;
    bool unique = true;            // ReadAlignment.isUnique(): not flagged MULTIPLE_ALN by the reader (RAFR:229-236, 284-291)
    int rg = -1;
    const uint8_t* bases = nullptr;
    const uint8_t* quals = nullptr;
    ReadState st;                  // current state
    int cur_from = INT32_MIN;      // current state is valid from this position on
    std::vector<Version> closed;   // earlier states
    // derived from st (recomputed after every mutation)
    std::vector<Anchor> anchors;
    bool derived = false;

    void derive();                                   // anchors per RA:747-834
    int read_offset(int pos) const;                  // RA:842-871
    int call_length(int pos, int* off = nullptr) const;  // alleleCallLength at the read base aligned to pos (0 = none)
    const Anchor* anchor_at(int pos) const;
    bool has_anchor_in(int lo, int hi) const;        // RA:1471-1478
    bool span_call(int pos, int last_pos, std::string& alleles, std::string& quals) const;  // RA:1008-1016,1034-1041 after the PR:137 gate
    char base_at(int off) const { return (char)bases[off]; }
    int soft_clip_start() const { return cig_op(st.cigar.front()) == kOpS ? cig_len(st.cigar.front()) : 0; }
    int soft_clip_end() const { return cig_op(st.cigar.back()) == kOpS ? cig_len(st.cigar.back()) : 0; }
    // mutations (each invalidates `derived`)
    bool move_indel(int anchor_pos, int new_pos);    // RA:1114-1153
    void realign_start(int new_first, int first_match, int ref_after, int read_after);   // RA:1372-1418
    void realign_end(int ref_before, int read_before, int final_ref_start, int final_match);  // RA:1427-1469
};

struct AdmissionOptions {
    int min_mq = 20, max_alns_per_start = 5;
    bool secondary = false, non_unique_primary = false;
    int ignore5 = 0, ignore3 = 0;
    int query_first = INT32_MIN;   // records ending before it are skipped before the start groups form (APG:316)
};

// Applies the record filter of RAFR:219-354 (flag set from APG:363-375), the duplicate-record rule and the per-read-
// group same-start cap of APG:407-433 to the records of one sequence; returns the admitted reads in pileup order.
std::vector<HostRead> admit_sequence(const BatchView& b, int64_t rec_lo, int64_t rec_hi, const AdmissionOptions& opt, int threads = 1);

// AlignmentsPileupGenerator.processFiles / chooseNextAln (APG:212-289): the order in which the records of several
// coordinate-sorted inputs are consumed.  Each input is first reduced to the records its own reader hands over (RAFR:219-354
// with the flag filter of APG:363-375, the duplicate-record rule per input); the next record is then the smallest head by
// (sequence index, first, last) — genome/GenomicRegionComparator.java:51-67 — ties going to the lowest input index.
// out: (input, record) pairs in consumption order.
void merge_order(const std::vector<BatchView>& inputs, const AdmissionOptions& opt, std::vector<std::pair<int, int64_t>>& out);

// The admitted reads of a sequence as the detectors keep them: 16 bytes per read.  Almost every read is one match block that no
// event of the realigner pre-pass ever touches; such a read needs no HostRead at all (the packer takes it straight from the
// batch).  Reads with any other CIGAR, and the reads under an event, are materialised (SeqPlan::heavy).
struct LiteRead {
    int32_t rec = 0;               // record index relative to the first record of the sequence in the batch
    int32_t first = 0, last = 0;   // reference span at admission; first is the ordering key of the pileup
    int32_t heavy = -1;            // >= 0: index of the read's HostRead; kLiteSimple: one match block, untouched; kLiteComplex: any other CIGAR, not materialised yet
};
constexpr int32_t kLiteSimple = -1, kLiteComplex = -2;
// admit_sequence's decisions (same filters, same order) without building the reads
std::vector<LiteRead> admit_sequence_lite(const BatchView& b, int64_t rec_lo, int64_t rec_hi, const AdmissionOptions& opt, int threads = 1);
// the HostRead admit_sequence builds for record rec
HostRead make_read(const BatchView& b, int64_t rec, const AdmissionOptions& opt);

inline int dna_index(char c) {
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; }
    return -1;
}
inline char up(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }
inline std::string upper(std::string s) { for (auto& c : s) c = up(c); return s; }

}  // namespace host
}  // namespace ngsep_b200
