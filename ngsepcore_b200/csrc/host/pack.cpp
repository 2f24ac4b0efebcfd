// Tile packer: host reads (with their realigner versions) -> the header / payload layout of include/ngsep_pileup.h.
// Contiguous ranges of reads are packed by independent threads into per-thread parts (kept between calls: no allocation or
// page faults after the first tile), then copied side by side into the tile's buffers with their offsets shifted.
#include <algorithm>
#include <array>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <pthread.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "pipeline.hpp"

namespace ngsep_b200 {
namespace host {

// ---- HostBuf ------------------------------------------------------------------------------------------------------------
template <typename T>
void HostBuf<T>::release() {
    if (p) {
        if (registered) ngsep_host_unregister(p);
        std::free(p);
    }
    p = nullptr; n = cap = 0; registered = false;
}
template <typename T>
void HostBuf<T>::reserve(size_t want, bool exact) {
    if (want <= cap && (registered || !pin || cap == 0)) return;
    if (want <= cap && pin && !registered) {   // asked to pin after the allocation
        registered = ngsep_host_register(p, cap * sizeof(T)) == NGSEP_OK;
        return;
    }
    release();
    size_t c = std::max<size_t>(exact ? want : want + want / 4, 1024);
    size_t bytes = (c * sizeof(T) + 4095) & ~(size_t)4095;
    p = (T*)std::aligned_alloc(4096, bytes);
    if (!p) throw std::bad_alloc();
    cap = bytes / sizeof(T);
    if (pin) registered = ngsep_host_register(p, bytes) == NGSEP_OK;
}
template struct HostBuf<ngsep_read_hdr>;
template struct HostBuf<uint32_t>;
template struct HostBuf<uint16_t>;
template struct HostBuf<uint8_t>;

bool wire_compact() {
    const char* e = std::getenv("NGSEP_WIRE_COMPACT");
    return !(e && std::atoi(e) == 0);
}

int host_threads() {
    if (const char* e = std::getenv("NGSEP_HOST_THREADS")) { int v = std::atoi(e); if (v >= 1 && v <= 256) return v; }
    unsigned hw = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(32u, hw ? hw : 1u));
}

namespace {

// f(0) .. f(n - 1) on the calling thread and up to n - 1 helpers.  The helpers are threads that live as long as the process (a tile
// is packed in three short parallel steps: starting and joining sixteen threads for each of them cost as much as a sixth of the
// packing itself); a second caller at the same time, or a child process after fork(), starts its own threads as before.
class Helpers {
public:
    void run(int n, const std::function<void(int)>& f) {
        if (n <= 1) { if (n == 1) f(0); return; }
        std::unique_lock<std::mutex> own(owner_, std::try_to_lock);
        if (!own.owns_lock() || dead_.load()) { spawn(n, f); return; }
        grow(n - 1);
        {
            std::lock_guard<std::mutex> g(m_);
            job_ = &f; n_ = n; next_.store(1); left_ = n - 1; gen_++;
        }
        cv_.notify_all();
        f(0);
        for (;;) { const int k = next_.fetch_add(1); if (k >= n) break; f(k); finish_one(); }
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [&] { return left_ == 0 && in_job_ == 0; });   // (no helper still holds the job: f dies with this call)
        job_ = nullptr;
    }
    static Helpers& get() { static Helpers* h = new Helpers(); return *h; }   // (never destroyed: its threads may outlive static destructors)
private:
    Helpers() { pthread_atfork(nullptr, nullptr, [] { get().dead_.store(true); }); }
    static void spawn(int n, const std::function<void(int)>& f) {
        std::vector<std::thread> th;
        for (int t = 1; t < n; t++) th.emplace_back(f, t);
        f(0);
        for (auto& x : th) x.join();
    }
    void grow(int want) {
        while ((int)threads_.size() < want) { threads_.emplace_back([this] { loop(); }); threads_.back().detach(); }
    }
    void finish_one() { std::lock_guard<std::mutex> g(m_); if (--left_ == 0 && in_job_ == 0) done_.notify_all(); }
    void loop() {
        unsigned long seen = 0;
        for (;;) {
            const std::function<void(int)>* job;
            int n;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_; job = job_; n = n_;
                if (job) in_job_++;
            }
            if (!job) continue;
            for (;;) { const int k = next_.fetch_add(1); if (k >= n) break; (*job)(k); finish_one(); }
            { std::lock_guard<std::mutex> g(m_); if (--in_job_ == 0 && left_ == 0) done_.notify_all(); }
        }
    }
    std::mutex owner_, m_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> threads_;
    const std::function<void(int)>* job_ = nullptr;
    int n_ = 0, left_ = 0, in_job_ = 0;
    unsigned long gen_ = 0;
    std::atomic<int> next_{0};
    std::atomic<bool> dead_{false};
};

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

// ---- eight bases / qualities at a time ------------------------------------------------------------------------------------
constexpr uint64_t k01 = 0x0101010101010101ull, k7f = 0x7f7f7f7f7f7f7f7full, k80 = 0x8080808080808080ull;
inline uint64_t zero_bytes(uint64_t v) { return ~(((v & k7f) + k7f) | v | k7f); }            // 0x80 in every byte of v that is zero (exact)
inline uint32_t gather_high_bits(uint64_t m) { return (uint32_t)((((m >> 7) & k01) * 0x0102040810204080ull) >> 56); }   // bit k = byte k

// 2-bit codes (A0 C1 G2 T3, either case) of eight ASCII bases; *bad = one bit per byte that is not a base of ACGTacgt
inline uint32_t encode8(uint64_t w, uint32_t* bad) {
    const uint64_t u = w & 0xDFDFDFDFDFDFDFDFull;   // upper case
    const uint64_t ok = zero_bytes(u ^ (0x41 * k01)) | zero_bytes(u ^ (0x43 * k01)) | zero_bytes(u ^ (0x47 * k01)) | zero_bytes(u ^ (0x54 * k01));
    *bad = gather_high_bits(~ok & k80);
    uint64_t x = (w >> 1) & (3 * k01);          // A 0, C 1, G 3, T 2
    x ^= (x >> 1) & k01;                        // A 0, C 1, G 2, T 3
    x = (x | (x >> 6)) & 0x000F000F000F000Full;
    x = (x | (x >> 12)) & 0x000000FF000000FFull;
    x = (x | (x >> 24)) & 0xFFFFull;
    return (uint32_t)x;
}
// one bit per byte with quality char - 33 <= 3 (CountsHelper.java:214: such a base is never counted)
inline uint32_t lowq8(uint64_t w) {
    const uint64_t ge37 = (((w & k7f) + (0x7f - 36) * k01) | w) & k80;
    return gather_high_bits(~ge37 & k80);
}
inline uint64_t load8(const uint8_t* p, int n, uint8_t fill) {   // n in 1..8 bytes, the rest = fill
    uint64_t w;
    if (n >= 8) { std::memcpy(&w, p, 8); return w; }
    uint8_t t[8];
    std::memset(t, fill, 8);
    std::memcpy(t, p, (size_t)n);
    std::memcpy(&w, t, 8);
    return w;
}

// How the tile's qualities travel (ngsep_tile.qual_bits): the bytes themselves, or 2- / 4-bit codes when the tile's qualities
// fall into at most 4 / 16 of the classes the path can tell apart — every q <= 3 alike (never counted, CH:214), min(30, q)
// otherwise (CH:91) — with one representative byte per class in the dictionary.
struct QualCoder {
    int bits = 8;
    uint8_t code_of[256];
    uint8_t dict[16];
    alignas(32) uint8_t class_code[32];   // the code of quality class k (qual_class), for the 32-at-a-time coder
    alignas(32) uint8_t class_known[32];  // 0xff: class k has a code (a coder taken over from an earlier tile may meet classes it has none for)
};
inline int qual_class(int c) { const int q = c - 33; return q <= 3 ? 0 : std::min(30, q); }
QualCoder choose_coder(const bool* seen, int min_bits = 0) {
    QualCoder qc;
    std::memset(qc.code_of, 0, sizeof qc.code_of);
    std::memset(qc.dict, 0, sizeof qc.dict);
    std::memset(qc.class_code, 0, sizeof qc.class_code);
    std::memset(qc.class_known, 0xff, sizeof qc.class_known);
    bool present[31] = {false};
    for (int c = 0; c < 256; c++) if (seen[c]) present[qual_class(c)] = true;
    int n = 0, code_of_class[31];
    for (int k = 0; k < 31; k++) { code_of_class[k] = -1; if (present[k]) code_of_class[k] = n++; }
    int bits = n <= 4 ? 2 : (n <= 16 ? 4 : 8);
    if (const char* e = std::getenv("NGSEP_QUAL_BITS")) { const int v = std::atoi(e); if (v == 8 || (v == 4 && bits == 2)) bits = v; }   // experiments: wider than needed
    if (bits < min_bits) bits = min_bits;   // a caller that needs the quality bytes themselves
    qc.bits = bits;
    if (bits == 8) return qc;
    for (int k = 0; k < 31; k++) if (code_of_class[k] >= 0) qc.dict[code_of_class[k]] = (uint8_t)(k == 0 ? 35 : 33 + k);
    for (int c = 0; c < 256; c++) { const int k = code_of_class[qual_class(c)]; qc.code_of[c] = (uint8_t)(k < 0 ? 0 : k); }
    for (int k = 0; k < 32; k++) { qc.class_code[k] = (uint8_t)((k < 31 && code_of_class[k] >= 0) ? code_of_class[k] : 0); qc.class_known[k] = (k < 31 && code_of_class[k] >= 0) ? 0xff : 0; }
    return qc;
}
// the coder an earlier tile used, from its dictionary (code c stands for the class of dict[c])
QualCoder coder_from_dict(int bits, const uint8_t* dict) {
    bool seen[256] = {false};
    for (int c = 0; c < (1 << bits) && c < 16; c++) if (dict[c]) seen[dict[c]] = true;
    QualCoder qc = choose_coder(seen, bits);
    return qc;
}
inline size_t qual_run_bytes(int L, int bits) { return ((((size_t)L * (size_t)bits + 7) / 8) + 3) & ~(size_t)3; }

// ---- 32 bases / qualities at a time (AVX2, chosen at run time; the 8-at-a-time code above is the portable path) ------------
#if defined(__x86_64__)
#define NGSEP_HAVE_AVX2_PATH 1
inline bool have_avx2() {
    static const bool on = [] { const char* e = std::getenv("NGSEP_NO_AVX2"); return !(e && std::atoi(e) != 0) && __builtin_cpu_supports("avx2"); }();
    return on;
}
// 32 two-bit values (one per byte of x, each 0..3) -> 64 bits, value k at bits 2k
__attribute__((target("avx2"))) inline uint64_t squeeze2_avx2(__m256i x) {
    const __m256i y = _mm256_maddubs_epi16(x, _mm256_set1_epi16(0x0401));        // x0 + 4 x1 per 16-bit element
    const __m256i z = _mm256_madd_epi16(y, _mm256_set1_epi32(0x00100001));       // y0 + 16 y1 per 32-bit element: one byte of output
    const __m256i g = _mm256_shuffle_epi8(z, _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                                             0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1));
    return (uint64_t)(uint32_t)_mm256_extract_epi32(g, 0) | ((uint64_t)(uint32_t)_mm256_extract_epi32(g, 4) << 32);
}
// the quality class (qual_class) of 32 quality characters
__attribute__((target("avx2"))) inline __m256i qual_class_avx2(__m256i c) {
    __m256i t = _mm256_min_epu8(_mm256_subs_epu8(c, _mm256_set1_epi8(33)), _mm256_set1_epi8(30));
    return _mm256_and_si256(t, _mm256_cmpgt_epi8(t, _mm256_set1_epi8(3)));
}
// bases2 / N mask / low-quality mask of one read (encode_read's arrays, already zeroed)
__attribute__((target("avx2"))) void encode_bits_avx2(const uint8_t* bases, const uint8_t* quals, int len, uint32_t* b2, uint32_t* nm, uint32_t* lq, uint32_t* any_bad) {
    const __m256i up = _mm256_set1_epi8((char)0xDF);
    for (int i = 0; i < len; i += 32) {
        const int n = std::min(32, len - i);
        __m256i v, q;
        if (n == 32) { v = _mm256_loadu_si256((const __m256i*)(bases + i)); q = _mm256_loadu_si256((const __m256i*)(quals + i)); }
        else {
            alignas(32) uint8_t tb[32], tq[32];
            std::memset(tb, 'A', 32); std::memset(tq, 0xff, 32);
            std::memcpy(tb, bases + i, (size_t)n); std::memcpy(tq, quals + i, (size_t)n);
            v = _mm256_load_si256((const __m256i*)tb); q = _mm256_load_si256((const __m256i*)tq);
        }
        const __m256i u = _mm256_and_si256(v, up);
        const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(u, _mm256_set1_epi8('A')), _mm256_cmpeq_epi8(u, _mm256_set1_epi8('C'))),
                                           _mm256_or_si256(_mm256_cmpeq_epi8(u, _mm256_set1_epi8('G')), _mm256_cmpeq_epi8(u, _mm256_set1_epi8('T'))));
        const uint32_t bad = ~(uint32_t)_mm256_movemask_epi8(ok);
        __m256i x = _mm256_and_si256(_mm256_srli_epi16(v, 1), _mm256_set1_epi8(3));          // A 0, C 1, G 3, T 2
        x = _mm256_xor_si256(x, _mm256_and_si256(_mm256_srli_epi16(x, 1), _mm256_set1_epi8(1)));   // A 0, C 1, G 2, T 3
        x = _mm256_and_si256(x, ok);                                                           // anything else: code 0, bit in the N mask
        const uint64_t c = squeeze2_avx2(x);
        b2[i >> 4] = (uint32_t)c;
        if (n > 16) b2[(i >> 4) + 1] = (uint32_t)(c >> 32);
        if (bad) { *any_bad |= bad; nm[i >> 5] = bad; }
        const __m256i ge37 = _mm256_cmpeq_epi8(_mm256_max_epu8(q, _mm256_set1_epi8(37)), q);
        lq[i >> 5] = ~(uint32_t)_mm256_movemask_epi8(ge37);
    }
}
// which quality classes occur in q[0, n): bit k of the result = class k
__attribute__((target("avx2"))) uint32_t scan_classes_avx2(const uint8_t* q, size_t n) {
    const __m256i bit_lut = _mm256_setr_epi8(1, 2, 4, 8, 16, 32, 64, (char)128, 1, 2, 4, 8, 16, 32, 64, (char)128, 1, 2, 4, 8, 16, 32, 64, (char)128, 1, 2, 4, 8, 16, 32, 64, (char)128);
    __m256i a0 = _mm256_setzero_si256(), a1 = a0, a2 = a0, a3 = a0;
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m256i cls = qual_class_avx2(_mm256_loadu_si256((const __m256i*)(q + i)));
        const __m256i bit = _mm256_shuffle_epi8(bit_lut, _mm256_and_si256(cls, _mm256_set1_epi8(7)));
        const __m256i grp = _mm256_and_si256(_mm256_srli_epi16(cls, 3), _mm256_set1_epi8(3));
        a0 = _mm256_or_si256(a0, _mm256_and_si256(bit, _mm256_cmpeq_epi8(grp, _mm256_setzero_si256())));
        a1 = _mm256_or_si256(a1, _mm256_and_si256(bit, _mm256_cmpeq_epi8(grp, _mm256_set1_epi8(1))));
        a2 = _mm256_or_si256(a2, _mm256_and_si256(bit, _mm256_cmpeq_epi8(grp, _mm256_set1_epi8(2))));
        a3 = _mm256_or_si256(a3, _mm256_and_si256(bit, _mm256_cmpeq_epi8(grp, _mm256_set1_epi8(3))));
    }
    uint32_t m = 0;
    alignas(32) uint8_t t[4][32];
    _mm256_store_si256((__m256i*)t[0], a0); _mm256_store_si256((__m256i*)t[1], a1); _mm256_store_si256((__m256i*)t[2], a2); _mm256_store_si256((__m256i*)t[3], a3);
    for (int g = 0; g < 4; g++) { uint8_t o = 0; for (int k = 0; k < 32; k++) o |= t[g][k]; m |= (uint32_t)o << (8 * g); }
    for (; i < n; i++) m |= 1u << qual_class(q[i]);
    return m;
}
// the 2- / 4-bit codes of the first (L & ~31) qualities of a read; returns how many qualities it coded (a multiple of 32)
__attribute__((target("avx2"))) int code_quals_avx2(const uint8_t* quals, int L, const QualCoder& qc, uint8_t* out, bool* unknown) {
    const __m256i k0 = _mm256_broadcastsi128_si256(_mm_load_si128((const __m128i*)qc.class_known)), k1 = _mm256_broadcastsi128_si256(_mm_load_si128((const __m128i*)(qc.class_known + 16)));
    __m256i all_known = _mm256_set1_epi8((char)0xff);
    const __m256i t0 = _mm256_broadcastsi128_si256(_mm_load_si128((const __m128i*)qc.class_code)), t1 = _mm256_broadcastsi128_si256(_mm_load_si128((const __m128i*)(qc.class_code + 16)));
    int i = 0;
    for (; i + 32 <= L; i += 32) {
        const __m256i cls = qual_class_avx2(_mm256_loadu_si256((const __m256i*)(quals + i)));
        const __m256i hi16 = _mm256_cmpgt_epi8(cls, _mm256_set1_epi8(15)), low4 = _mm256_and_si256(cls, _mm256_set1_epi8(15));
        const __m256i code = _mm256_blendv_epi8(_mm256_shuffle_epi8(t0, cls), _mm256_shuffle_epi8(t1, low4), hi16);
        all_known = _mm256_and_si256(all_known, _mm256_blendv_epi8(_mm256_shuffle_epi8(k0, cls), _mm256_shuffle_epi8(k1, low4), hi16));
        if (qc.bits == 2) {
            const uint64_t c = squeeze2_avx2(code);
            std::memcpy(out + (i >> 2), &c, 8);
        } else {
            const __m256i y = _mm256_maddubs_epi16(code, _mm256_set1_epi16(0x1001));   // c0 + 16 c1 per 16-bit element
            const __m256i pk = _mm256_packus_epi16(y, y);                              // per 128-bit half: its eight bytes, twice
            const uint64_t lo = (uint64_t)_mm256_extract_epi64(pk, 0), hi = (uint64_t)_mm256_extract_epi64(pk, 2);
            std::memcpy(out + (i >> 1), &lo, 8);
            std::memcpy(out + (i >> 1) + 8, &hi, 8);
        }
    }
    if (_mm256_movemask_epi8(_mm256_cmpeq_epi8(all_known, _mm256_set1_epi8((char)0xff))) != -1) *unknown = true;
    return i;
}
#else
inline bool have_avx2() { return false; }
#endif

// Per-thread output of a contiguous range of reads.
struct Part {
    std::vector<ngsep_read_hdr> hdr;
    std::vector<uint32_t> payload;
    std::vector<uint8_t> qual;
    std::vector<uint16_t> group;               // read-group id per record (only when the tile carries groups)
    size_t n_hdr = 0, n_pay = 0, n_qual = 0;   // filled sizes (the vectors only grow)
    size_t n_full = 0;                         // payload words of the records in the full layout (n_pay counts the compact stream when that is packed)
    int max_span = 1, max_back = 0;
    uint64_t aligned_bases = 0;
    bool bad = false;
    bool unknown_quality = false;              // a quality class the tile's coder has no code for (a coder taken over from an earlier tile)
};

// bases2 / N mask / low-quality mask of one read, computed once and shared by its versions
struct ReadBits { uint32_t* b2; uint32_t* nm; uint32_t* lq; int nb, nw; bool has_n; };

void encode_bits(const uint8_t* bases, const uint8_t* quals, int len, ReadBits& o) {
    o.nb = (len + 15) / 16; o.nw = (len + 31) / 32;
    std::memset(o.b2, 0, sizeof(uint32_t) * (size_t)(o.nb + 2 * o.nw));
    o.nm = o.b2 + o.nb; o.lq = o.nm + o.nw;
    uint32_t any_bad = 0;
#ifdef NGSEP_HAVE_AVX2_PATH
    if (have_avx2()) { encode_bits_avx2(bases, quals, len, o.b2, o.nm, o.lq, &any_bad); o.has_n = any_bad != 0; return; }
#endif
    const CodeLut& L = lut();
    for (int i = 0; i < len; i += 8) {
        const int n = std::min(8, len - i);
        uint32_t bad;
        uint32_t c = encode8(load8(bases + i, n, 'A'), &bad);
        if (bad) {   // rare: a base outside ACGT — its code is 0 and its bit goes to the N mask
            c = 0;
            for (int k = 0; k < n; k++) c |= (uint32_t)(L.v[bases[i + k]] & 3u) << (2 * k);
            any_bad |= bad;
            o.nm[i >> 5] |= bad << (i & 31);
        }
        o.b2[i >> 4] |= c << ((i & 15) * 2);
        o.lq[i >> 5] |= lowq8(load8(quals + i, n, 0xff)) << (i & 31);
    }
    o.has_n = any_bad != 0;
}

// one read's qualities into the tile's quality array (padded run of qual_run_bytes): the bytes, or their 2- / 4-bit codes
inline void code_quality_run(const uint8_t* quals, int L, const QualCoder& qc, uint8_t* q, size_t padded, bool* unknown) {
    if (qc.bits == 8) {
        std::memcpy(q, quals, (size_t)L);
        for (size_t k = (size_t)L; k < padded; k++) q[k] = (uint8_t)'!';
        return;
    }
    int i = 0;
#ifdef NGSEP_HAVE_AVX2_PATH
    if (have_avx2()) i = code_quals_avx2(quals, L, qc, q, unknown);
#endif
    for (int k = i; k < L; k++) if (!qc.class_known[qual_class(quals[k])]) *unknown = true;   // (the qualities the loops below code one by one)
    if (qc.bits == 2) {
        size_t o = (size_t)i >> 2;
        for (; i + 4 <= L; i += 4, o++)
            q[o] = (uint8_t)(qc.code_of[quals[i]] | (qc.code_of[quals[i + 1]] << 2) | (qc.code_of[quals[i + 2]] << 4) | (qc.code_of[quals[i + 3]] << 6));
        if (i < L) { uint8_t b = 0; for (int k = 0; i + k < L; k++) b |= (uint8_t)(qc.code_of[quals[i + k]] << (2 * k)); q[o++] = b; }
        for (; o < padded; o++) q[o] = 0;
    } else {
        size_t o = (size_t)i >> 1;
        for (; i + 2 <= L; i += 2, o++) q[o] = (uint8_t)(qc.code_of[quals[i]] | (qc.code_of[quals[i + 1]] << 4));
        if (i < L) q[o++] = qc.code_of[quals[i]];
        for (; o < padded; o++) q[o] = 0;
    }
}

// One record of the tile.  admit_first: the ordering start; cigar / nc / st_first / ign_*: the read state the record describes
inline bool emit_record(int admit_first, int len, bool neg, const uint32_t* cigar, int nc, int st_first, int ign_s, int ign_e, const ReadBits& rb, bool windowed, int from,
                        int to, uint32_t qoff, Part& out, bool compact) {
    if (len > 0xffff || nc > 0xff) { out.bad = true; return false; }
    ngsep_read_hdr h;
    h.first = admit_first;
    h.meta = (uint32_t)len | ((uint32_t)nc << 16) | (neg ? NGSEP_M_NEG : 0u) | (rb.has_n ? NGSEP_M_NMASK : 0u) | (windowed ? NGSEP_M_WINDOW : 0u);
    if (nc == 1 && !rb.has_n && !windowed && cigar[0] == (((uint32_t)len << 3) | 3u)) h.meta |= NGSEP_M_SIMPLE;
    h.ign_start = (int16_t)ign_s;
    h.ign_end = (int16_t)ign_e;
    h.off = (uint32_t)out.n_pay;
    out.hdr[out.n_hdr++] = h;
    uint32_t* p = out.payload.data() + out.n_pay;
    out.n_full += (windowed ? 3u : 0u) + (size_t)nc + (size_t)rb.nb + (rb.has_n ? (size_t)rb.nw : 0u) + (size_t)rb.nw + 1u;
    if (compact && (h.meta & NGSEP_M_SIMPLE)) {
        // NGSEP_LAYOUT_COMPACT: the CIGAR word and the low-quality mask of a single-block record are rebuilt on the device
        std::memcpy(p, rb.b2, sizeof(uint32_t) * (size_t)rb.nb); p += rb.nb;
    } else {
        if (windowed) { *p++ = (uint32_t)from; *p++ = (uint32_t)to; *p++ = (uint32_t)st_first; }
        for (int k = 0; k < nc; k++) *p++ = cigar[(size_t)k];
        std::memcpy(p, rb.b2, sizeof(uint32_t) * (size_t)rb.nb); p += rb.nb;
        if (rb.has_n) { std::memcpy(p, rb.nm, sizeof(uint32_t) * (size_t)rb.nw); p += rb.nw; }
        std::memcpy(p, rb.lq, sizeof(uint32_t) * (size_t)rb.nw); p += rb.nw;
    }
    *p++ = qoff;
    out.n_pay = (size_t)(p - out.payload.data());
    int ref_last = st_first - 1;
    for (int k = 0; k < nc; k++) if (cig_ref(cigar[k])) ref_last += cig_len(cigar[k]);
    out.max_span = std::max(out.max_span, ref_last - admit_first + 1);
    out.max_back = std::max(out.max_back, admit_first - st_first);
    out.aligned_bases += (uint64_t)len;
    return true;
}
inline bool emit_record(const HostRead& r, const ReadState& st, const ReadBits& rb, bool windowed, int from, int to, uint32_t qoff, Part& out, bool compact) {
    return emit_record(r.admit_first, r.L, r.neg, st.cigar.data(), (int)st.cigar.size(), st.first, st.ign_s, st.ign_e, rb, windowed, from, to, qoff, out, compact);
}

// What the packer walks.  Read i of a source is a HostRead (any CIGAR, realigner versions) or — heavy(i) == nullptr — an untouched
// single-block read that exists only as a record of the batch (PlainRead).
struct PlainRead { int first = 0, L = 0, rg = -1; bool neg = false; int ign_s = 0, ign_e = 0; const uint8_t* bases = nullptr; const uint8_t* quals = nullptr; };
struct VectorSource {
    const std::vector<HostRead>* v;
    const HostRead* heavy(size_t i) const { return &(*v)[i]; }
    PlainRead plain(size_t) const { return PlainRead(); }
};
struct PlanSource {
    const SeqPlan* p;
    const HostRead* heavy(size_t i) const { const int32_t h = p->lite[i].heavy; return h >= 0 ? &p->heavy[(size_t)h] : nullptr; }
    PlainRead plain(size_t i) const {
        const LiteRead& l = p->lite[i];
        const BatchView& b = *p->batch;
        const int64_t rec = p->rec_base + l.rec;
        PlainRead r;
        r.first = l.first; r.L = l.last - l.first + 1; r.rg = b.rg ? b.rg[rec] : -1;
        r.neg = (b.flags[rec] & 0x10) != 0;
        if (r.neg) { r.ign_e = p->ignore5; r.ign_s = p->ignore3; } else { r.ign_s = p->ignore5; r.ign_e = p->ignore3; }   // RA:613-644
        r.bases = b.bases + b.base_off[rec];
        r.quals = b.quals + b.base_off[rec];
        return r;
    }
};

template <typename Src>
void pack_range(const Src& src, size_t lo, size_t hi, const QualCoder& qc, Part& out, bool compact, bool with_groups) {
    out.n_hdr = out.n_pay = out.n_qual = out.n_full = 0;
    out.max_span = 1; out.max_back = 0; out.aligned_bases = 0; out.bad = false; out.unknown_quality = false;
    size_t n_rec = 0, n_words = 0, n_q = 0;
    int max_len = 0;
    for (size_t i = lo; i < hi; i++) {   // exact upper bounds: one reservation per array
        const HostRead* hr = src.heavy(i);
        const int L = hr ? hr->L : src.plain(i).L;
        const size_t per = 3 + (size_t)((L + 15) / 16) + 2 * (size_t)((L + 31) / 32) + 1;
        size_t rec = 1, cig = 1;
        if (hr) {
            cig = hr->st.cigar.size();
            const bool plain = hr->closed.empty() && hr->st.first == hr->admit_first;
            if (!plain) { rec += hr->closed.size(); for (const Version& v : hr->closed) cig += v.st.cigar.size(); }
        }
        n_rec += rec; n_words += rec * per + cig; n_q += qual_run_bytes(L, qc.bits);
        max_len = std::max(max_len, L);
    }
    if (out.hdr.size() < n_rec) out.hdr.resize(n_rec + n_rec / 8);
    if (out.payload.size() < n_words) out.payload.resize(n_words + n_words / 8);
    if (out.qual.size() < n_q) out.qual.resize(n_q + n_q / 8);
    if (with_groups && out.group.size() < n_rec) out.group.resize(n_rec + n_rec / 8);
    std::vector<uint32_t> scratch((size_t)((max_len + 15) / 16 + 2 * ((max_len + 31) / 32)) + 4);
    ReadBits rb{};
    rb.b2 = scratch.data();
    for (size_t i = lo; i < hi; i++) {
        const HostRead* hr = src.heavy(i);
        PlainRead pr;
        if (!hr) pr = src.plain(i);
        const int L = hr ? hr->L : pr.L;
        const uint8_t* bases = hr ? hr->bases : pr.bases;
        const uint8_t* quals = hr ? hr->quals : pr.quals;
        // the read's qualities go once into the quality array (every version of the read points at them)
        const uint32_t qoff = (uint32_t)(out.n_qual / 4);
        uint8_t* q = out.qual.data() + out.n_qual;
        const size_t padded = qual_run_bytes(L, qc.bits);
        code_quality_run(quals, L, qc, q, padded, &out.unknown_quality);
        out.n_qual += padded;
        if (L > 0xffff) { out.bad = true; continue; }
        encode_bits(bases, quals, L, rb);
        const size_t first_rec = out.n_hdr;
        if (!hr) {
            const uint32_t cig = ((uint32_t)L << 3) | (uint32_t)kOpM;
            emit_record(pr.first, L, pr.neg, &cig, 1, pr.first, pr.ign_s, pr.ign_e, rb, false, 0, 0, qoff, out, compact);
        } else {
            const HostRead& r = *hr;
            if (r.closed.empty() && r.st.first == r.admit_first) emit_record(r, r.st, rb, false, 0, 0, qoff, out, compact);
            else {
                for (const Version& v : r.closed) emit_record(r, v.st, rb, true, v.from, v.to, qoff, out, compact);
                emit_record(r, r.st, rb, true, r.cur_from, INT_MAX, qoff, out, compact);
            }
        }
        if (with_groups) {   // one id per record: every version of a read carries the read's group
            const int rg = hr ? hr->rg : pr.rg;
            const uint16_t g = (rg >= 0 && rg < 0xffff) ? (uint16_t)rg : (uint16_t)0xffff;
            for (size_t k = first_rec; k < out.n_hdr; k++) out.group[k] = g;
        }
    }
}

template <typename Src>
bool pack_source(const Src& src, size_t lo, size_t hi, PackedTile& out, bool with_groups, std::string* err, int min_qual_bits, bool compact, QualHint* hint = nullptr) {
    out.group.clear();
    const size_t n = hi > lo ? hi - lo : 0;
    const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)host_threads(), n / 8192 + 1));   // small tiles are not worth a thread each
    static thread_local std::vector<Part> tl_parts;   // per calling thread: reused from tile to tile
    if ((int)tl_parts.size() < nt) tl_parts.resize((size_t)nt);
    std::vector<Part>& parts = tl_parts;              // (worker threads must not name the thread_local themselves)
    auto range_of = [&](int t, size_t& a, size_t& b) { a = lo + n * (size_t)t / (size_t)nt; b = lo + n * (size_t)(t + 1) / (size_t)nt; };
    // which quality classes occur in the tile -> bytes, or 2- / 4-bit codes
    std::vector<std::array<bool, 256>> seen((size_t)nt);
    auto scan = [&](int t) {
        size_t a, b;
        range_of(t, a, b);
        unsigned char hit[256];
        std::memset(hit, 0, sizeof hit);
        uint32_t m = 0;
        const bool vec = have_avx2();
        for (size_t i = a; i < b; i++) {
            const HostRead* hr = src.heavy(i);
            const uint8_t* q; int L;
            if (hr) { q = hr->quals; L = hr->L; } else { const PlainRead pr = src.plain(i); q = pr.quals; L = pr.L; }
#ifdef NGSEP_HAVE_AVX2_PATH
            if (vec) { m |= scan_classes_avx2(q, (size_t)L); continue; }   // the classes that occur (choose_coder only looks at classes)
#endif
            for (int k = 0; k < L; k++) hit[q[k]] = 1;
        }
        if (vec) for (int k = 0; k < 31; k++) if ((m >> k) & 1u) hit[33 + k] = 1;   // ... each reported through a character of the class
        for (int c = 0; c < 256; c++) seen[(size_t)t][(size_t)c] = hit[c] != 0;
    };
    // The coder of the previous tile is taken over when the caller keeps one (tiles of one run share their quality alphabet): the
    // pass over the qualities that finds the tile's classes is skipped, and the packing itself notices a class without a code —
    // the tile is then packed again the long way.
    const bool speculate = hint && hint->valid && (hint->bits == 2 || hint->bits == 4) && (int)hint->bits >= min_qual_bits;
    QualCoder qc;
    if (speculate) qc = coder_from_dict((int)hint->bits, hint->dict);
    else {
        Helpers::get().run(nt, scan);
        bool any[256];
        for (int c = 0; c < 256; c++) { any[c] = false; for (int t = 0; t < nt; t++) any[c] = any[c] || seen[(size_t)t][(size_t)c]; }
        qc = choose_coder(any, min_qual_bits);
    }
    out.qual_bits = (uint32_t)qc.bits;
    std::memcpy(out.qual_dict, qc.dict, 16);
    out.layout = compact ? NGSEP_LAYOUT_COMPACT : NGSEP_LAYOUT_FULL;
    Helpers::get().run(nt, [&](int t) { size_t a, b; range_of(t, a, b); pack_range(src, a, b, qc, parts[(size_t)t], compact, with_groups); });
    // contiguous ranges were packed independently: payload offsets move by the words in front of the range, quality offsets
    // (the last payload word of every record) by the quality words in front of it
    size_t nh = 0, np = 0, nq = 0;
    std::vector<size_t> bh((size_t)nt), bp((size_t)nt), bq((size_t)nt);
    bool bad = false;
    out.n_words_full = 0;
    if (speculate) for (int t = 0; t < nt; t++) if (parts[t].unknown_quality) { hint->valid = false; return pack_source(src, lo, hi, out, with_groups, err, min_qual_bits, compact, hint); }
    for (int t = 0; t < nt; t++) { bh[t] = nh; bp[t] = np; bq[t] = nq; nh += parts[t].n_hdr; np += parts[t].n_pay; nq += parts[t].n_qual; bad = bad || parts[t].bad; out.n_words_full += parts[t].n_full; }
    if (bad) {
        if (err) *err = "read longer than 65535 bases or with more than 255 CIGAR operations (after realignment): not supported by the device layout";
        out.hdr.clear(); out.payload.clear(); out.qual.clear();
        return false;
    }
    out.hdr.resize(nh); out.payload.resize(np); out.qual.resize(nq);
    if (with_groups) out.group.resize(nh);
    out.max_span = 1; out.max_back = 0; out.aligned_bases = 0;
    auto place = [&](int t) {
        Part& p = parts[(size_t)t];
        const uint32_t dp = (uint32_t)bp[t], dq = (uint32_t)(bq[t] / 4);
        ngsep_read_hdr* oh = out.hdr.data() + bh[t];
        uint32_t* op = out.payload.data() + bp[t];
        std::memcpy(op, p.payload.data(), p.n_pay * sizeof(uint32_t));
        std::memcpy(out.qual.data() + bq[t], p.qual.data(), p.n_qual);
        if (with_groups && p.n_hdr) std::memcpy(out.group.data() + bh[t], p.group.data(), p.n_hdr * sizeof(uint16_t));
        for (size_t i = 0; i < p.n_hdr; i++) {
            ngsep_read_hdr h = p.hdr[i];
            const size_t end = (i + 1 < p.n_hdr) ? p.hdr[i + 1].off : p.n_pay;
            op[end - 1] += dq;          // qoff
            h.off += dp;
            oh[i] = h;
        }
    };
    Helpers::get().run(nt, place);
    if (hint) { hint->valid = true; hint->bits = out.qual_bits; std::memcpy(hint->dict, out.qual_dict, 16); }
    for (int t = 0; t < nt; t++) { out.max_span = std::max(out.max_span, parts[t].max_span); out.max_back = std::max(out.max_back, parts[t].max_back); out.aligned_bases += parts[t].aligned_bases; }
    return true;
}

}  // namespace

bool pack_reads(const std::vector<HostRead>& reads, size_t lo, size_t hi, PackedTile& out, bool with_groups, std::string* err, int min_qual_bits, bool compact) {
    return pack_source(VectorSource{&reads}, lo, hi, out, with_groups, err, min_qual_bits, compact);
}
bool pack_plan_reads(const SeqPlan& plan, size_t lo, size_t hi, PackedTile& out, bool with_groups, std::string* err, int min_qual_bits, bool compact, QualHint* hint) {
    return pack_source(PlanSource{&plan}, lo, hi, out, with_groups, err, min_qual_bits, compact, hint);
}

ngsep_tile PackedTile::view(int seq_idx, int first, int last) const {
    ngsep_tile t;
    t.seq_idx = seq_idx; t.tile_first = first; t.tile_last = last;
    t.n_reads = (int)hdr.size();
    t.max_span = max_span; t.max_back = max_back;
    t.n_words = layout == NGSEP_LAYOUT_COMPACT ? n_words_full : payload.size();
    t.hdr = hdr.data();
    t.payload = payload.data();
    t.group = group.size() == hdr.size() && !group.empty() ? group.data() : nullptr;
    t.qual = qual.data();
    t.n_qual = qual.size();
    t.qual_bits = qual_bits;
    std::memcpy(t.qual_dict, qual_dict, 16);
    t.layout = layout;
    t.n_wire_words = payload.size();
    return t;
}

}  // namespace host
}  // namespace ngsep_b200
