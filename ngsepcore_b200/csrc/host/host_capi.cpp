// C entry points of libngsep_host.so: the host pipeline (C++ mirror of the reference's Java detector) for callers
// without a C++ toolchain — the Python harness (ctypes) and, in a JVM deployment, the JNI glue of INTEGRATION.md.
// The device work goes through the C ABI of include/ngsep_pileup.h; nothing here computes the hot path on the CPU.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "../../../include/ngsep_host.h"
#include "pipeline.hpp"

using namespace ngsep_b200::host;

namespace {

thread_local uint32_t g_last_qual_bits = 8;
thread_local uint8_t g_last_qual_dict[16] = {0};
thread_local uint32_t g_last_layout = 0;
thread_local uint64_t g_last_words_full = 0;
void remember_quality(const PackedTile& t) {
    g_last_qual_bits = t.qual_bits; std::memcpy(g_last_qual_dict, t.qual_dict, 16);
    g_last_layout = t.layout; g_last_words_full = t.layout == NGSEP_LAYOUT_COMPACT ? t.n_words_full : (uint64_t)t.payload.size();
}

struct Handle {
    DetectorOptions opt;
    Genome genome;
    GpuSingleSampleVariantsDetector* det = nullptr;
    bool multi = false;
    std::vector<SampleDef> samples;
    int n_read_groups = 0;
    std::string vcf, err;
    ~Handle() { delete det; }
};

BatchView view_of(const ngsep_host_batch* b) {
    BatchView v;
    v.n = b->n; v.seq_idx = b->seq_idx; v.first = b->first; v.flags = b->flags; v.mapq = b->mapq; v.rg = b->rg; v.nh = b->nh;
    v.cigar_off = b->cigar_off; v.cigar = b->cigar; v.base_off = b->base_off; v.bases = b->bases; v.quals = b->quals; v.name_hash = b->name_hash;
    return v;
}

DetectorOptions convert(const ngsep_detector_options* o) {
    DetectorOptions d;
    if (o->sample_id) d.sample_id = o->sample_id;
    d.ploidy = o->ploidy; d.print_sample_ploidy = o->print_sample_ploidy != 0; d.min_mq = o->min_mq;
    if (o->query_seq && o->query_seq[0]) { d.query_seq = o->query_seq; d.has_query = true; }
    d.query_first = o->query_first; d.query_last = o->query_last;
    d.ignore_lower = o->ignore_lower_case_ref != 0; d.max_alns_per_start = o->max_alns_per_start_pos;
    d.non_unique_primary = o->process_non_unique_primary != 0; d.secondary = o->process_secondary != 0;
    d.ignore5 = o->ignore5; d.ignore3 = o->ignore3; d.het_rate = o->het_rate; d.het_modified = o->het_rate_set != 0;
    d.max_base_qs = o->max_base_qs; d.min_quality = o->min_quality; d.call_embedded = o->call_embedded_snvs != 0; d.csb = o->calc_strand_bias != 0;
    d.device = o->device; d.tile_len = o->tile_len; d.shard = o->shard; d.n_shards = o->n_shards < 1 ? 1 : o->n_shards; d.host_only = o->host_only != 0;
    d.synchronous = o->synchronous != 0;
    d.n_devices = o->n_devices < 1 ? 1 : o->n_devices;
    if (o->known_strs && o->known_strs[0]) d.known_strs = o->known_strs;
    if (o->known_variants && o->known_variants[0]) d.known_variants = o->known_variants;
    return d;
}

}  // namespace

extern "C" {

void ngsep_detector_options_default(ngsep_detector_options* o) {
    std::memset(o, 0, sizeof(*o));
    o->ploidy = 2; o->min_mq = 20; o->query_last = 1000000000; o->max_alns_per_start_pos = 5; o->het_rate = 0.001;
    o->max_base_qs = 30; o->min_quality = 40; o->n_shards = 1;
}

int ngsep_ssvd_create(const ngsep_detector_options* o, void** out) {
    if (!o || !out) return NGSEP_EINVAL;
    Handle* h = new Handle();
    h->opt = convert(o);
    *out = h;
    return NGSEP_OK;
}

int ngsep_msvd_create(const ngsep_detector_options* o, int n_read_groups, const char* const* rg_ids, const char* const* rg_samples,
                      double min_allele_depth_frequency, void** out) {
    if (!o || !out || n_read_groups < 1 || !rg_ids) return NGSEP_EINVAL;
    Handle* h = new Handle();
    h->opt = convert(o);
    h->opt.min_allele_depth_freq = min_allele_depth_frequency;
    h->multi = true;
    h->n_read_groups = n_read_groups;
    std::vector<std::string> ids, sms;
    for (int i = 0; i < n_read_groups; i++) { ids.emplace_back(rg_ids[i]); sms.emplace_back((rg_samples && rg_samples[i]) ? rg_samples[i] : ""); }
    h->samples = samples_from_read_groups(ids, sms, h->opt.ploidy);
    *out = h;
    return NGSEP_OK;
}

int ngsep_ssvd_destroy(void* p) { delete (Handle*)p; return NGSEP_OK; }

const char* ngsep_ssvd_error(void* p) { return p ? ((Handle*)p)->err.c_str() : ""; }

int ngsep_ssvd_set_genome(void* p, int n, const char* const* names, const int64_t* lengths, const char* const* bases) {
    Handle* h = (Handle*)p;
    if (!h || n <= 0) return NGSEP_EINVAL;
    h->genome.names.clear(); h->genome.seqs.clear();
    for (int i = 0; i < n; i++) { h->genome.names.emplace_back(names[i]); h->genome.seqs.emplace_back(bases[i], (size_t)lengths[i]); }
    delete h->det;
    if (h->multi) h->det = new GpuMultisampleVariantsDetector(h->opt, h->samples, h->n_read_groups);
    else h->det = new GpuSingleSampleVariantsDetector(h->opt);
    h->det->set_genome(&h->genome);
    return NGSEP_OK;
}

int ngsep_ssvd_run(void* p, const ngsep_host_batch* b, int with_header) {
    Handle* h = (Handle*)p;
    if (!h || !b) return NGSEP_EINVAL;
    if (!h->det) { h->err = "The reference genome file is a required parameter"; return NGSEP_ESTATE; }
    h->vcf.clear();
    int rc = h->det->run(view_of(b), h->vcf, with_header != 0);
    if (rc != NGSEP_OK) h->err = h->det->error();
    return rc;
}

const char* ngsep_ssvd_vcf(void* p, size_t* len) {
    Handle* h = (Handle*)p;
    if (len) *len = h->vcf.size();
    return h->vcf.c_str();
}

int ngsep_ssvd_stats(void* p, ngsep_run_stats* s) {
    Handle* h = (Handle*)p;
    if (!h || !h->det || !s) return NGSEP_EINVAL;
    const RunStats& r = h->det->stats();
    s->records = r.records; s->admitted = r.admitted; s->versions = r.versions; s->events = r.events; s->indel_calls = r.indel_calls;
    s->snv_calls = r.snv_calls; s->covered_positions = r.covered_positions; s->candidate_positions = r.candidate_positions;
    s->aligned_bases = r.aligned_bases; s->tiles = r.tiles; s->kernel_launches = r.kernel_launches; s->kernel_ms = r.kernel_ms;
    s->prepass_s = r.prepass_s; s->pack_s = r.pack_s; s->device_s = r.device_s; s->total_s = r.total_s;
    return NGSEP_OK;
}

// Pack-only entry (no device): admission + event pre-pass versions are NOT applied here; it packs the admitted
// records of one sequence exactly as the detector would for an indel-free input.  Used by bench.py to build
// device-resident tiles and by tests that compare the C++ packer with the numpy one.
int ngsep_host_pack(const ngsep_host_batch* b, int seq_idx, int min_mq, int max_alns_per_start, int ignore5, int ignore3,
                    ngsep_read_hdr** hdr, uint32_t** payload, uint8_t** qual, int64_t* n_reads, uint64_t* n_words, uint64_t* n_qual, int* max_span,
                    uint64_t* aligned_bases) {
    if (!b || !hdr || !payload || !qual) return NGSEP_EINVAL;
    BatchView v = view_of(b);
    int64_t lo = 0;
    while (lo < v.n && v.seq_idx[lo] != seq_idx) lo++;
    int64_t hi = lo;
    while (hi < v.n && v.seq_idx[hi] == seq_idx) hi++;
    AdmissionOptions ao;
    ao.min_mq = min_mq; ao.max_alns_per_start = max_alns_per_start; ao.ignore5 = ignore5; ao.ignore3 = ignore3;
    auto T0 = std::chrono::steady_clock::now();
    std::vector<HostRead> reads = admit_sequence(v, lo, hi, ao);
    auto T1 = std::chrono::steady_clock::now();
    PackedTile t;
    pack_reads(reads, 0, reads.size(), t);
    remember_quality(t);
    auto T2 = std::chrono::steady_clock::now();
    if (std::getenv("NGSEP_HOST_TIMING")) std::fprintf(stderr, "admit %.3f s pack %.3f s\n", std::chrono::duration<double>(T1 - T0).count(), std::chrono::duration<double>(T2 - T1).count());
    *n_reads = (int64_t)t.hdr.size();
    *n_words = t.payload.size();
    *max_span = t.max_span;
    if (aligned_bases) *aligned_bases = t.aligned_bases;
    *hdr = (ngsep_read_hdr*)std::malloc(t.hdr.size() * sizeof(ngsep_read_hdr) + 16);
    *payload = (uint32_t*)std::malloc(t.payload.size() * 4 + 64);
    *qual = (uint8_t*)std::malloc(t.qual.size() + 64);
    if (!*hdr || !*payload || !*qual) return NGSEP_ENOMEM;
    *n_qual = t.qual.size();
    std::memcpy(*hdr, t.hdr.data(), t.hdr.size() * sizeof(ngsep_read_hdr));
    std::memcpy(*payload, t.payload.data(), t.payload.size() * 4);
    std::memcpy(*qual, t.qual.data(), t.qual.size());
    return NGSEP_OK;
}

// The tile the detector itself would submit for a whole sequence (admission + event pre-pass versions + packing).
int ngsep_ssvd_pack(void* p, const ngsep_host_batch* b, int seq_idx, ngsep_read_hdr** hdr, uint32_t** payload, uint8_t** qual, int64_t* n_reads,
                    uint64_t* n_words, uint64_t* n_qual, int* max_span, int* max_back, int* tile_first, int* tile_last, uint64_t* aligned_bases) {
    Handle* h = (Handle*)p;
    if (!h || !h->det || !b || !hdr || !payload || !qual) return NGSEP_EINVAL;
    PackedTile t;
    int rc = h->det->pack_sequence(view_of(b), seq_idx, t, *tile_first, *tile_last);
    if (rc != NGSEP_OK) { h->err = h->det->error(); return rc; }
    remember_quality(t);
    *n_reads = (int64_t)t.hdr.size();
    *n_words = t.payload.size();
    *max_span = t.max_span;
    *max_back = t.max_back;
    if (aligned_bases) *aligned_bases = t.aligned_bases;
    *hdr = (ngsep_read_hdr*)std::malloc(t.hdr.size() * sizeof(ngsep_read_hdr) + 16);
    *payload = (uint32_t*)std::malloc(t.payload.size() * 4 + 64);
    *qual = (uint8_t*)std::malloc(t.qual.size() + 64);
    if (!*hdr || !*payload || !*qual) return NGSEP_ENOMEM;
    *n_qual = t.qual.size();
    std::memcpy(*hdr, t.hdr.data(), t.hdr.size() * sizeof(ngsep_read_hdr));
    std::memcpy(*payload, t.payload.data(), t.payload.size() * 4);
    std::memcpy(*qual, t.qual.data(), t.qual.size());
    return NGSEP_OK;
}

int ngsep_host_merge_order(int n_inputs, const ngsep_host_batch* inputs, int min_mq, int process_non_unique_primary, int process_secondary,
                           int32_t** input_idx, int64_t** record_idx, int64_t* n_out) {
    if (n_inputs < 1 || !inputs || !input_idx || !record_idx || !n_out) return NGSEP_EINVAL;
    std::vector<BatchView> views;
    for (int i = 0; i < n_inputs; i++) views.push_back(view_of(&inputs[i]));
    AdmissionOptions ao;
    ao.min_mq = min_mq; ao.non_unique_primary = process_non_unique_primary != 0; ao.secondary = process_secondary != 0;
    std::vector<std::pair<int, int64_t>> order;
    merge_order(views, ao, order);
    *n_out = (int64_t)order.size();
    *input_idx = (int32_t*)std::malloc(order.size() * sizeof(int32_t) + 8);
    *record_idx = (int64_t*)std::malloc(order.size() * sizeof(int64_t) + 8);
    if (!*input_idx || !*record_idx) return NGSEP_ENOMEM;
    for (size_t k = 0; k < order.size(); k++) { (*input_idx)[k] = order[k].first; (*record_idx)[k] = order[k].second; }
    return NGSEP_OK;
}

void ngsep_host_free(void* p) { std::free(p); }

// CoverageStats (discovery/CoverageStatisticsCalculator.java): the coverage table of the batch as malloc'ed text (ngsep_host_free).
int ngsep_coverage_stats(int n_seqs, const char* const* names, const int64_t* lengths, const ngsep_host_batch* b, int min_mq, int device, char** text) {
    if (n_seqs <= 0 || !names || !lengths || !b || !text) return NGSEP_EINVAL;
    Genome g;
    for (int i = 0; i < n_seqs; i++) { g.names.emplace_back(names[i]); g.seqs.emplace_back((size_t)lengths[i], 'N'); }   // only the lengths matter here
    std::string out, err;
    int rc = run_coverage_stats(g, view_of(b), min_mq, device, out, err);
    if (rc != NGSEP_OK) return rc;
    *text = (char*)std::malloc(out.size() + 1);
    if (!*text) return NGSEP_ENOMEM;
    std::memcpy(*text, out.c_str(), out.size() + 1);
    return NGSEP_OK;
}

int ngsep_relative_allele_counts(int n_seqs, const char* const* names, const int64_t* lengths, const ngsep_host_batch* b, int min_rd, int max_rd, int min_base_qs,
                                 int secondary, int device, char** text) {
    return ngsep_relative_allele_counts_regions(n_seqs, names, lengths, b, min_rd, max_rd, min_base_qs, secondary, nullptr, nullptr, device, text);
}

int ngsep_relative_allele_counts_regions(int n_seqs, const char* const* names, const int64_t* lengths, const ngsep_host_batch* b, int min_rd, int max_rd,
                                         int min_base_qs, int secondary, const char* regions_to_filter, const char* regions_to_select, int device, char** text) {
    if (n_seqs <= 0 || !names || !lengths || !b || !text) return NGSEP_EINVAL;
    Genome g;
    for (int i = 0; i < n_seqs; i++) { g.names.emplace_back(names[i]); g.seqs.emplace_back((size_t)lengths[i], 'N'); }   // only the lengths matter here
    std::string out, err;
    int rc = run_relative_allele_counts(g, view_of(b), min_rd, max_rd, min_base_qs, secondary != 0, device, out, err, regions_to_filter ? regions_to_filter : "",
                                        regions_to_select ? regions_to_select : "");
    if (rc != NGSEP_OK) return rc;
    *text = (char*)std::malloc(out.size() + 1);
    if (!*text) return NGSEP_ENOMEM;
    std::memcpy(*text, out.c_str(), out.size() + 1);
    return NGSEP_OK;
}

int ngsep_host_pack_layout(uint32_t* layout, uint64_t* n_words_full) {
    if (!layout || !n_words_full) return NGSEP_EINVAL;
    *layout = g_last_layout;
    *n_words_full = g_last_words_full;
    return NGSEP_OK;
}

int ngsep_host_pack_quality(uint32_t* qual_bits, uint8_t* qual_dict16) {
    if (!qual_bits || !qual_dict16) return NGSEP_EINVAL;
    *qual_bits = g_last_qual_bits;
    std::memcpy(qual_dict16, g_last_qual_dict, 16);
    return NGSEP_OK;
}

}  // extern "C"
