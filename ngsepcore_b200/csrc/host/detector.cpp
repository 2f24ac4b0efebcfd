// GpuSingleSampleVariantsDetector: drop-in for SingleSampleVariantsDetector.findSNVs (SSVD:896-968) over in-memory,
// coordinate-sorted alignment records.  Per sequence:
//   1. admission            (RAFR filters, APG:407-433 cap)                               host, O(reads)
//   2. event pre-pass       indel realigner + span>1 calls at the sparse event positions  host, O(events x depth)
//   3. pack + ngsep_tile_submit / ngsep_tile_fetch per region tile                        device: every position
//   4. ordered merge        embedded / indel suppression (SSVPL:150-160), sort, VCF text  host, O(calls)
// Sequences are independent (every listener is reset at onSequenceStart): stages 1-2 of all sequences run in parallel on
// the host threads, stage 3 streams the tiles of all sequences through page-locked ring buffers with up to three tiles in
// flight (the next tile is packed while the device works on the previous ones), stage 4 runs in parallel again.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <deque>
#include <queue>
#include <set>
#include <thread>

#include "pipeline.hpp"

namespace ngsep_b200 {
namespace host {

namespace {
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}

GpuSingleSampleVariantsDetector::GpuSingleSampleVariantsDetector(const DetectorOptions& o) : opt_(o) {
    if (!opt_.het_modified && opt_.ploidy == 1) opt_.het_rate = 0.000001;  // SSVD:591-593
}

GpuSingleSampleVariantsDetector::~GpuSingleSampleVariantsDetector() {
    if (warm_.joinable()) warm_.join();
    rings_.clear();   // page-locked buffers go first (they unregister themselves)
    for (void* c : ctxs_) if (c) ngsep_ctx_destroy(c);
}

void GpuSingleSampleVariantsDetector::set_genome(const Genome* g) {
    genome_ = g;
    ref_uploaded_.assign((size_t)std::max(1, opt_.n_devices), std::vector<bool>(g->seqs.size(), false));
}

void GpuSingleSampleVariantsDetector::warm_up() {
    if (opt_.host_only || warm_.joinable() || !ctxs_.empty()) return;
    warm_ = std::thread([this]() { ensure_ctx(); });   // (a failure is found again, and reported, by run()'s own ensure_ctx)
}

int GpuSingleSampleVariantsDetector::ensure_ctx() {
    if (!ctxs_.empty()) return NGSEP_OK;
    const int n = std::max(1, opt_.n_devices);
    for (int k = 0; k < n; k++) {
        void* c = nullptr;
        int rc = create_ctx(opt_.device + k, &c);
        if (rc != NGSEP_OK) { for (void* x : ctxs_) ngsep_ctx_destroy(x); ctxs_.clear(); ctx_ = nullptr; return rc; }
        ctxs_.push_back(c);
    }
    ctx_ = ctxs_[0];
    rings_.resize((size_t)n);
    for (auto& r : rings_) r.resize((size_t)kHostRing);
    return NGSEP_OK;
}

int GpuSingleSampleVariantsDetector::create_ctx(int device, void** out) {
    void* ctx_ = nullptr;   // (local: the members are set by ensure_ctx)
    ngsep_params p{};
    p.max_base_qs = opt_.max_base_qs;
    p.min_quality = opt_.min_quality;
    p.het_rate = opt_.het_rate;
    p.ploidy = opt_.ploidy;
    p.flags = (opt_.csb ? NGSEP_F_STRAND_BIAS : 0u) | (opt_.ignore_lower ? NGSEP_F_IGNORE_LOWER_REF : 0u);
    p.n_samples = 1;
    double tables[NGSEP_N_TABLES];
    host_tables(opt_.het_rate, tables);
    int rc = ngsep_ctx_create(device, &p, tables, NGSEP_N_TABLES, &ctx_);
    if (rc != NGSEP_OK) { err_ = ngsep_last_error(nullptr); ctx_ = nullptr; }
    *out = ctx_;
    return rc;
}

namespace {
// f(k) for k in [0, n) on up to `threads` host threads, longest jobs first (`weight`), each k exactly once
template <typename F>
void parallel_jobs(size_t n, int threads, const std::vector<size_t>& order, F&& f) {
    if (threads <= 1 || n <= 1) { for (size_t k = 0; k < n; k++) f(order[k]); return; }
    std::atomic<size_t> next{0};
    auto work = [&]() { for (;;) { const size_t k = next.fetch_add(1); if (k >= n) return; f(order[k]); } };
    std::vector<std::thread> th;
    const int nt = (int)std::min<size_t>((size_t)threads, n);
    for (int t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto& x : th) x.join();
}
}  // namespace

int GpuSingleSampleVariantsDetector::run(const BatchView& b, std::string& out, bool with_header) {
    double t0 = now_s();
    if (!genome_) { err_ = "The reference genome file is a required parameter"; return NGSEP_EINVAL; }  // SSVD:596
    if (warm_.joinable()) warm_.join();
    int rc = opt_.host_only ? NGSEP_OK : ensure_ctx();
    if (rc != NGSEP_OK) return rc;
    if (ref_uploaded_.size() != (size_t)std::max(1, opt_.n_devices)) ref_uploaded_.assign((size_t)std::max(1, opt_.n_devices), std::vector<bool>(genome_->seqs.size(), false));
    if (!opt_.known_strs.empty() && opt_.known_variants.empty() && !strs_loaded_) {   // SSVD:906-913 (the else branch of -knownVariants)
        if (!load_known_strs(opt_.known_strs, *genome_, strs_, err_)) return NGSEP_EINVAL;
        strs_loaded_ = true;
    }
    if (!opt_.known_variants.empty() && !known_loaded_) {   // SSVD:897-905
        if (!load_known_variants(opt_.known_variants, *genome_, known_, err_)) return NGSEP_EINVAL;
        known_loaded_ = true;
        n_known_ = 0;
        for (auto& vs : known_) n_known_ += vs.size();
    }
    if (opt_.ploidy >= 3 && with_groups()) { err_ = "ploidy >= 3 (the pool algorithm) is built for SingleSampleVariantsDetector only"; return NGSEP_EINVAL; }
    if (with_header) out += header_text();
    stats_.records += (uint64_t)b.n;
    // ---- the sequences of the batch, in file order (APG:310-322 for a query sequence) ------------------------------------
    std::vector<SeqJob> jobs;
    bool query_found = false;
    int64_t i = 0;
    while (i < b.n) {
        int seq = b.seq_idx[i];
        int64_t j = i;
        while (j < b.n && b.seq_idx[j] == seq) j++;
        bool take = true;
        if (opt_.has_query) {
            bool is_q = seq >= 0 && seq < (int)genome_->names.size() && genome_->names[seq] == opt_.query_seq;
            if (is_q) query_found = true;
            else if (query_found) break;
            take = is_q;
        }
        if (take && seq >= 0 && seq < (int)genome_->seqs.size()) { jobs.emplace_back(); jobs.back().seq = seq; jobs.back().lo = i; jobs.back().hi = j; }
        i = j;
    }
    const int threads = opt_.synchronous ? 1 : host_threads();
    std::vector<size_t> order(jobs.size());
    for (size_t k = 0; k < order.size(); k++) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return jobs[x].hi - jobs[x].lo > jobs[y].hi - jobs[y].lo; });
    // ---- stages 1-2: admission + event pre-pass, every sequence on its own host thread (a run with fewer sequences than threads
    // also splits the work inside a sequence) -----------------------------------------------------------------------------------
    // (measured on the 16-core B200 host, yeast 30x: giving the LONGEST of many sequences threads of its own makes the stage slower,
    // 0.057 -> 0.065 s — the admitted reads are fresh memory and page faults serialise the threads —, so only a run with fewer sequences
    // than threads splits a sequence)
    inner_threads_ = (jobs.empty() || (int)jobs.size() >= threads) ? 1 : std::max(1, threads / (int)jobs.size());
    double tp = now_s();
    double t_plan = tp, t_tally = tp, t_dev = tp;
    auto after_plans = [&]() -> int {
        for (auto& job : jobs) {
            if (job.rc != NGSEP_OK) { err_ = job.plan.err; return job.rc; }
            stats_.add_host(job.plan.stats);
        }
        return NGSEP_OK;
    };
    auto tallies = [&]() -> int {
        if (deferred_tallies() && !known_mode()) {
            int r = tally_span_events(jobs);
            if (r != NGSEP_OK) return r;
            parallel_jobs(jobs.size(), threads, order, [&](size_t k) { resolve_fired_events(jobs[k].plan); });
            for (auto& job : jobs) stats_.indel_calls += job.plan.stats.indel_calls_late;
        }
        return NGSEP_OK;
    };
    if (opt_.host_only || opt_.synchronous) {
        parallel_jobs(jobs.size(), threads, order, [&](size_t k) { jobs[k].rc = plan_sequence(b, jobs[k].seq, jobs[k].lo, jobs[k].hi, jobs[k].plan); });
        stats_.prepass_s += now_s() - tp;
        t_plan = now_s();
        rc = after_plans();
        if (rc != NGSEP_OK) return rc;
        rc = tallies();
        if (rc != NGSEP_OK) return rc;
        t_tally = now_s();
        // ---- stage 3: the device ---------------------------------------------------------------------------------------
        if (!opt_.host_only) {
            size_t k = 0;
            rc = run_device(jobs, [&]() -> long { return k < jobs.size() ? (long)k++ : -1L; });
            if (rc != NGSEP_OK) return rc;
        }
        t_dev = now_s();
    } else {
        // The sequences are planned on the host threads while this thread already packs and submits the tiles of the sequences whose
        // plan is ready.  Both go through the sequences shortest first: the planners (sixteen at a time, each several times faster
        // than the packing of its sequence) stay ahead of this thread from the first small sequence on, the long sequences are
        // planned while the short ones are packed, and the tiles reach the ring in the same order in every run — a slot sees the
        // same tile sizes again, so nothing is re-allocated or re-registered after the first run.
        std::vector<size_t> asc(order.rbegin(), order.rend());
        std::stable_sort(asc.begin(), asc.end(), [&](size_t x, size_t y) { return jobs[x].hi - jobs[x].lo < jobs[y].hi - jobs[y].lo; });
        std::mutex mq;
        std::condition_variable cvq;
        std::vector<char> ready(jobs.size(), 0);
        size_t handed = 0;
        bool failed = false;
        double plan_wall = 0;
        std::thread planner([&]() {
            parallel_jobs(jobs.size(), threads, asc, [&](size_t k) {
                jobs[k].rc = plan_sequence(b, jobs[k].seq, jobs[k].lo, jobs[k].hi, jobs[k].plan);
                { std::lock_guard<std::mutex> g(mq); ready[k] = 1; if (jobs[k].rc != NGSEP_OK) failed = true; }
                cvq.notify_all();
            });
            plan_wall = now_s() - tp;
        });
        rc = run_device(jobs, [&]() -> long {
            std::unique_lock<std::mutex> g(mq);
            if (handed == jobs.size()) return -1L;
            const size_t k = asc[handed];
            cvq.wait(g, [&] { return ready[k] != 0 || failed; });
            if (failed) return -1L;   // a sequence could not be planned: nothing more is submitted
            handed++;
            return (long)k;
        });
        planner.join();
        stats_.prepass_s += plan_wall;
        t_plan = t_tally = now_s();
        const int prc = after_plans();
        if (prc != NGSEP_OK) return prc;
        if (rc != NGSEP_OK) return rc;
        t_dev = now_s();
        rc = tallies();   // (one device call for the events of every sequence: after the last plan)
        if (rc != NGSEP_OK) return rc;
        t_tally = now_s();
    }
    // ---- stage 4: ordered merge + text, per sequence ------------------------------------------------------------------------
    parallel_jobs(jobs.size(), threads, order, [&](size_t k) { if (!jobs[k].plan.lite.empty()) emit_text(jobs[k]); });
    size_t bytes = 0;
    for (auto& job : jobs) bytes += job.text.size();
    out.reserve(out.size() + bytes);
    for (auto& job : jobs) { out += job.text; stats_.snv_calls += job.snv_records; }
    {   // the plans hold millions of small allocations: release them on the host threads as well
        parallel_jobs(jobs.size(), threads, order, [&](size_t k) { std::vector<LiteRead>().swap(jobs[k].plan.lite); std::deque<HostRead>().swap(jobs[k].plan.heavy); });
    }
    stats_.total_s += now_s() - t0;
    if (std::getenv("NGSEP_TIMING"))
        std::fprintf(stderr, "run: setup %.4f  planning (wall) %.4f  device loop %.4f  tallies %.4f  text + release %.4f  total %.4f s%s\n", tp - t0, stats_.prepass_s,
                     (opt_.host_only || opt_.synchronous) ? t_dev - t_tally : t_dev - tp, (opt_.host_only || opt_.synchronous) ? t_tally - t_plan : t_tally - t_dev,
                     now_s() - std::max(t_dev, t_tally), now_s() - t0, (opt_.host_only || opt_.synchronous) ? "" : "  (planning runs under the device loop)");
    return NGSEP_OK;
}

// Stage 3.  Tiles of all sequences, in order, through the rings of page-locked host buffers: pack tile k (host threads) while
// the devices still work on the tiles before it (ngsep_tile_submit only enqueues copies and kernels).  With several devices
// the tiles are dealt to them in turn; every device has its own context, ring and tiles in flight, and the calls of all of
// them meet again in the per-sequence merge (no collective: SURVEY 8e).
int GpuSingleSampleVariantsDetector::run_device(std::vector<SeqJob>& jobs, const std::function<long()>& next_job) {
    const size_t n_dev = ctxs_.size();
    struct InFlight { SeqJob* job; size_t site_base, n_sites; };   // (-knownVariants: the tile's piece of plan.sites)
    std::vector<std::deque<InFlight>> q(n_dev);
    const bool known = known_mode();
    std::vector<int32_t> site_pos;
    std::vector<uint16_t> site_codes;
    std::vector<int> slot(n_dev, 0);
    const size_t max_in_flight = opt_.synchronous ? 1 : (size_t)kHostRing - 1;
    double submit_s = 0, fetch_s = 0;   // (NGSEP_TIMING)
    const char* hint_env = std::getenv("NGSEP_PACK_HINT");   // =0: every tile finds its own quality classes (A/B switch)
    QualHint* const hint = (hint_env && std::atoi(hint_env) == 0) ? nullptr : &qual_hint_;
    auto fetch_oldest = [&](size_t dev) -> int {
        double td = now_s();
        ngsep_tile_stats ts{};
        const InFlight f = q[dev].front();
        const size_t before = f.job->snvs.size();
        int rc = fetch_tile(ctxs_[dev], *f.job, ts);
        q[dev].pop_front();
        if (rc != NGSEP_OK) return rc;
        if (opt_.ploidy >= 3 && !known && !with_groups()) {   // pools: the supported columns, each followed by the slots that continue its evidence bytes
            std::vector<ngsep_snv_call>& snvs = f.job->snvs;
            size_t k = before;
            while (k < snvs.size()) {
                const ngsep_snv_call& c = snvs[k];
                if (c.kind != NGSEP_CALL_POOL || c.sample < 0) { err_ = "unexpected record from a pool context"; return NGSEP_ESTATE; }
                const size_t n_ev = (size_t)c.sample, inl = std::min<size_t>(n_ev, 128);
                const size_t extra = n_ev > 128 ? (n_ev - 128 + sizeof(ngsep_snv_call) - 1) / sizeof(ngsep_snv_call) : 0;
                if (k + extra >= snvs.size()) { err_ = "truncated pool column"; return NGSEP_ESTATE; }
                SeqJob::PoolColumn pc;
                pc.col = c;
                pc.ev.resize(n_ev);
                std::memcpy(pc.ev.data(), c.logcond, inl);
                if (n_ev > 128) std::memcpy(pc.ev.data() + 128, (const uint8_t*)&snvs[k + 1], n_ev - 128);
                f.job->pool_cols.push_back(std::move(pc));
                k += 1 + extra;
            }
            snvs.resize(before);
        }
        if (known && opt_.ploidy >= 3) {   // pools: one NGSEP_CALL_POOL column per site (post_max = its index), continuation slots behind
            std::vector<ngsep_snv_call>& snvs = f.job->snvs;
            size_t k = before;
            while (k < snvs.size()) {
                const ngsep_snv_call& c = snvs[k];
                const size_t n_ev = (size_t)std::max(0, c.sample), inl = std::min<size_t>(n_ev, 128);
                const size_t extra = n_ev > 128 ? (n_ev - 128 + sizeof(ngsep_snv_call) - 1) / sizeof(ngsep_snv_call) : 0;
                const long site = (long)c.post_max;
                if (c.kind != NGSEP_CALL_POOL || site < 0 || (size_t)site >= f.n_sites || k + extra >= snvs.size()) { err_ = "unexpected record from a tile of known sites"; return NGSEP_ESTATE; }
                const size_t idx = f.site_base + (size_t)site;
                f.job->site_calls[idx] = c;
                f.job->site_have[idx] = 1;
                f.job->site_ev[idx].resize(n_ev);
                std::memcpy(f.job->site_ev[idx].data(), c.logcond, inl);
                if (n_ev > 128) std::memcpy(f.job->site_ev[idx].data() + 128, (const uint8_t*)&snvs[k + 1], n_ev - 128);
                k += 1 + extra;
            }
            snvs.resize(before);
        } else if (known) {   // record k of a site tile belongs to site k of the list it was submitted with
            std::vector<ngsep_snv_call>& snvs = f.job->snvs;
            for (size_t k = before; k < snvs.size(); k++) {
                const ngsep_snv_call& c = snvs[k];
                if (c.kind != NGSEP_CALL_SITE || c.sample < 0 || (size_t)c.sample >= f.n_sites) { err_ = "unexpected record from a tile of known sites"; return NGSEP_ESTATE; }
                f.job->site_calls[f.site_base + (size_t)c.sample] = c;
                f.job->site_have[f.site_base + (size_t)c.sample] = 1;
            }
            snvs.resize(before);
        }
        stats_.device_s += now_s() - td;
        fetch_s += now_s() - td;
        stats_.covered_positions += ts.covered_positions;
        stats_.candidate_positions += ts.candidate_positions;
        stats_.aligned_bases += ts.aligned_bases;
        stats_.kernel_ms += ts.kernel_ms;
        stats_.kernel_launches += ts.kernel_launches;
        stats_.tiles++;
        return NGSEP_OK;
    };
    auto drain = [&]() {   // after an error: nothing may stay in flight
        for (size_t dev = 0; dev < n_dev; dev++) while (!q[dev].empty()) { ngsep_tile_stats ts{}; fetch_tile(ctxs_[dev], *q[dev].front().job, ts); q[dev].pop_front(); }
    };
    size_t dealt = 0;
    for (long jk = next_job(); jk >= 0; jk = next_job()) {
        SeqJob& job = jobs[(size_t)jk];
        if (job.rc != NGSEP_OK) continue;
        SeqPlan& plan = job.plan;
        if (plan.lite.empty()) continue;
        const int seq = job.seq;
        const int64_t tlen = tile_len();
        int tile_no = 0;
        const std::vector<KnownVariant>* kvars = (known && (size_t)seq < known_.size()) ? &known_[(size_t)seq] : nullptr;
        if (known) { job.site_calls.assign(plan.sites.size(), ngsep_snv_call{}); job.site_have.assign(plan.sites.size(), 0); job.site_ev.assign(plan.sites.size(), {}); }
        size_t site_next = 0;
        for (int64_t tf = plan.win_first; tf <= plan.win_last; tf += tlen, tile_no++) {
            const int tl = (int)std::min<int64_t>(plan.win_last, tf + tlen - 1);
            size_t site_base = 0, n_sites = 0;
            if (known) {   // the known SNV sites of the tile; a tile without one has nothing to report
                while (site_next < plan.sites.size() && (*kvars)[(size_t)plan.sites[site_next]].first < tf) site_next++;
                site_base = site_next;
                while (site_next < plan.sites.size() && (*kvars)[(size_t)plan.sites[site_next]].first <= tl) site_next++;
                n_sites = site_next - site_base;
                if (n_sites == 0) continue;
            }
            if (opt_.n_shards > 1 && (tile_no % opt_.n_shards) != opt_.shard) continue;
            size_t r_lo, r_hi;   // records whose ordering start can reach the tile
            tile_records(plan, (int)tf, tl, r_lo, r_hi);
            if (r_lo >= r_hi) continue;
            const size_t dev = dealt % n_dev;
            if (!ref_uploaded_[dev][(size_t)seq]) {
                int rc = ngsep_ref_upload(ctxs_[dev], seq, genome_->seqs[seq].data(), (int64_t)genome_->seqs[seq].size());
                if (rc != NGSEP_OK) { err_ = ngsep_last_error(ctxs_[dev]); drain(); return rc; }
                ref_uploaded_[dev][(size_t)seq] = true;
            }
            while (q[dev].size() >= max_in_flight) { int rc = fetch_oldest(dev); if (rc != NGSEP_OK) { drain(); return rc; } }
            PackedTile& tile = rings_[dev][(size_t)slot[dev]];
            tile.set_pinned(!opt_.synchronous);
            {   // the buffers of a ring grow together (a buffer takes the largest capacity any of them has needed): tiles of different
                // sizes cycling through the ring stop re-registering page-locked memory after the first few
                size_t ch = 0, cp = 0, cq = 0, cg = 0;
                for (const PackedTile& r : rings_[dev]) { ch = std::max(ch, r.hdr.cap); cp = std::max(cp, r.payload.cap); cq = std::max(cq, r.qual.cap); cg = std::max(cg, r.group.cap); }
                tile.hdr.reserve(ch, true); tile.payload.reserve(cp, true); tile.qual.reserve(cq, true);
                if (with_groups()) tile.group.reserve(cg, true);
            }
            double tp = now_s();
            if (!pack_plan_reads(plan, r_lo, r_hi, tile, with_groups(), &err_, 0, wire_compact(), hint)) { drain(); return NGSEP_EINVAL; }
            stats_.pack_s += now_s() - tp;
            if (tile.hdr.empty()) continue;
            double td = now_s();
            ngsep_tile tv = tile.view(seq, (int)tf, tl);
            int rc = NGSEP_OK;
            if (known) {
                site_pos.clear(); site_codes.clear();
                for (size_t k = 0; k < n_sites; k++) {
                    const KnownVariant& v = (*kvars)[(size_t)plan.sites[site_base + k]];
                    site_pos.push_back(v.first); site_codes.push_back(v.site_code);
                }
                rc = ngsep_tile_sites(ctxs_[dev], (int32_t)n_sites, site_pos.data(), site_codes.data());
            }
            if (rc == NGSEP_OK) rc = ngsep_tile_submit(ctxs_[dev], &tv);
            if (rc != NGSEP_OK) { err_ = ngsep_last_error(ctxs_[dev]); drain(); return rc; }
            stats_.device_s += now_s() - td;
            submit_s += now_s() - td;
            q[dev].push_back(InFlight{&job, site_base, n_sites});
            slot[dev] = (slot[dev] + 1) % kHostRing;
            dealt++;
        }
    }
    const double t_drain = now_s();
    for (size_t dev = 0; dev < n_dev; dev++) while (!q[dev].empty()) { int rc = fetch_oldest(dev); if (rc != NGSEP_OK) { drain(); return rc; } }
    if (std::getenv("NGSEP_TIMING")) std::fprintf(stderr, "device loop: %zu tiles, submit %.4f s, fetch %.4f s (of which the last tiles in flight %.4f s)\n", dealt, submit_s, fetch_s, now_s() - t_drain);
    return NGSEP_OK;
}

int GpuSingleSampleVariantsDetector::fetch_tile(void* ctx, SeqJob& job, ngsep_tile_stats& ts) {
    // into a buffer that lives as long as the detector (no fresh pages per tile), then only the tile's calls move to the job
    size_t n = 0;
    if (call_buf_.size() < 65536) call_buf_.resize(65536);
    int rc = ngsep_tile_fetch(ctx, call_buf_.data(), call_buf_.size(), &n, nullptr, &ts);
    if (rc == NGSEP_EOVERFLOW && n > call_buf_.size()) {
        call_buf_.resize(n + n / 4 + 16);
        rc = ngsep_tile_fetch(ctx, call_buf_.data(), call_buf_.size(), &n, nullptr, &ts);
    }
    if (rc != NGSEP_OK) { err_ = ngsep_last_error(ctx); return rc; }
    job.snvs.insert(job.snvs.end(), call_buf_.begin(), call_buf_.begin() + (ptrdiff_t)n);
    return NGSEP_OK;
}

int GpuSingleSampleVariantsDetector::plan_sequence(const BatchView& b, int seq, int64_t lo, int64_t hi, SeqPlan& plan) {
    const int seq_len = (int)genome_->seqs[seq].size();
    double t0 = now_s();
    std::vector<LiteRead>& reads = plan.lite;
    std::set<int>& suppressed = plan.suppressed;
    // ---- 1. admission -------------------------------------------------------------------------------------------
    AdmissionOptions ao;
    ao.min_mq = opt_.min_mq; ao.max_alns_per_start = opt_.max_alns_per_start; ao.secondary = opt_.secondary;
    ao.non_unique_primary = opt_.non_unique_primary; ao.ignore5 = opt_.ignore5; ao.ignore3 = opt_.ignore3;
    int64_t q_lo = lo, q_hi = hi;
    if (opt_.has_query) {  // records starting after queryLast end the run (APG:315); earlier ones are filtered below
        while (q_hi > q_lo && b.first[q_hi - 1] > opt_.query_last) q_hi--;
    }
    if (opt_.has_query) ao.query_first = opt_.query_first;
    plan.batch = &b; plan.rec_base = q_lo; plan.ignore5 = opt_.ignore5; plan.ignore3 = opt_.ignore3;
    plan.heavy.clear();
    reads = admit_sequence_lite(b, q_lo, q_hi, ao, inner_threads_);
    static const char* too_long = "read longer than 65535 bases or with more than 255 CIGAR operations: not supported by the device layout";
    for (const LiteRead& l : reads) if (b.base_off[q_lo + l.rec + 1] - b.base_off[q_lo + l.rec] > 0xffff) { plan.err = too_long; return NGSEP_EINVAL; }
    // the HostRead of a read, made when the read needs one: any CIGAR but one match block (at admission into the sweep), or the
    // first event that finds the read in its pileup
    auto heavy_of = [&](LiteRead& l) -> HostRead* {
        if (l.heavy < 0) {
            plan.heavy.push_back(make_read(b, q_lo + l.rec, ao));
            l.heavy = (int32_t)plan.heavy.size() - 1;
            plan.heavy.back().derive();
        }
        return &plan.heavy[(size_t)l.heavy];
    };
    plan.stats.admitted += reads.size();
    if (reads.empty()) { plan.win_first = 1; plan.win_last = seq_len; return NGSEP_OK; }
    const int win_first = plan.win_first = opt_.has_query ? std::max(1, opt_.query_first) : 1;
    const int win_last = plan.win_last = opt_.has_query ? std::min(seq_len, opt_.query_last) : seq_len;

    // ---- 2. event pre-pass ----------------------------------------------------------------------------------------
    Realigner realigner;
    realigner.genome = genome_;
    realigner.seq = seq;
    int last_indel_end = 0;                           // SSVPL:143
    {
        std::priority_queue<int, std::vector<int>, std::greater<int>> heap;
        std::vector<uint32_t> active;      // indexes into `reads`
        std::vector<HostRead*> pileup;
        std::vector<ReadState> before;
        size_t next = 0;
        auto cur_first = [&](const LiteRead& l) { return l.heavy >= 0 ? plan.heavy[(size_t)l.heavy].st.first : l.first; };
        auto cur_last = [&](const LiteRead& l) { return l.heavy >= 0 ? plan.heavy[(size_t)l.heavy].st.last : l.last; };
        // -knownSTRs: every input locus is an event at its first position, whether a read anchors an indel there or not, and the
        // positions behind it are embedded (IRPL:89-99); indels inside a locus are not events of their own
        static const std::vector<StrLocus> no_loci;
        static const std::vector<KnownVariant> no_known;
        // -knownVariants: the input variants are the realigner's loci as well (IRPL:141-157), and every variant is genotyped at its
        // first position (SSVPL:162-176): each of them is an event
        const bool known = known_mode();
        const std::vector<KnownVariant>& kvars = (known && (size_t)seq < known_.size()) ? known_[(size_t)seq] : no_known;
        const std::vector<StrLocus> kloci = known ? known_variant_loci(kvars) : std::vector<StrLocus>();
        const std::vector<StrLocus>& loci = known ? kloci : ((strs_loaded_ && (size_t)seq < strs_.size()) ? strs_[(size_t)seq] : no_loci);
        for (const StrLocus& l : loci) {
            if (l.starter) heap.push(l.first);
            if (!known && l.last > l.first) plan.embedded.push_back({l.first + 1, l.last});
        }
        for (const KnownVariant& v : kvars) heap.push(v.first);
        IndelCallerOptions kio;
        kio.max_base_qs = opt_.max_base_qs; kio.het_rate = opt_.het_rate; kio.min_quality = opt_.min_quality; kio.csb = false; kio.ploidy = opt_.ploidy;
        auto locus_at = [&](int pos) -> const StrLocus* {
            auto it = std::upper_bound(loci.begin(), loci.end(), pos, [](int v, const StrLocus& l) { return v < l.first; });
            if (it == loci.begin()) return nullptr;
            --it;
            return pos <= it->last ? &*it : nullptr;
        };
        while (next < reads.size() || !heap.empty()) {
            int s = next < reads.size() ? reads[next].first : INT_MAX;
            int e = heap.empty() ? INT_MAX : heap.top();
            if (s <= e) {  // every read starting at or before an event is pending when the event is processed (APG:377-403)
                LiteRead& l = reads[next];
                if (l.heavy == kLiteComplex) {   // (a single match block anchors no indel: nothing to derive)
                    HostRead* r = heavy_of(l);
                    if (r->st.cigar.size() > 0xff) { plan.err = too_long; return NGSEP_EINVAL; }
                    for (auto& a : r->anchors) heap.push(a.pos);
                }
                active.push_back((uint32_t)next++);
                continue;
            }
            while (!heap.empty() && heap.top() == e) heap.pop();
            if (e < win_first || e > win_last) continue;  // no pileup outside the query window (APG:481-484)
            bool prev_covered = false;   // (known variants) was e - 1 under a read?  Then a pileup — an empty one — is still handed out at e
            if (known && e - 1 >= win_first) for (uint32_t k : active) if (cur_first(reads[k]) <= e - 1 && e - 1 <= cur_last(reads[k])) { prev_covered = true; break; }
            size_t w = 0;
            for (size_t k = 0; k < active.size(); k++) if (cur_last(reads[active[k]]) >= e) active[w++] = active[k];  // APG:464-471
            active.resize(w);
            pileup.clear();
            for (uint32_t k : active) if (cur_first(reads[k]) <= e && e <= cur_last(reads[k])) pileup.push_back(heavy_of(reads[k]));       // PR:154-156
            const StrLocus* locus = locus_at(e);
            // inside an input locus: embedded, the realigner does nothing (IRPL:96-98)
            bool realign = !(locus && !(locus->starter && locus->first == e));
            if (realign && locus) { if (pileup.empty()) realign = false; }
            else if (realign) {
                bool any = false;
                for (HostRead* r : pileup) if (r->anchor_at(e)) { any = true; break; }
                if (!any) realign = false;
            }
            EventOutcome ev;
            if (realign) {
                before.clear();
                for (HostRead* r : pileup) before.push_back(r->st);
                ev = realigner.on_event(e, pileup, locus);
            }
            if (ev.fired) {
                plan.stats.events++;
                for (size_t k = 0; k < pileup.size(); k++) {
                    HostRead* r = pileup[k];
                    if (!r->derived) r->derive();
                    if (r->st == before[k]) continue;
                    if (r->cur_from == e) continue;  // cannot happen (one event per position); keeps windows well formed
                    r->closed.push_back(Version{r->cur_from, e, before[k]});
                    r->cur_from = e;
                    plan.stats.versions++;
                    for (auto& a : r->anchors) if (a.pos > e) heap.push(a.pos);
                }
            }
            if (known) {
                // ---- SingleSampleVariantPileupListener.onPileup with input variants (SSVPL:162-176): the variants that start at e, in
                // list order, given that the generator hands out a pileup here: a position under a read, or the first position behind
                // the last read of a covered stretch (APG:475-498: the reads that ended are only dropped after that empty pileup)
                if (pileup.empty() && !prev_covered) continue;
                auto it = std::lower_bound(kvars.begin(), kvars.end(), e, [](const KnownVariant& v, int p) { return v.first < p; });
                for (; it != kvars.end() && it->first == e; ++it) {
                    const int vi = (int)(it - kvars.begin());
                    if (it->is_snv) plan.sites.push_back(vi);
                    else plan.known_indels.push_back({vi, opt_.ploidy >= 3 ? genotype_known_indel_pool_at(e, it->alleles, pileup, kio) : genotype_known_indel_at(e, it->alleles, pileup, kio)});
                }
                continue;
            }
            if (!ev.fired) continue;
            if (deferred_tallies()) {   // the listener's part follows the device's tallies (resolve_fired_events)
                SeqPlan::FiredEvent fe;
                fe.e = e; fe.ev = ev;
                std::string ref;
                fe.ref_ok = genome_->get(seq, e, e + ev.span - 1, ref);
                fe.ref_lower = fe.ref_ok && ref[0] >= 'a' && ref[0] <= 'z';
                if (fe.ref_ok && ev.span > 1 && !(opt_.ignore_lower && fe.ref_lower)) {
                    fe.input = (int)plan.span_inputs.size();
                    plan.span_inputs.emplace_back();
                    gather_span_event(e, upper(ref), ev.str, ev.str && !ev.new_str, pileup, opt_.max_base_qs, plan.span_inputs.back());
                }
                plan.fired.push_back(fe);
                continue;
            }
            // ---- SingleSampleVariantPileupListener.onPileup at the event (SSVPL:146-161) ----
            bool emb = e <= last_indel_end;
            if (ev.str && !ev.new_str && e >= last_indel_end) { last_indel_end = e + ev.span - 1; emb = false; }   // an input STR (SSVPL:150)
            if (emb && !opt_.call_embedded) continue;     // SSVPL:193 — nothing is called here (interval already recorded)
            std::string ref;
            if (!genome_->get(seq, e, e + ev.span - 1, ref)) { suppressed.insert(e); continue; }   // SSVPL:195-196
            if (opt_.ignore_lower && ref[0] >= 'a' && ref[0] <= 'z') { suppressed.insert(e); continue; }
            if (emb || ev.span <= 1) continue;            // SNV path: the device decides
            on_span_event(e, upper(ref), ev, pileup, plan, last_indel_end);
        }
    }
    plan.stats.prepass_s += now_s() - t0;
    // reach of every read over all of its versions, to pick the records of a tile (halo = read span)
    plan.max_reach = 1;
    plan.max_back = 0;
    for (const LiteRead& l : reads) {
        if (l.heavy == kLiteComplex) { plan.err = "internal error: a read with a CIGAR of its own left the sweep without its HostRead"; return NGSEP_ESTATE; }
        if (l.heavy < 0) { plan.max_reach = std::max(plan.max_reach, l.last - l.first + 1); continue; }
        const HostRead& r = plan.heavy[(size_t)l.heavy];
        int lastp = r.st.last, firstp = r.st.first;
        for (auto& v : r.closed) { lastp = std::max(lastp, v.st.last); firstp = std::min(firstp, v.st.first); }
        plan.max_reach = std::max(plan.max_reach, lastp - r.admit_first + 1);
        plan.max_back = std::max(plan.max_back, r.admit_first - firstp);
    }
    return NGSEP_OK;
}

void GpuSingleSampleVariantsDetector::on_span_event(int e, const std::string& ref_upper, const EventOutcome& ev, std::vector<HostRead*>& pileup, SeqPlan& plan,
                                                    int& last_indel_end) {
    IndelCallerOptions io;
    io.max_base_qs = opt_.max_base_qs; io.het_rate = opt_.het_rate; io.min_quality = opt_.min_quality; io.csb = opt_.csb; io.ploidy = opt_.ploidy;
    IndelRecord rec;
    const bool input_str = ev.str && !ev.new_str;
    if (input_str) plan.suppressed.insert(e);      // no SNV is tried at an input locus when the indel alleles give no call (SSVPL:264)
    if (call_indel_at(e, ref_upper, ev.str, pileup, io, rec, input_str)) {
        plan.suppressed.insert(e);
        last_indel_end = rec.last;                 // SSVPL:157-158
        if (rec.last > e) plan.embedded.push_back({e + 1, rec.last});
        plan.indels.push_back(std::move(rec));
        plan.stats.indel_calls++;
    }
}

// The allele tallies (per-call fits, counts, strand counts, diagonal of logConditionalProbs) of every span > 1 event of the run in ONE
// device call (k_indel_tallies); the mixture terms and the posteriors are finished here (finish_tally)
int GpuSingleSampleVariantsDetector::tally_span_events(std::vector<SeqJob>& jobs) {
    size_t n_ev = 0;
    for (auto& job : jobs) n_ev += job.plan.span_inputs.size();
    if (n_ev == 0) return NGSEP_OK;
    std::vector<int32_t> allele_first{0}, call_first{0}, allele_off{0}, call_off{0};
    std::string allele_chars, call_chars;
    std::vector<uint8_t> call_quals, call_neg;
    for (auto& job : jobs)
        for (const SpanEventInput& in : job.plan.span_inputs) {
            for (auto& a : in.alleles) { allele_chars += a; allele_off.push_back((int32_t)allele_chars.size()); }
            for (auto& c : in.calls) {
                call_chars += c.allele;
                call_quals.insert(call_quals.end(), c.quals.begin(), c.quals.begin() + (ptrdiff_t)c.allele.size());
                call_off.push_back((int32_t)call_chars.size());
                call_neg.push_back(c.neg ? 1 : 0);
            }
            allele_first.push_back((int32_t)allele_off.size() - 1);
            call_first.push_back((int32_t)call_off.size() - 1);
        }
    if (allele_chars.size() > 0x7ffffff0ull || call_chars.size() > 0x7ffffff0ull) { err_ = "too many span events for one tally call"; return NGSEP_EINVAL; }
    ngsep_indel_events ev{};
    ev.n_events = (int32_t)n_ev; ev.n_alleles = (int32_t)allele_off.size() - 1; ev.n_calls = (int32_t)call_off.size() - 1;
    ev.allele_first = allele_first.data(); ev.call_first = call_first.data(); ev.allele_off = allele_off.data(); ev.call_off = call_off.data();
    ev.allele_chars = allele_chars.data(); ev.call_chars = call_chars.data(); ev.call_quals = call_quals.data(); ev.call_neg = call_neg.data();
    ev.log_err = indel_log_error();
    std::vector<int32_t> total(n_ev), counts((size_t)ev.n_alleles), strand(2 * (size_t)ev.n_alleles), call_index((size_t)ev.n_calls + 1);
    std::vector<double> diag((size_t)ev.n_alleles), call_fit((size_t)ev.n_calls + 1);
    int rc = ngsep_indel_tallies(ctx_, &ev, total.data(), counts.data(), strand.data(), diag.data(), call_index.data(), call_fit.data());
    if (rc != NGSEP_OK) { err_ = ngsep_last_error(ctx_); return rc; }
    size_t k = 0;
    for (auto& job : jobs) {
        job.plan.span_tallies.resize(job.plan.span_inputs.size());
        for (size_t i = 0; i < job.plan.span_inputs.size(); i++, k++) {
            const int a0 = allele_first[k], n = allele_first[k + 1] - a0, c0 = call_first[k], c1 = call_first[k + 1];
            CallTallies ct;
            ct.total = total[k];
            ct.counts.assign(counts.begin() + a0, counts.begin() + a0 + n);
            ct.strand.resize((size_t)n);
            for (int x = 0; x < n; x++) ct.strand[(size_t)x] = {strand[2 * (size_t)(a0 + x)], strand[2 * (size_t)(a0 + x) + 1]};
            ct.diag.assign(diag.begin() + a0, diag.begin() + a0 + n);
            ct.call_index.assign(call_index.begin() + c0, call_index.begin() + c1);
            ct.call_fit.assign(call_fit.begin() + c0, call_fit.begin() + c1);
            if (opt_.ploidy >= 3) { job.plan.span_call_tallies.resize(job.plan.span_inputs.size()); job.plan.span_call_tallies[i] = std::move(ct); }   // pools: finished per hypothesis
            else finish_tally(ct, (size_t)n, opt_.het_rate, job.plan.span_tallies[i]);   // the part that hangs on libm stays with the host's
        }
    }
    return NGSEP_OK;
}

// SingleSampleVariantPileupListener.onPileup (SSVPL:146-161) + discoverVariantWithSpan (SSVPL:257-273) over the fired events of one
// sequence, in order, with the device's tallies
void GpuSingleSampleVariantsDetector::resolve_fired_events(SeqPlan& plan) {
    IndelCallerOptions io;
    io.max_base_qs = opt_.max_base_qs; io.het_rate = opt_.het_rate; io.min_quality = opt_.min_quality; io.csb = opt_.csb; io.ploidy = opt_.ploidy;
    const bool pool = opt_.ploidy >= 3;
    int last_indel_end = 0;   // SSVPL:143
    for (const SeqPlan::FiredEvent& fe : plan.fired) {
        const int e = fe.e;
        const EventOutcome& ev = fe.ev;
        const bool input_str = ev.str && !ev.new_str;
        bool emb = e <= last_indel_end;
        if (input_str && e >= last_indel_end) { last_indel_end = e + ev.span - 1; emb = false; }   // SSVPL:150
        if (emb && !opt_.call_embedded) continue;            // SSVPL:193
        if (!fe.ref_ok) { plan.suppressed.insert(e); continue; }                                   // SSVPL:195-196
        if (opt_.ignore_lower && fe.ref_lower) { plan.suppressed.insert(e); continue; }
        if (emb || ev.span <= 1) continue;                   // SNV path: the device decides
        if (input_str) plan.suppressed.insert(e);            // no SNV is tried at an input locus when the indel alleles give no call (SSVPL:264)
        IndelRecord rec;
        const bool have = fe.input < 0 ? false
                          : (pool ? finish_indel_pool(plan.span_inputs[(size_t)fe.input], plan.span_call_tallies[(size_t)fe.input], io, opt_.ploidy, rec)
                                  : finish_indel_call(plan.span_inputs[(size_t)fe.input], plan.span_tallies[(size_t)fe.input], io, rec));
        if (have) {
            plan.suppressed.insert(e);
            last_indel_end = rec.last;                       // SSVPL:157-158
            if (rec.last > e) plan.embedded.push_back({e + 1, rec.last});
            plan.indels.push_back(std::move(rec));
            plan.stats.indel_calls_late++;
        }
    }
    std::vector<SpanEventInput>().swap(plan.span_inputs);
    std::vector<IndelTally>().swap(plan.span_tallies);
    std::vector<CallTallies>().swap(plan.span_call_tallies);
}

void GpuSingleSampleVariantsDetector::tile_records(const SeqPlan& plan, int tf, int tl, size_t& r_lo, size_t& r_hi) const {
    const auto& reads = plan.lite;
    r_lo = std::lower_bound(reads.begin(), reads.end(), tf - plan.max_reach + 1, [](const LiteRead& r, int v) { return r.first < v; }) - reads.begin();
    r_hi = std::upper_bound(reads.begin(), reads.end(), tl + plan.max_back, [](int v, const LiteRead& r) { return v < r.first; }) - reads.begin();
}

// -knownVariants: one record per input variant that met a pileup, in the order of the input list (the listener appends them in that
// order and SSVD:937 sorts by first, last — the order of the list — with a stable sort)
void GpuSingleSampleVariantsDetector::emit_known_text(SeqJob& job) {
    const std::string& chrom = genome_->names[job.seq];
    const std::vector<KnownVariant>& kvars = known_[(size_t)job.seq];
    SeqPlan& plan = job.plan;
    size_t si = 0, ii = 0;
    while (si < plan.sites.size() || ii < plan.known_indels.size()) {
        const bool take_site = ii >= plan.known_indels.size() || (si < plan.sites.size() && plan.sites[si] < plan.known_indels[ii].first);
        const int vi = take_site ? plan.sites[si] : plan.known_indels[ii].first;
        const KnownVariant& v = kvars[(size_t)vi];
        bool mine = true;   // records go to the shard that owns the tile of their position
        if (opt_.n_shards > 1) mine = (((v.first - plan.win_first) / tile_len()) % opt_.n_shards) == opt_.shard;
        if (mine) {
            if (take_site) {
                const bool have = si < job.site_have.size() && job.site_have[si];
                if (opt_.host_only) {   // (host-only test runs report the host-made records only)
                } else if (opt_.ploidy >= 3) {   // the pool branch of genotypeVariantSample (SSVPL:366-370) on the site's column
                    std::vector<int> al;
                    for (auto& a : v.alleles) al.push_back(dna_index(a[0]));
                    const ngsep_snv_call* col = have ? &job.site_calls[si] : nullptr;
                    SampleIndelCall sc = genotype_known_snv_pool(col, have ? job.site_ev[si].data() : nullptr, al, opt_.ploidy, opt_.het_rate, opt_.min_quality);
                    job.text += known_pool_snv_vcf_line(v, col, sc, chrom, opt_.ploidy);
                    job.snv_records++;
                } else {
                    job.text += known_snv_vcf_line(v, have ? &job.site_calls[si] : nullptr, chrom, opt_.ploidy);   // (nullptr: no record of the tile reaches the site)
                    job.snv_records++;
                }
            } else job.text += known_indel_vcf_line(v, plan.known_indels[ii].second, chrom, opt_.ploidy);
        }
        if (take_site) si++; else ii++;
    }
}

void GpuSingleSampleVariantsDetector::emit_text(SeqJob& job) {
    if (known_mode()) { emit_known_text(job); return; }
    const std::string& chrom = genome_->names[job.seq];
    SeqPlan& plan = job.plan;
    std::vector<ngsep_snv_call>& snvs = job.snvs;
    const std::set<int>& suppressed = plan.suppressed;
    // embedded positions: inside called indels / STRs and inside input loci — as one sorted list of disjoint intervals
    std::vector<std::pair<int, int>> embedded = plan.embedded;
    std::sort(embedded.begin(), embedded.end());
    {
        size_t w = 0;
        for (size_t k = 0; k < embedded.size(); k++) {
            if (w > 0 && embedded[k].first <= embedded[w - 1].second + 1) embedded[w - 1].second = std::max(embedded[w - 1].second, embedded[k].second);
            else embedded[w++] = embedded[k];
        }
        embedded.resize(w);
    }
    // ---- 4. ordered merge + text (SSVD:933-968) -------------------------------------------------------------------------
    std::sort(snvs.begin(), snvs.end(), [](const ngsep_snv_call& a, const ngsep_snv_call& c) { return a.pos < c.pos; });
    struct Line { int first, last, order; std::string text; };
    std::vector<Line> lines;
    lines.reserve(snvs.size() + plan.indels.size());
    size_t ei = 0;
    for (auto& c : snvs) {
        if (suppressed.count(c.pos)) continue;
        while (ei < embedded.size() && embedded[ei].second < c.pos) ei++;
        bool emb = ei < embedded.size() && embedded[ei].first <= c.pos && c.pos <= embedded[ei].second;
        if (emb && !opt_.call_embedded) continue;
        lines.push_back({c.pos, c.pos, (int)lines.size(), snv_vcf_line(c, chrom, opt_.ploidy, opt_.csb, emb)});
        job.snv_records++;
    }
    if (opt_.ploidy >= 3) {   // pools / polyploids: discoverSNV's pool branch on the columns the device found supported (SSVPL:240-254)
        std::sort(job.pool_cols.begin(), job.pool_cols.end(), [](const SeqJob::PoolColumn& a, const SeqJob::PoolColumn& c) { return a.col.pos < c.col.pos; });
        size_t pe = 0;
        for (auto& pc : job.pool_cols) {
            const int pos = pc.col.pos;
            if (suppressed.count(pos)) continue;
            while (pe < embedded.size() && embedded[pe].second < pos) pe++;
            const bool emb = pe < embedded.size() && embedded[pe].first <= pos && pos <= embedded[pe].second;
            if (emb && !opt_.call_embedded) continue;
            std::vector<int> alleles;
            PoolCall call;
            if (!discover_pool_snv(pc.col, pc.ev.data(), opt_.ploidy, opt_.het_rate, opt_.min_quality, alleles, call)) continue;
            lines.push_back({pos, pos, (int)lines.size(), pool_snv_vcf_line(pos, alleles, call, chrom, opt_.ploidy, emb)});
            job.snv_records++;
        }
    }
    for (auto& r : plan.indels) {
        bool mine = true;   // host-made records go to the shard that owns the tile of their position
        if (opt_.n_shards > 1) mine = (((r.pos - plan.win_first) / tile_len()) % opt_.n_shards) == opt_.shard;
        if (mine) lines.push_back({r.pos, r.last, (int)lines.size(), indel_vcf_line(r, chrom, opt_.ploidy)});
    }
    std::stable_sort(lines.begin(), lines.end(), [](const Line& a, const Line& c) { return a.first != c.first ? a.first < c.first : a.last < c.last; });
    size_t bytes = 0;
    for (auto& l : lines) bytes += l.text.size();
    job.text.reserve(bytes);
    for (auto& l : lines) job.text += l.text;
}

}  // namespace host
}  // namespace ngsep_b200

namespace ngsep_b200 {
namespace host {
// Admission + event pre-pass + packing of one whole sequence: exactly the tile run_sequence submits when tile_len == 0.
int GpuSingleSampleVariantsDetector::pack_sequence(const BatchView& b, int seq, PackedTile& tile, int& tile_first, int& tile_last) {
    if (!genome_) { err_ = "The reference genome file is a required parameter"; return NGSEP_EINVAL; }
    int64_t lo = 0;
    while (lo < b.n && b.seq_idx[lo] != seq) lo++;
    int64_t hi = lo;
    while (hi < b.n && b.seq_idx[hi] == seq) hi++;
    SeqPlan plan;
    inner_threads_ = host_threads();   // one sequence: its admission may use every host thread
    int rc = plan_sequence(b, seq, lo, hi, plan);
    if (rc != NGSEP_OK) { err_ = plan.err; return rc; }
    const double tpk = now_s();
    // (NGSEP_PACK_HINT=1, test hook: this one-tile entry point hands the quality coding on from call to call like run_device does from tile to tile)
    if (!pack_plan_reads(plan, 0, plan.lite.size(), tile, false, &err_, 0, wire_compact(), (std::getenv("NGSEP_PACK_HINT") && std::atoi(std::getenv("NGSEP_PACK_HINT")) != 0) ? &qual_hint_ : nullptr)) return NGSEP_EINVAL;
    if (std::getenv("NGSEP_TIMING")) std::fprintf(stderr, "pack_sequence: prepass %.4f s, pack %.4f s\n", plan.stats.prepass_s, now_s() - tpk);
    tile_first = plan.win_first;
    tile_last = plan.win_last;
    return NGSEP_OK;
}
}  // namespace host
}  // namespace ngsep_b200
