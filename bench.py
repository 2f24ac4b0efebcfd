#!/usr/bin/env python
"""bench.py — genotyped positions/s of the pileup / SNVQ hot path (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port) on the host cores

A "step" = one pass of the hot path over the whole workload: BASELINE configs[1], a yeast-like 12.16 Mb genome at
30x (2x100 bp), simulated with planted SNVs/indels (ngsepcore_b200/simulate.py), one region tile per sequence, all
tiles walked by ONE k_pileup_windows launch followed by ONE k_genotype_columns launch (ngsep_resident_run(-1)).
  value  = positions with >= 1 covering admitted read / device time of the step, tiles resident in HBM (CUDA events on
           the context stream; the step's input is 344 MB in HBM > 126 MB L2, so no step is served from cache);
  e2e    = the same metric through the C ABI with HOST buffers: ngsep_tile_submit (H2D from page-locked memory +
           kernels) + ngsep_tile_fetch (D2H of the calls) per tile, all inside the timed region;
  roofline.achieved = ALGORITHMIC bytes (sum_reads(16 + 4 n_cigar + ceil(L/4) + L) + 1 B/reference base +
           64 B/call) / device time of the WHOLE step, k_pileup_windows + k_genotype_columns (CUDA events around the two
           launches, same stream); the per-kernel shares sit beside it.  The window kernel does not need every one of those
           bytes: quality bytes are read only for the few positions that reach the genotype kernel (it works from a 1-bit
           "quality <= 3" mask), so its measured DRAM traffic (roofline.traffic, from the committed ncu capture) is BELOW the
           algorithmic bytes; k_pileup_windows.frac_of_traffic gives the fraction of peak that traffic amounts to.
N > 1 (torchrun): every rank owns one GPU and an independent, equally sized region shard (weak scaling, no data-path
collective — SURVEY.md 8e); value = positions of all ranks / max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "genotyped_positions_per_sec"
UNIT = "positions/s"
WORKLOAD = "yeast-like 12.16 Mb x 30x, 2x100 bp simulated alignments with planted SNVs/indels (BASELINE configs[1]), one tile per sequence"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_traffic():
    """DRAM bytes of one k_pileup_windows launch over this workload, from the committed ncu capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_pileup_windows.json")
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "r01_ncu_pileup_windows.json")
    try:
        d = json.load(open(p))
        unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        return sum(float(d[k]["value"]) * unit[d[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    except Exception:
        return None


def _workload(rank: int, depth: float, scale: float):
    from ngsepcore_b200 import simulate
    contigs = simulate.YEAST_CONTIGS
    if scale != 1.0:
        contigs = [(n, max(20000, int(l * scale))) for n, l in contigs]
    seeds = (1001 + 17 * rank, 2002 + 17 * rank, 3003 + 17 * rank)
    return simulate.simulate(contigs=contigs, depth=depth, seeds=seeds)


def _algorithmic_bytes(batch, admitted_hdrs, genome_len, n_calls):
    """per admitted record 16 + 4 n_cigar + ceil(L/4) + L, + 1 byte per reference base + 64 per call."""
    tot = 0
    for hdr in admitted_hdrs:
        L = (hdr["meta"] & 0xFFFF).astype(np.int64)
        nc = ((hdr["meta"] >> 16) & 0xFF).astype(np.int64)
        tot += int((16 + 4 * nc + (L + 3) // 4 + L).sum())
    return tot + genome_len + 64 * n_calls


class _ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25", "-i", str(device)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f:
            t = [x.strip() for x in line.split(",")]
            if len(t) < 8:
                continue
            try:
                sm.append(float(t[1])); mx.append(float(t[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def reduce_over_ranks(dist, device, sums, maxima):
    """Whole-job aggregates: units are SUMMED over ranks, times take the MAX over ranks (contract: value = units of all ranks /
    max-over-ranks time).  dist is None for a single process."""
    if dist is None:
        return [float(x) for x in sums], [float(x) for x in maxima]
    import torch
    w = torch.tensor([float(x) for x in sums], dtype=torch.float64, device=device)
    dist.all_reduce(w, op=dist.ReduceOp.SUM)
    v = torch.tensor([float(x) for x in maxima], dtype=torch.float64, device=device)
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return [float(x) for x in w.tolist()], [float(x) for x in v.tolist()]


def _dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        # The communicator setup prints "NCCL version ..." on STDOUT (whatever NCCL_DEBUG says): this program prints ONE JSON
        # line there, so file descriptor 1 points at stderr while the process group and its first collective come up.
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist_mod.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
        dist = dist_mod
    return rank, world, local, dist


# ----------------------------------------------------------------------------------------------------------------------
# CUDA arm
# ----------------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    rank, world, local, dist = _dist_setup(args.gpus)
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if dist is not None:
        dist.barrier()
    from ngsepcore_b200 import capi, hostapi
    device = local if world > 1 else 0

    t_sim = time.time()
    genome, batch, _ = _workload(rank, args.depth, args.scale)
    names = list(genome)
    t_sim = time.time() - t_sim

    # host stages once (admission + event pre-pass + packing): the very tiles the detector submits
    t_pack = time.time()
    det = hostapi.Detector(device=device, host_only=True)
    det.set_genome(genome)
    tiles = [det.pack(batch, ci) for ci in range(len(names))]
    t_pack = time.time() - t_pack
    for t in tiles:  # page-lock the host buffers, as a JVM host does for its direct ByteBuffers
        if t["hdr"].size:
            capi.host_register(t["hdr"]); capi.host_register(t["payload"]); capi.host_register(t["qual"])
    h2d_bytes = sum(t["hdr"].nbytes + t["payload"].nbytes + t["qual"].nbytes for t in tiles)
    resident_bytes = sum(t["hdr"].nbytes + 4 * t["layout"][1] + t["qual"].nbytes for t in tiles)   # the payload in its full layout, as staged in HBM
    aligned_bases = sum(t["aligned_bases"] for t in tiles)
    genome_len = sum(len(v) for v in genome.values())

    # the same mode at every N: every host buffer of a tile is copied (qualities travel as 2-bit codes when the tile's
    # qualities allow it, ngsep_tile.qual_bits); --quality-in-place leaves page-locked quality buffers on the host instead
    ctx = capi.Context(device=device, flags=capi.F_QUALITY_IN_PLACE if args.quality_in_place else 0)
    for ci, n in enumerate(names):
        ctx.ref_upload(ci, genome[n])
    slots = []
    for ci, t in enumerate(tiles):
        if t["hdr"].size:
            slots.append(ctx.resident_add(ci, t["first"], t["last"], t["hdr"], t["payload"], t["qual"], t["max_span"], t["max_back"], quality=t["quality"], layout=t["layout"]))

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- resident (HBM) measurement: every tile of the genome in ONE launch per step ---------------------------------------
    for _ in range(max(3, args.warmup)):
        ctx.resident_run(-1)
    st = ctx.resident_stats()
    covered, calls, cand = st.covered_positions, st.calls, st.candidate_positions
    launches = st.kernel_launches
    sampler = _ClockSampler(device) if rank == 0 else None
    barrier()
    pileup_ms = 0.0
    kernels_ms = 0.0
    step_ms = []
    for _ in range(args.steps):
        ctx.timer_start()
        ctx.resident_run(-1)
        step_ms.append(ctx.timer_stop())
        rs = ctx.resident_stats()
        pileup_ms += rs.pileup_ms    # k_pileup_windows alone (events around the launch, same stream)
        kernels_ms += rs.kernel_ms   # + k_genotype_columns
    barrier()
    total_ms = float(sum(step_ms))
    # ---- end-to-end through the C ABI with host buffers ------------------------------------------------------------------
    IN_FLIGHT = 3   # tiles between ngsep_tile_submit and ngsep_tile_fetch (the C ABI pipelines up to 4)

    def e2e_step():
        n_calls, pending, copied, in_place, back = 0, 0, 0, 0, 0
        for ci, t in enumerate(tiles):
            if not t["hdr"].size:
                continue
            if pending == IN_FLIGHT:
                c, _, st = ctx.fetch(sort=False)
                n_calls += len(c); pending -= 1; copied += st.h2d_bytes; in_place += st.quality_in_place; back += st.d2h_bytes
            ctx.submit(ci, t["first"], t["last"], t["hdr"], t["payload"], t["qual"], t["max_span"], t["max_back"], quality=t["quality"], layout=t["layout"])
            pending += 1
        while pending:
            c, _, st = ctx.fetch(sort=False)
            n_calls += len(c); pending -= 1; copied += st.h2d_bytes; in_place += st.quality_in_place; back += st.d2h_bytes
        return n_calls, copied, in_place, back
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        n_calls_e2e, e2e_copied, e2e_in_place, e2e_back = e2e_step()
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if sampler else None

    # ---- whole detector from decoded records (host stages + device + VCF text), informational --------------------------------
    pipeline = None
    if not args.no_pipeline and rank == 0:   # (one rank: the detector uses every host thread, N of them at once would time their contention)
        det2 = hostapi.Detector(device=device)
        det2.set_genome(genome)
        runs = []
        for _ in range(6):                   # run 1: context creation, device allocations, page-locked buffers sized; runs 2 .. 6 are what a long-lived detector does
            s0 = det2.stats()
            w0 = (s0.covered_positions, s0.prepass_s, s0.pack_s, s0.device_s)
            t0 = time.perf_counter()
            vcf = det2.run(batch)
            dt = time.perf_counter() - t0
            st2 = det2.stats()
            runs.append((dt, st2.covered_positions - w0[0], st2.prepass_s - w0[1], st2.pack_s - w0[2], st2.device_s - w0[3]))
        later = sorted(runs[2:])
        pipeline_s, pcov, ppre, ppack, pdev = later[len(later) // 2]      # the median of runs 3 .. 6
        pipeline = {"value": pcov / pipeline_s, "unit": UNIT, "seconds": pipeline_s, "prepass_s": ppre,
                    "pack_s": ppack, "device_s": pdev, "vcf_records": len([l for l in vcf.splitlines() if l and l[0] != "#"]),
                    "host_threads": min(32, os.cpu_count() or 1), "run_seconds": [round(r[0], 4) for r in runs],
                    "note": "whole detector from decoded records to VCF text, one detector object run six times (run_seconds; the first run creates the "
                            "device context and sizes the page-locked buffers): the median of runs 3 .. 6.  Admission + realigner pre-pass of the "
                            "sequences on the host threads (prepass_s = wall time of that stage, which runs under the device loop), tiles packed into "
                            "page-locked ring buffers with up to three tiles in flight (pack_s, device_s = host time spent packing / in submit + fetch), VCF text"}

    (covered_all, bases_all, h2d_all, calls_all, launches_all, d2h_all), (total_ms, e2e_s, pileup_ms, kernels_ms) = reduce_over_ranks(
        dist, "cuda", sums=[covered, aligned_bases, e2e_copied, calls, launches, e2e_back], maxima=[total_ms, e2e_s, pileup_ms, kernels_ms])

    if rank == 0:
        peak, peak_src = _peaks()
        alg_bytes = _algorithmic_bytes(batch, [t["hdr"] for t in tiles], genome_len, calls)
        quality_bytes_alg = sum(int((t["hdr"]["meta"] & 0xFFFF).astype(np.int64).sum()) for t in tiles)
        ms_per_step = total_ms / args.steps
        # achieved = algorithmic bytes of the path / device-resident time of the path = BOTH kernels of the step
        step_kernels_s = kernels_ms / args.steps / 1e3
        achieved = alg_bytes / step_kernels_s / 1e9
        win_s = pileup_ms / args.steps / 1e3
        # what each kernel reads of the canonical input: the window kernel headers + CIGAR + 2-bit bases + reference, the genotype
        # kernel the quality bytes and the call records
        own_win = alg_bytes - quality_bytes_alg - 64 * calls
        traffic = _ncu_traffic() if (args.scale == 1.0 and args.depth == 30.0) else None
        line = {
            "metric": METRIC, "value": covered_all / (ms_per_step / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "depth": args.depth, "genome_bp": genome_len, "reads": int(batch.n), "tiles": len(slots),
                       "genotyped_positions_per_gpu": int(covered), "candidate_positions": int(cand), "open_columns": int(st.open_columns), "calls": int(calls),
                       "l2": "inputs larger than L2 (%.0f MB of packed tiles resident in HBM, walked every step)" % (resident_bytes / 1e6),
                       "parallelism": f"region shards x{world}, no collective"},
            "aligned_bases_per_sec": bases_all / (ms_per_step / 1e3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic,
                         "kernel": "the device step: k_pileup_windows + k_genotype_columns (one launch each)", "algorithmic_bytes_per_step": alg_bytes,
                         "kernels_ms_per_step": kernels_ms / args.steps, "launches_per_step": 2,
                         "k_pileup_windows": {"ms_per_step": pileup_ms / args.steps, "share_of_step": pileup_ms / kernels_ms,
                                              "own_algorithmic_bytes": own_win, "own_frac": own_win / win_s / 1e9 / peak,
                                              "frac_of_traffic": (traffic / win_s / 1e9 / peak) if traffic else None},
                         "k_genotype_columns": {"ms_per_step": (kernels_ms - pileup_ms) / args.steps, "share_of_step": 1.0 - pileup_ms / kernels_ms},
                         "traffic_note": "ncu dram__bytes_read+write of one k_pileup_windows launch on this workload (committed capture under profiles/); "
                                         "quality bytes are only read for the columns that reach k_genotype_columns",
                         "peak_source": peak_src},
            "e2e": {"value": covered_all * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d_all),
                    "d2h_bytes_per_step": int(d2h_all), "ms_per_step": e2e_s / args.steps * 1e3,
                    "host_input_bytes_per_step": int(h2d_bytes * world), "same_calls_as_resident_run": bool(n_calls_e2e == calls),
                    "quality_in_place_tiles": int(e2e_in_place),
                    "quality_bits": sorted({int(t["quality"][0]) for t in tiles}),
                    "wire_layout": sorted({"compact" if t["layout"][0] else "full" for t in tiles}),
                    "note": "h2d_bytes_per_step = bytes ngsep_tile_submit copies (record headers + payload); the quality array (%.0f MB) stays in the "
                            "caller's page-locked buffers and k_genotype_columns reads what it needs in place over PCIe "
                            "(--quality-in-place)" % ((h2d_bytes - e2e_copied) / 1e6) if e2e_in_place else
                            "every host buffer of the step is copied host -> device, the same mode at every N; qualities travel as "
                            "quality_bits-bit codes (ngsep_tile.qual_bits: the simulated qualities fall into 4 classes the path can tell apart), "
                            "single-block records without their CIGAR word and low-quality mask (ngsep_tile.layout compact: rebuilt on the device "
                            "at staging, inside the timed region)"},
            "pipeline": pipeline,
            "gpu_launches": int(launches_all * args.steps),
            "clocks": clocks,
            "host": {"simulate_s": t_sim, "pack_s": t_pack},
        }
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(genome, batch, names)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle (C++ port of the reference's Java path) on the host cores
# ----------------------------------------------------------------------------------------------------------------------
def _sample_batch(batch, names, max_bp):
    """First sequences of the workload up to ~max_bp reference bases."""
    take, tot = [], 0
    for ci, l in enumerate(batch.contig_lengths):
        if tot and tot + l > max_bp:
            break
        take.append(ci); tot += l
    idx = np.nonzero(np.isin(batch.seq_idx, take))[0]
    return batch.select(idx), take, tot


def cpu_baseline(genome, batch, names, max_bp: int = 1_500_000):
    """Oracle, one thread, on a bounded sample of the same workload (about 10-30 s of CPU work)."""
    from ngsepcore_b200 import simulate
    from tests import oracle_util as ou
    sub, take, tot = _sample_batch(batch, names, max_bp)
    with tempfile.TemporaryDirectory() as d:
        fa, sam = os.path.join(d, "r.fa"), os.path.join(d, "r.sam")
        simulate.write_fasta(fa, genome)
        simulate.write_sam(sam, sub)
        _, _, gp, sec = ou.run_ssvd(ou.load_oracle(), fa, sam)
    return {"value": gp / sec, "unit": UNIT, "cores": 1, "kind": "port", "seconds": sec,
            "sample": f"first {len(take)} sequences of the workload ({tot} bp, {sub.n} records), C++ restatement of NGSEP 4.3.2 (no JVM in the image)"}


def _ref_write(job):
    """Worker: SAM text of one region's records (untimed setup; forked, the batch is inherited)."""
    from ngsepcore_b200 import simulate
    sam, keep = job
    simulate.write_sam(sam, _REF["batch"].select(keep))
    return sam


def _ref_worker(job):
    fa, sam, seq, first, last = job
    from tests import oracle_util as ou
    args = ("-querySeq", seq, "-first", str(first), "-last", str(last))
    _, _, gp, sec = ou.run_ssvd(ou.load_oracle(), fa, sam, args=args)
    return gp, sec


_REF = {}


def run_reference(args):
    """The reference's CPU implementation of the path on all host cores, over the WHOLE workload of the CUDA arm (same
    simulated genome and reads).  The reference's only parallel mode is region splitting (-querySeq/-first/-last, SURVEY.md
    fact 0.3): the genome is cut into about two regions per core, one process each.  Timed: SAM text parsing, admission, the
    per-position pileup walk with the realigner and both callers, VCF text — what `java -jar NGSEPcore.jar
    SingleSampleVariantsDetector` does between reading its input and writing its output."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])      # the checker only: no CUDA library is loaded in this arm
    from ngsepcore_b200 import simulate
    genome, batch, _ = _workload(0, args.depth, args.scale)
    names = list(genome)
    cores = max(1, min(os.cpu_count() or 1, 64))
    total = sum(batch.contig_lengths)
    sample_bp = min(total, args.reference_sample_bp)
    d = tempfile.mkdtemp(prefix="ngsep_ref_")
    fa = os.path.join(d, "r.fa")
    simulate.write_fasta(fa, genome)
    # about two regions per core, balanced by reference length; each worker reads only the records that can reach its region
    per = max(20000, -(-sample_bp // (2 * cores)))
    jobs, writes, taken, n_records = [], [], 0, 0
    for ci, L in enumerate(batch.contig_lengths):
        if taken >= sample_bp:
            break
        lo, hi = np.searchsorted(batch.seq_idx, ci, "left"), np.searchsorted(batch.seq_idx, ci, "right")
        first_arr = batch.first[lo:hi]
        for f in range(1, L + 1, per):
            l = min(L, f + per - 1)
            keep = lo + np.nonzero((first_arr >= f - 1000) & (first_arr <= l))[0]
            sam = os.path.join(d, f"{names[ci]}_{f}.sam")
            writes.append((sam, keep))
            jobs.append((fa, sam, names[ci], f, l))
            n_records += len(keep)
        taken += L
    _REF["batch"] = batch
    pool = mp.get_context("fork").Pool(cores)
    pool.map(_ref_write, writes, chunksize=1)
    times, positions = [], 0
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res = pool.map(_ref_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
            positions = sum(r[0] for r in res)
    pool.close()
    total_s = sum(times)
    value = positions * len(times) / total_s
    whole = taken >= total
    sample = (("the whole workload" if whole else f"the first {taken} bp of the workload") +
              f" ({n_records} records incl. lead-ins) cut into {len(jobs)} regions, oracle port (C++ restatement of NGSEP 4.3.2; no JVM in the image) on {cores} processes")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_s / len(times) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": WORKLOAD, "depth": args.depth, "genome_bp": int(total), "reads": int(batch.n), "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# BASELINE configs[2] and [4] with inputs generated on the device (ngsep_synth_*): no read crosses PCIe
# ----------------------------------------------------------------------------------------------------------------------
GRCH38_LIKE = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422, 135086622, 133275309,
               114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415]
SYNTH_CONFIGS = {
    "c3": dict(contigs=[64444167], depth=60.0, tile=2_000_000, gc=0.44, what="synthetic 64 Mb (chr20-sized) sequence x 60x (BASELINE configs[2])"),
    "c5": dict(contigs=GRCH38_LIKE, depth=30.0, tile=4_000_000, gc=0.41, what="synthetic 3.1 Gb human-scale genome (24 sequences) x 30x (BASELINE configs[4])"),
    "c1s": dict(contigs=[230218, 813184, 316620, 1531933, 576874, 270161, 1090940, 562643, 439888, 745751, 666816, 1078177, 924431, 784333, 1091291, 948066, 85779],
                depth=30.0, tile=2_000_000, gc=0.38, what="yeast-sized genome (17 sequences, 12.16 Mb) x 30x with device-generated reads (configs[1]'s shape without indels)"),
    "panel": dict(contigs=[2_000_000], depth=1000.0, tile=125_000, gc=0.41, what="2 Mb targeted panel x 1000x (BASELINE configs[4], depth stress)"),
}


def run_synth(args):
    """Strong scaling: the tiles of ONE genome are dealt round-robin to the ranks (region sharding, no collective); every rank
    generates its own tiles in HBM, walks them in one launch per step and keeps its calls; value = positions of all ranks / max time."""
    rank, world, local, dist = _dist_setup(args.gpus)
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if dist is not None:
        dist.barrier()
    from ngsepcore_b200 import capi
    cfg = SYNTH_CONFIGS[args.config]
    device = local if world > 1 else 0
    ctx = capi.Context(device=device)
    tiles = [(ci, f, min(L, f + cfg["tile"] - 1)) for ci, L in enumerate(cfg["contigs"]) for f in range(1, L + 1, cfg["tile"])]
    mine = tiles[rank::world]
    t0 = time.time()
    for ci in sorted({t[0] for t in mine}):
        ctx.synth_reference(ci, cfg["contigs"][ci], seed=1031, gc=cfg["gc"])
    slots = [ctx.synth_tile(ci, f, l, depth=cfg["depth"], seed=3033) for ci, f, l in mine]
    gen_s = time.time() - t0
    shapes = [ctx.resident_download(s, data=False) for s in slots]
    n_reads = sum(t["n_reads"] for t in shapes)
    in_bytes = sum(t["n_reads"] * 16 + t["n_words"] * 4 + t["n_qual"] for t in shapes)
    for _ in range(max(3, args.warmup)):
        ctx.resident_run(-1)
    st = ctx.resident_stats()
    sampler = _ClockSampler(device) if rank == 0 else None
    if dist is not None:
        dist.barrier()
    step_ms, pileup_ms, kernels_ms = [], 0.0, 0.0
    for _ in range(args.steps):
        ctx.timer_start()
        ctx.resident_run(-1)
        step_ms.append(ctx.timer_stop())
        rs = ctx.resident_stats()
        pileup_ms += rs.pileup_ms
        kernels_ms += rs.kernel_ms
    if dist is not None:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    L = 100
    alg = n_reads * (16 + 4 + (L + 3) // 4 + L) + sum(l - f + 1 for _, f, l in mine) + 64 * st.calls
    (covered, bases, calls, reads, alg_all, in_all), (total_ms, pileup_ms, kernels_ms, gen_s) = reduce_over_ranks(
        dist, "cuda", sums=[st.covered_positions, st.aligned_bases, st.calls, n_reads, alg, in_bytes], maxima=[float(sum(step_ms)), pileup_ms, kernels_ms, gen_s])
    if rank == 0:
        peak, peak_src = _peaks()
        ms = total_ms / args.steps
        per_gpu_alg = alg_all / world
        line = {"metric": METRIC, "value": covered / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic (generated on the device)",
                "config": {"workload": cfg["what"] + ", reads generated in HBM (ngsep_synth_tile), tiles of %d positions dealt round-robin to the GPUs" % cfg["tile"],
                           "depth": cfg["depth"], "genome_bp": int(sum(cfg["contigs"])), "reads": int(reads), "tiles": len(tiles),
                           "genotyped_positions": int(covered), "calls": int(calls), "resident_input_bytes": int(in_all),
                           "l2": "inputs larger than L2 (%.1f GB of packed tiles per GPU resident in HBM, walked every step)" % (in_all / world / 1e9),
                           "parallelism": f"region tiles x{world}, no collective"},
                "aligned_bases_per_sec": bases / (ms / 1e3),
                "roofline": {"bound": "hbm", "achieved": per_gpu_alg / (kernels_ms / args.steps / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": per_gpu_alg / (kernels_ms / args.steps / 1e3) / 1e9 / peak, "traffic": None,
                             "kernel": "the device step: k_pileup_windows + k_genotype_columns (per GPU, slowest rank)", "algorithmic_bytes_per_step_per_gpu": per_gpu_alg,
                             "kernels_ms_per_step": kernels_ms / args.steps, "k_pileup_windows_ms_per_step": pileup_ms / args.steps, "peak_source": peak_src},
                "e2e": None, "gpu_launches": int(2 * args.steps * world), "clocks": clocks, "host": {"generate_s": gen_s}}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ngsepcore_b200", choices=["ngsepcore_b200", "reference"])
    ap.add_argument("--depth", type=float, default=30.0)
    ap.add_argument("--scale", type=float, default=1.0, help="scale the sequence lengths (debugging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the informational whole-detector run")
    ap.add_argument("--quality-in-place", action="store_true", help="e2e: leave page-locked quality buffers on the host (NGSEP_F_QUALITY_IN_PLACE)")
    ap.add_argument("--reference-sample-bp", type=int, default=1 << 40, help="reference arm: reference bases per step (default: the whole workload)")
    ap.add_argument("--config", default="c1", choices=["c1"] + sorted(SYNTH_CONFIGS),
                    help="c1 (default): BASELINE configs[1], host-simulated yeast 30x; c3 / c5 / panel: configs[2] / [4] with device-generated inputs (strong scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config != "c1":
        run_synth(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
