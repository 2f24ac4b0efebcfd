#!/usr/bin/env python
"""Summary of one kernel launch from an .ncu-rep (ncu --set full): selected raw metrics -> JSON, plus the per-region and
per-line instruction / stall-sample shares as text.   ncu_summary.py REPORT.ncu-rep OUT_PREFIX "source note"
Writes OUT_PREFIX.json and OUT_PREFIX_source.txt (what profiles/ holds)."""
import csv, json, os, subprocess, sys, tempfile

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sectors.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps"]


def main():
    rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    here = os.path.dirname(os.path.abspath(__file__))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = {"value": vals[i], "unit": units[i]}
    if "Kernel Name" in hdr:
        d["kernel"] = vals[hdr.index("Kernel Name")]
    d["_source"] = note
    json.dump(d, open(out + ".json", "w"), indent=1)
    with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as f:
        f.write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout)
        src_csv = f.name
    src = os.path.join(os.path.dirname(here), "ngsepcore_b200", "csrc", "pileup_kernels.cu")
    a = subprocess.run([sys.executable, os.path.join(here, "ncu_regions.py"), src_csv, src, "16"], capture_output=True, text=True).stdout
    b = subprocess.run([sys.executable, os.path.join(here, "ncu_lines.py"), src_csv, "30"], capture_output=True, text=True).stdout
    open(out + "_source.txt", "w").write(a + b)
    os.unlink(src_csv)
    print(json.dumps({k: v["value"] for k, v in d.items() if isinstance(v, dict)}, indent=0)[:1500])


if __name__ == "__main__":
    main()
