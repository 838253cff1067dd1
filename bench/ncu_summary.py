"""Extract the per-kernel figures quoted in DESIGN.md / bench.py from an ncu report
(`ncu --set full ... -o X`): `python bench/ncu_summary.py X.ncu-rep key` merges them under `key`
into profiles/r01_ncu_kernel_summaries.json (values converted to base units: bytes, ms, %)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12,
         "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}


def main():
    rep, key = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    rec = {"Kernel Name": vals[hdr.index("Kernel Name")], "source": os.path.basename(rep)}
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP:
            try:
                x = float(v.replace(",", ""))
            except ValueError:
                continue
            if u in SCALE and ("bytes" in h or "time_duration" in h):
                x *= SCALE[u]
            rec[h] = x
    if "dram__bytes_read.sum" in rec:
        rec["dram_bytes_total"] = rec["dram__bytes_read.sum"] + rec["dram__bytes_write.sum"]
        rec["units"] = "bytes, ms, percent, ratios"
    path = os.path.join(ROOT, "profiles", "r01_ncu_kernel_summaries.json")
    allk = json.load(open(path)) if os.path.exists(path) else {}
    allk[key] = rec
    json.dump(allk, open(path, "w"), indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
