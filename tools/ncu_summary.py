"""Summarise an .ncu-rep (ncu --set full): per-kernel duration, DRAM traffic, throughput, occupancy and
the top warp-stall reasons.  Usage: python tools/ncu_summary.py file.ncu-rep [> profiles/xyz.md]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %peak"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 cyc %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occ %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "occ lim regs"),
    ("launch__occupancy_limit_shared_mem", "occ lim smem"),
    ("smsp__inst_executed.sum", "warp insts"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
]


def main(fn):
    raw = subprocess.run(["ncu", "-i", fn, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    print(f"# {fn}\n")
    for r in rows[2:]:
        print(f"## {r[idx['Kernel Name']][:110]}")
        for k, label in KEYS:
            if k in idx:
                print(f"- {label}: {r[idx[k]]} {units[idx[k]]}")
        st = sorted(((float(r[idx[h]] or 0), h) for h in stall), reverse=True)[:5]
        print("- top stalls (warps per issue): " + ", ".join(
            f"{h.split('stalled_')[1].split('_per_issue')[0]}={v:.2f}" for v, h in st))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
