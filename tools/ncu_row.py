"""One table row per ncu report: the figures profiles/README.md tabulates (duration, DRAM bytes, warp-instructions,
issue rate, active threads per instruction, warps per SM, ALU pipe, registers, top stall reasons).

usage: python tools/ncu_row.py report1.ncu-rep [report2.ncu-rep ...]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("ms", "gpu__time_duration.sum"),
    ("dram_rd", "dram__bytes_read.sum"),
    ("dram_wr", "dram__bytes_write.sum"),
    ("warp_inst", "smsp__inst_executed.sum"),
    ("thr/inst", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("issue/sched", "smsp__issue_active.avg.per_cycle_active"),
    ("warps_active_%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("alu_pipe_%", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
]


def row(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return None
    h, u, v = rows[0], rows[1], rows[2]
    d = {"kernel": v[h.index("Kernel Name")].split("(")[0] if "Kernel Name" in h else "?"}
    for name, key in KEYS:
        if key in h:
            i = h.index(key)
            d[name] = f"{v[i]} {u[i]}".strip()
    stalls = []
    for i, k in enumerate(h):
        if "average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            stalls.append((float(v[i]), k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
    d["stalls/issue"] = " ".join(f"{n}={x:.2f}" for x, n in sorted(stalls, reverse=True)[:5])
    return d


def main():
    for rep in sys.argv[1:]:
        d = row(rep)
        print(rep)
        if d:
            for k, v in d.items():
                print(f"  {k:16s} {v}")


if __name__ == "__main__":
    main()
