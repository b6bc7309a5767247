"""Print the metrics of an ncu report that the kernel tables in profiles/README.md quote.

usage: python tools/ncu_keys.py <report.ncu-rep> [more substrings to match]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__warps_eligible.avg.per_cycle_active", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct", "smsp__sass_average_data_bytes_per_sector_mem_global_op_st.pct",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    H, U, V = rows[0], rows[1], rows[2]
    d = {h: (U[i], V[i]) for i, h in enumerate(H)}
    for k in KEYS:
        if k in d:
            print(f"{k:75s} {d[k][1]:>18s} {d[k][0]}")
    for h in H:
        if "issue_stalled" in h and "per_issue_active" in h:
            v = float(d[h][1])
            if v > 0.04:
                print(f"stall {h.split('stalled_')[1].split('_per')[0]:30s} {v:.3f}")
        elif any(e in h for e in extra):
            print(f"{h:75s} {d[h][1]:>18s} {d[h][0]}")


if __name__ == "__main__":
    main()
