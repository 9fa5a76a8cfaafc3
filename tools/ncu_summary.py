"""Extract the judged metrics of one kernel from an .ncu-rep (read here, no GPU needed).
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_atom.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
    "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none: `{path}`\n")
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        print(f"## {name}\n\n| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {vals[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
