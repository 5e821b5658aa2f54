#!/usr/bin/env python3
"""ncu raw page (ncu -i X.ncu-rep --page raw --csv) -> the per-kernel summary CSVs kept under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/summarize_ncu.py /tmp/raw.csv profiles/r02_ncu_full_<what>_summary.csv "<command line that was profiled>"
"""
import csv
import sys

KEYS = """Kernel Name
gpu__time_duration.sum
dram__bytes_read.sum
dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
smsp__issue_active.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__warps_active.avg.pct_of_peak_sustained_active
l1tex__t_sector_hit_rate.pct
lts__t_sector_hit_rate.pct
launch__registers_per_thread
launch__shared_mem_per_block_dynamic
launch__grid_size
launch__block_size
launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem
sm__cycles_elapsed.max
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio""".split("\n")


def main():
    raw, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w", newline="") as fh:
        if note:
            fh.write(f"# {note}\n")
        w = csv.writer(fh)
        short = [r[hdr.index("Kernel Name")].replace("void ", "").split("(")[0] for r in data]
        w.writerow(["metric", "unit"] + short)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                w.writerow([k, units[i]] + [r[i] for r in data])


if __name__ == "__main__":
    main()
