"""Dev tool: print the key metrics of .ncu-rep files (run here, no GPU needed)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_tma_ld.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max"]
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "wait", "not_selected", "math_pipe_throttle", "lg_throttle", "mio_throttle",
          "branch_resolving", "dispatch_stall", "no_instruction", "membar", "sleeping", "drain", "imc_miss", "tex_throttle", "selected"]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units, vals = rows[0], rows[1], rows[2]
        print("==", path, "|", vals[hdr.index("Kernel Name")][:90])
        for w in WANT:
            if w in hdr:
                print(f"  {w:70s} {vals[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
        st = []
        for s in STALLS:
            k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if k in hdr:
                st.append((float(vals[hdr.index(k)]), s))
        print("  stalls per issue:", ", ".join(f"{s} {v:.2f}" for v, s in sorted(st, reverse=True)[:8]))


if __name__ == "__main__":
    main()
