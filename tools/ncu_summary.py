"""Summarise an .ncu-rep (raw page CSV) into the handful of numbers DESIGN/profiles quote."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio", "lts__t_sectors_op_read.sum", "lts__t_bytes.sum",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__cycles_elapsed.max", "local_load_requests", "smsp__inst_executed_op_local_ld.sum"]
STALL = "smsp__average_warps_issue_stalled_"
def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== kernel", r[hdr.index("Kernel Name")], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for k in KEYS:
            if k in hdr:
                print("  %-75s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        st = [(float(r[i].replace(",", "")), h) for i, h in enumerate(hdr) if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i]]
        for v, h in sorted(st, reverse=True)[:8]:
            print("  stall %-60s %.2f" % (h[len(STALL):-len("_per_issue_active.ratio")], v))
if __name__ == "__main__":
    main(sys.argv[1])
