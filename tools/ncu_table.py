"""One markdown table row per .ncu-rep: the numbers DESIGN.md / profiles/SUMMARY.md quote.
Usage: python tools/ncu_table.py label=path.ncu-rep ...   (sectors per request and hit rates for the map-query kernels)"""
import csv, subprocess, sys
M = {
    "t_us": "gpu__time_duration.sum",
    "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_hit": "lts__t_sector_hit_rate.pct", "l1_hit": "l1tex__t_sector_hit_rate.pct",
    "sec_req": "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio",
    "gld_sectors": "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "gld_req": "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "lanes": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "warps": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "inst": "smsp__inst_executed.sum", "regs": "launch__registers_per_thread",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}
STALL = "smsp__average_warps_issue_stalled_"
def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    def get(name):
        if name not in hdr: return None
        i = hdr.index(name)
        try: v = float(r[i].replace(",", ""))
        except ValueError: return None
        return v * UNIT.get(units[i], 1.0)
    d = {k: get(v) for k, v in M.items()}
    st = [(float(r[i].replace(",", "")), h[len(STALL):-len("_per_issue_active.ratio")]) for i, h in enumerate(hdr)
          if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i]]
    d["stalls"] = ", ".join("%s %.1f" % (n, v) for v, n in sorted(st, reverse=True)[:3] if n != "selected")
    d["kernel"] = r[hdr.index("Kernel Name")].split("(")[0]
    d["grid"] = r[hdr.index("Grid Size")]
    return d
def main(args):
    print("| kernel (capture) | time µs | DRAM rd+wr MB | DRAM GB/s (% of peak) | L2 hit % | L1 hit % | sectors / global-load request | issue active % | lanes active /32 | warps active % | regs | top stalls (cycles per issue) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for a in args:
        label, path = a.split("=", 1)
        d = load(path)
        mb = ((d["rd"] or 0) + (d["wr"] or 0)) / 1e6
        gbs = mb / 1e3 / (d["t_us"] * 1e-6) if d["t_us"] else 0
        sr = d["sec_req"] if d["sec_req"] is not None else ((d["gld_sectors"] / d["gld_req"]) if d["gld_sectors"] and d["gld_req"] else None)
        print("| `%s` (%s) | %.1f | %.1f | %.0f (%.1f) | %.1f | %.1f | %s | %.1f | %.1f | %.1f | %d | %s |" % (
            d["kernel"], label, d["t_us"], mb, gbs, d["dram_pct"] or 0, d["l2_hit"] or 0, d["l1_hit"] or 0,
            ("%.2f" % sr) if sr is not None else "n/a", d["issue"] or 0, d["lanes"] or 0, d["warps"] or 0, int(d["regs"] or 0), d["stalls"]))
if __name__ == "__main__":
    main(sys.argv[1:])
