"""One-screen summary of a bench.py JSON line."""
import json, sys
d = json.load(open(sys.argv[1]))
print("value %.1f %s  ms/step %.3f  e2e %.1f (ceiling %s GB/s, frac %s)  launches %d" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"],
      d["e2e"].get("pcie_h2d_gbs_all_ranks_concurrently"), d["e2e"].get("frac_of_pcie"), d["gpu_launches"]))
print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], d["roofline"].get("traffic_source"))
for k, v in d["stages"].items():
    print("  %-14s total %8.3f ms  avg %8.4f ms  %7.1f GB/s  frac %.4f" % (k, v["ms_total"], v["avg_ms"], v["achieved_gbs"], v["frac"]))
print("  iters", d["config"]["mean_gn_iterations"], d["config"].get("gn_iterations_hist"), "corr", d["mean_corr"], "ok", d["ok_registrations"], "/", d.get("registrations"))
print("latency", d["latency"])
for leg in ("whole_bag_pairs", "imu_batch", "vlp16_online", "cpu_baseline"):
    if d.get(leg):
        print(leg, json.dumps(d[leg])[:1800])
print("clocks", d["clocks"])
