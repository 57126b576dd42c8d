"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares."""
import csv, sys, collections
def main(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = None; agg = collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r and "Metric Value" in r:
            hdr = {n: i for i, n in enumerate(r)}; continue
        if hdr is None or len(r) <= hdr["Metric Value"]: continue
        if r[hdr["Metric Name"]] != "gpu__time_duration.sum": continue
        name = r[hdr["Kernel Name"]].split("(")[0]
        v = float(r[hdr["Metric Value"]].replace(",", "")); u = r[hdr["Metric Unit"]]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values()) or 1.0
    print("%-28s %8s %12s %10s %7s" % ("kernel", "launches", "total_us", "avg_us", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-28s %8d %12.1f %10.2f %6.1f%%" % (k, n, t, t / n, 100 * t / tot))
    print("%-28s %8d %12.1f" % ("TOTAL", sum(a[0] for a in agg.values()), tot))
if __name__ == "__main__":
    main(sys.argv[1])
