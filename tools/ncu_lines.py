"""Top CUDA source lines of an .ncu-rep by warp-stall samples (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
def main(path, top=30):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None; cur_file = ""; agg = {}
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
        if len(r) > 6 and r[0] == "Line No": hdr = {n: i for i, n in enumerate(r)}; hdr_list = r; continue
        if hdr is None or len(r) < 10 or r[0] == "": continue
        def f(name):
            # first occurrence index
            if name not in hdr_list: return 0.0
            i = hdr_list.index(name)
            try: return float(r[i].replace(",", ""))
            except Exception: return 0.0
        key = (cur_file, r[0])
        a = agg.setdefault(key, [0, 0, 0, 0, 0, 0, 0, 0, r[1].strip()])
        a[0] += f("# Samples"); a[1] += f("Instructions Executed"); a[2] += f("stall_barrier"); a[3] += f("stall_short_sb")
        a[4] += f("stall_long_sb"); a[5] += f("stall_wait"); a[6] += f("stall_mio"); a[7] += f("L1 Wavefronts Shared Excessive")
    tot_s = sum(a[0] for a in agg.values()) or 1; tot_i = sum(a[1] for a in agg.values()) or 1
    print("total samples %d, total warp-instructions %d" % (tot_s, tot_i))
    for (fn, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%5.1f%% smp %5.1f%% inst | bar %4d ssb %4d lsb %4d wait %4d mio %4d xs_smem %7d | %s:%s %s" %
              (100 * a[0] / tot_s, 100 * a[1] / tot_i, a[2], a[3], a[4], a[5], a[6], a[7], fn, ln, a[8][:90]))
if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
