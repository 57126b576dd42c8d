"""profiles/dram_traffic.json from `ncu --set full` reports: dram__bytes_read.sum + dram__bytes_write.sum per launch of the
named kernels, stamped with the sha256 of the kernel's source file so that bench.py refuses a capture of older code
(roofline.traffic is then null with the reason).  Usage: python tools/ncu_traffic.py <profile dir under profiles/ or gpurun_out/> ...
A stage of several kernels (k0_organise = k0_classify + k0_scatter, k1_extract = k1_extract + k1c_lessflat) sums its kernels."""
import csv, hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGES = {   # bench stage / kernel -> (source file, [kernel name prefixes that make it up])
    "k5_assoc": ("k5_mapping.cu", ["k5_assoc"]), "k5_lin": ("k5_mapping.cu", ["k5_lin"]),
    "k1_extract": ("k1_extract.cu", ["k1_extract", "k1c_lessflat"]), "k0_organise": ("k0_organise.cu", ["k0_classify", "k0_scatter"]),
    "k3_assoc": ("k3_odometry.cu", ["k3_assoc"]), "k3_gn": ("k3_odometry.cu", ["k3_gn"]), "k6_imu": ("k6_imu.cu", ["k6_imu"]),
    "k7_stack_ds": ("k7_map.cu", ["k7_ds_bin", "k7_ds_emit"]),
}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

def dram_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return None
    hdr, units, r = rows[0], rows[1], rows[2]
    def val(k):
        i = hdr.index(k)
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1)
    return r[hdr.index("Kernel Name")], val("dram__bytes_read.sum"), val("dram__bytes_write.sum"), float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))

def main(dirs):
    path = os.path.join(ROOT, "profiles", "dram_traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    for d in dirs:
        reps = {f[len("full_"):-len(".ncu-rep")]: os.path.join(d, f) for f in os.listdir(d) if f.startswith("full_") and f.endswith(".ncu-rep")}
        for stage, (src, kernels) in STAGES.items():
            got = [(k, dram_of(reps[k])) for k in kernels if k in reps]
            if len(got) != len(kernels) or any(g[1] is None for g in got):
                continue
            rd, wr = sum(g[1][1] for g in got), sum(g[1][2] for g in got)
            table[stage] = {"dram_bytes_per_launch": int(rd + wr), "read": int(rd), "write": int(wr), "kernels": [g[1][0] for g in got],
                            "source": src, "source_sha256": hashlib.sha256(open(os.path.join(ROOT, "vil_sensor_fusion_b200", "csrc", src), "rb").read()).hexdigest(),
                            "profile": os.path.relpath(d, ROOT) + " (ncu --set full --clock-control none, one launch each)"}
            print(stage, table[stage]["dram_bytes_per_launch"], table[stage]["kernels"])
    json.dump(table, open(path, "w"), indent=1)

if __name__ == "__main__":
    main(sys.argv[1:])
