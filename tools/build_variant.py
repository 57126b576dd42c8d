"""Build a tuning variant of libvlo.so with extra nvcc flags: python tools/build_variant.py name -DSEG_PTS=16 ...
-> vil_sensor_fusion_b200/lib/variants/libvlo_<name>.so (load it with VLO_LIB_PATH)."""
import concurrent.futures as cf, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vil_sensor_fusion_b200 import build as b
name, extra = sys.argv[1], sys.argv[2:]
obj_dir = os.path.join(b.OBJDIR, "variant_" + name); os.makedirs(obj_dir, exist_ok=True)
out_dir = os.path.join(b.LIBDIR, "variants"); os.makedirs(out_dir, exist_ok=True)
CSRC = os.environ.get("VLO_CSRC", b.CSRC)
srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu") and not f.startswith("synth_"))
def one(src):
    obj = os.path.join(obj_dir, src[:-3] + ".o")
    r = subprocess.run([b.NVCC] + b.NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
    if r.returncode: raise RuntimeError(r.stderr)
    return obj
with cf.ThreadPoolExecutor(8) as ex: objs = list(ex.map(one, srcs))
out = os.path.join(out_dir, "libvlo_%s.so" % name)
subprocess.run([b.NVCC, "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"], check=True)
print(out)
