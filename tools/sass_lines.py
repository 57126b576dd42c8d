"""Static SASS instruction count per source function-level bucket of one kernel (code-size / i-cache budget).
Usage: python tools/sass_lines.py obj.o kernel_substring [depth]
Attributes each instruction to the chain of inlined call sites (innermost first) cut at `depth` frames."""
import collections, os, re, subprocess, sys, tempfile
def main(obj, kern, depth=1):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    out = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    fn = ""; chain = []; new_block = True; cnt = collections.Counter(); total = 0
    for line in out.splitlines():
        if ".text." in line and line.strip().startswith(".section"):
            fn = line
        m = re.search(r'//## File "([^"]*)", line (\d+)', line)
        if m:
            if new_block: chain = []; new_block = False
            chain.append("%s:%s" % (os.path.basename(m.group(1)), m.group(2)))
            continue
        if re.match(r"^\s+/\*[0-9a-f]+\*/\s", line):
            new_block = True
            if kern in fn:
                total += 1
                # chain is innermost .. outermost; report the outermost `depth` frames below the kernel body
                key = " < ".join(chain[-depth - 1:-1][::-1]) if len(chain) > 1 else (chain[0] if chain else "?")
                cnt[key] += 1
    print("kernel %s: %d instructions, %.1f KB" % (kern, total, total * 16 / 1024))
    for k, v in cnt.most_common(40):
        print("%6d  %s" % (v, k))
if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 1)
