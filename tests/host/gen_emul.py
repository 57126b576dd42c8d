"""TEST INFRASTRUCTURE: rewrite `kernel<<<grid, block, smem, stream>>>(args)` into `emu_launch(kernel, dim3(grid), dim3(block), args)`
so that a .cu file of csrc/ -- kernels AND its real host launcher -- compiles with g++ against tests/host/cuda_emul.h."""
import re
import sys


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite(src: str) -> str:
    # cudaLaunchCooperativeKernel((void *)kernel, grid, block, args, smem, stream) -> emu_launch_coop(kernel, grid, block, args)
    def coop(m):
        a = _split_top(m.group(1))
        return "emu_launch_coop(%s, %s, %s, %s)" % (re.sub(r"^\(void \*\)", "", a[0]), a[1], a[2], a[3])
    src = re.sub(r"cudaLaunchCooperativeKernel\(((?:[^()]|\([^()]*\))*)\)", coop, src)
    out, i = "", 0
    while True:
        j = src.find("<<<", i)
        if j < 0:
            return out + src[i:]
        k = src.find(">>>", j)
        # kernel name (optionally with template arguments) directly before <<<
        m = re.search(r"([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*$", src[i:j])
        name = m.group(1)
        cfg = _split_top(src[j + 3:k])
        assert src[k + 3] == "(", src[k:k + 20]
        out += src[i:i + m.start(1)] + "emu_launch(%s, dim3(%s), dim3(%s), " % (name, cfg[0], cfg[1])
        i = k + 4


if __name__ == "__main__":
    open(sys.argv[2], "w").write(rewrite(open(sys.argv[1]).read()))
