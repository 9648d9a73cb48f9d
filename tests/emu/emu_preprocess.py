"""tests/emu/emu_preprocess.py SRC.cu OUT.cpp -- TEST INFRASTRUCTURE ONLY.
Rewrites the two CUDA constructs g++ cannot parse: `kernel<<<grid, block, smem, stream>>>(args);` becomes
`EMU_LAUNCH(kernel, grid, block, smem, args);` and `extern __shared__ double name[];` becomes a pointer to the emulation's
dynamic shared-memory buffer.  Everything else is handled by macros in include/cuda_runtime.h."""
import re
import sys


def split_top(s):
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


def rewrite_launches(line):
    while True:
        pos = line.find("<<<")
        if pos < 0:
            return line
        # kernel expression before <<<: an identifier, optionally with template arguments
        e = pos
        while e > 0 and line[e - 1] == " ":
            e -= 1
        st = e
        if line[st - 1] == ">":
            depth = 0
            while True:
                st -= 1
                if line[st] == ">":
                    depth += 1
                elif line[st] == "<":
                    depth -= 1
                    if depth == 0:
                        break
        while st > 0 and (line[st - 1].isalnum() or line[st - 1] == "_"):
            st -= 1
        kern = line[st:e]

        class M:                                   # same interface as the former regex match
            def start(self): return st
            def group(self, n): return "(" + kern + ")" if "<" in kern else kern
        m = M()
        a = pos + 3
        b = line.index(">>>", a)
        cfg = split_top(line[a:b])
        assert len(cfg) in (2, 3, 4), line
        while len(cfg) < 3:
            cfg.append("0")
        p = line.index("(", b)
        depth, q = 0, p
        while True:
            if line[q] == "(":
                depth += 1
            elif line[q] == ")":
                depth -= 1
                if depth == 0:
                    break
            q += 1
        args = line[p + 1:q].strip()
        line = line[:m.start()] + "EMU_LAUNCH(%s, %s, %s, %s%s)" % (m.group(1), cfg[0], cfg[1], cfg[2], (", " + args) if args else "") + line[q + 1:]


src, dst = sys.argv[1:3]
out = []
for line in open(src):
    if "<<<" in line and not line.lstrip().startswith("//"):
        line = rewrite_launches(line)
    line = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?double\s+(\w+)\s*\[\s*\]\s*;", r"double* \1 = (double*)emu::dyn_smem();", line)
    line = re.sub(r'#include\s+"(\w+)\.cu"', r'#include "\1.cpp"', line)      # a .cu that includes another .cu: use its rewritten copy
    out.append(line)
open(dst, "w").write('#line 1 "%s"\n' % src + "".join(out))
