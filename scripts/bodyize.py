"""One-off source transform used while introducing batched launches: turn `__global__ void K(args) {body}` into
`__device__ __forceinline__ void K_body(args) {body}` + `__global__ void K(args) { K_body(names...); }` so that the generic
batched trampoline (common.cuh, svae_multi_kernel) can run the unchanged body.  usage: bodyize.py file.cu K1 K2 ..."""
import re
import sys


def split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "(<[":
            depth += 1
        elif ch in ")>]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def main():
    path, names = sys.argv[1], sys.argv[2:]
    src = open(path).read()
    for k in names:
        m = re.search(r"(template\s*<[^>]*>\s*)?__global__\s+void\s+(__launch_bounds__\([^)]*\)\s*)?" + re.escape(k) + r"\s*\(", src)
        assert m, k
        tmpl, lb = m.group(1) or "", m.group(2) or ""
        i = m.end()
        depth, j = 1, i
        while depth:
            depth += {"(": 1, ")": -1}.get(src[j], 0)
            j += 1
        args = src[i:j - 1]
        names_ = [re.search(r"([A-Za-z_][A-Za-z0-9_]*)\s*$", a).group(1) for a in split_args(args)]
        b0 = src.index("{", j)
        depth, e = 1, b0 + 1
        while depth:
            depth += {"{": 1, "}": -1}.get(src[e], 0)
            e += 1
        body = src[b0:e]
        targs = ""
        if tmpl:
            tn = [re.search(r"([A-Za-z_][A-Za-z0-9_]*)\s*$", a).group(1) for a in split_args(re.search(r"<(.*)>", tmpl, re.S).group(1))]
            targs = "<" + ", ".join(tn) + ">"
        new = (tmpl + "__device__ __forceinline__ void " + k + "_body(" + args + ") " + body + "\n" + tmpl + "__global__ void " + lb + k +
               "(" + args + ") { " + k + "_body" + targs + "(" + ", ".join(names_) + "); }")
        src = src[:m.start()] + new + src[e:]
    open(path, "w").write(src)


main()
