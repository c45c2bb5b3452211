#!/usr/bin/env python
"""Where a divergent kernel spends its warp instructions: contiguous SASS regions of one profiled launch with their
share of the executed warp instructions, executions per warp and ACTIVE LANES, attributed to the source line of the
kernel body they were inlined into (nvdisasm -gi on the library's cubin).

    python tools/ncu_regions.py REP.ncu-rep LIB.so KERNEL_SUBSTRING [top]

KERNEL_SUBSTRING selects the mangled kernel in the cubin (e.g. env_step_kernelIfLi0EfLb1).  Development aid."""
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_source import page


def sass_rows(rep):
    rows = page(rep, "sass")
    hdr = rows[1]
    ia, isrc, iex, ithr = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "Thread Instructions Executed"))
    out, base = [], None
    for r in rows[2:]:
        if len(r) <= ithr or not r[ia].startswith("0x"):
            continue
        a = int(r[ia], 16)
        base = a if base is None else base
        ex, th = int(r[iex]), int(r[ithr])
        out.append((a - base, ex, th / ex if ex else 0.0, r[isrc].strip()))
    return out


def line_map(lib, kernel):
    """offset -> [innermost ... outermost] 'file:line' chain"""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout
        m = re.search(r"^\s*\.section\s+\.text\.(\S*%s\S*?),.*$" % re.escape(kernel), txt, re.M)
        if not m:
            continue
        body = txt[m.end():]
        nxt = re.search(r"^\s*\.section\s", body, re.M)
        body = body[:nxt.start()] if nxt else body
        out, chain, last = {}, [], []
        for l in body.split("\n"):
            s = l.strip()
            if s.startswith("//##"):
                r = re.search(r'File ".*/([^/"]+)", line (\d+)', s)
                if r:
                    chain.append(f"{r.group(1)}:{r.group(2)}")
            else:
                mm = re.match(r"/\*([0-9a-f]{4,6})\*/", s)
                if mm:
                    last = chain or last
                    out[int(mm.group(1), 16)] = last
                    chain = []
        return out
    return {}


def main():
    rep, lib, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    rows = sass_rows(rep)
    lm = line_map(lib, kernel)
    tot = sum(r[1] for r in rows)
    warps = max(r[1] for r in rows[:4]) or 1                     # the prologue runs once per warp
    print(f"# {tot} warp instructions, {warps} warps, {tot / warps:.0f} per warp, "
          f"{sum(r[1] * r[2] for r in rows) / tot:.1f} lanes on average")
    for lo, hi in ((0, 4), (4, 8), (8, 12), (12, 16), (16, 24), (24, 33)):
        s = sum(r[1] for r in rows if lo <= r[2] < hi and r[1])
        print(f"#   lanes [{lo:2d},{hi:2d}): {100.0 * s / tot:5.1f} % of the warp instructions")
    regs, cur = [], None
    for off, ex, lanes, _ in rows:
        if ex == 0:
            if cur:
                regs.append(cur); cur = None
            continue
        if cur and abs(cur[4] - lanes) < 0.35 and abs(cur[5] - ex) / ex < 0.6:
            cur[1] = off; cur[2] += ex; cur[3] += 1
        else:
            if cur:
                regs.append(cur)
            cur = [off, off, ex, 1, lanes, ex]
    if cur:
        regs.append(cur)
    for r in sorted(regs, key=lambda r: -r[2])[:top]:
        where = collections.Counter()
        inner = collections.Counter()
        for off in range(r[0], r[1] + 16, 16):
            ch = lm.get(off)
            if ch:
                where[ch[-1]] += 1
                inner[" <- ".join(ch[:3])] += 1
        w = ", ".join(k for k, _ in where.most_common(2))
        i = inner.most_common(1)[0][0] if inner else "?"
        print(f"{r[0]:05x}-{r[1]:05x} {100.0 * r[2] / tot:5.2f} %  {r[3]:4d} instr  x{r[5] / warps:4.2f}/warp  lanes {r[4]:4.1f}  in {w}  ({i})")


if __name__ == "__main__":
    main()
