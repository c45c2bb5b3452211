#!/usr/bin/env python
"""Per-source-line and per-SASS-instruction views of one profiled kernel launch in an .ncu-rep (reads with `ncu -i`).

    python tools/ncu_source.py REP lines [N]     # top N CUDA source lines by executed warp instructions (+ samples)
    python tools/ncu_source.py REP sass          # every SASS instruction: offset, warp-instr executed, lanes, samples
"""
import csv
import io
import subprocess
import sys


def page(rep, view):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", view, "--csv"],
                         capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def sass(rep):
    rows = page(rep, "sass")
    hdr = rows[1]
    ia, isrc, ismp, iex, ithr = (hdr.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed",
                                                         "Thread Instructions Executed"))
    base = None
    tot_i = tot_s = 0
    out = []
    for r in rows[2:]:
        if len(r) <= ithr or not r[ia].startswith("0x"):
            continue
        a = int(r[ia], 16)
        base = a if base is None else base
        ex, th, smp = int(r[iex]), int(r[ithr]), int(r[ismp])
        tot_i += ex; tot_s += smp
        out.append((a - base, ex, th / ex if ex else 0.0, smp, r[isrc].strip()))
    print(f"# total warp instructions {tot_i}, samples {tot_s}")
    for off, ex, lanes, smp, txt in out:
        print(f"{off:05x} {100.0 * ex / tot_i:6.3f}%i {100.0 * smp / max(tot_s, 1):6.3f}%s {lanes:5.1f} {txt}")


def lines(rep, top):
    rows = page(rep, "cuda,sass")
    recs, fname, hdr = [], "", None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit() and len(r) > 8:
            ismp, iex, ithr = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
            try:
                recs.append((int(r[iex]), int(r[ithr]), int(r[ismp]), fname, int(r[0]), r[1].strip()))
            except ValueError:
                pass
    ti = sum(x[0] for x in recs); ts = sum(x[2] for x in recs)
    print(f"# warp instructions {ti}, thread instructions {sum(x[1] for x in recs)}, samples {ts}")
    for ex, th, smp, f, ln, src in sorted(recs, reverse=True)[:top]:
        print(f"{100.0 * ex / ti:5.1f}% inst {100.0 * smp / max(ts, 1):5.1f}% smp lanes {th / ex if ex else 0:4.1f}  {f}:{ln}  {src[:110]}")


if __name__ == "__main__":
    rep, mode = sys.argv[1], sys.argv[2]
    if mode == "sass":
        sass(rep)
    else:
        lines(rep, int(sys.argv[3]) if len(sys.argv) > 3 else 40)
