#!/usr/bin/env python
"""Per-SASS-instruction stall-reason view of one profiled launch: python tools/ncu_stalls.py REP [reason] [top]
Lists the instructions with the most samples of `reason` (default stall_no_inst) and the totals per reason."""
import csv, io, subprocess, sys
rep = sys.argv[1]; reason = sys.argv[2] if len(sys.argv) > 2 else "stall_no_inst"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
idx = {c: hdr.index(c) for c in cols}
ia, isrc, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
recs = []
for r in rows[2:]:
    if len(r) <= max(idx.values()) or not r[ia].startswith("0x"):
        continue
    recs.append((int(r[ia], 16), r[isrc].strip(), int(r[iex]), {c: int(r[idx[c]] or 0) for c in cols}))
base = recs[0][0]
tot = {c: sum(x[3][c] for x in recs) for c in cols}
alls = sum(tot.values())
print("# totals:", ", ".join(f"{c[6:]} {100.0 * v / alls:.1f}%" for c, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
for a, s, ex, st in sorted(recs, key=lambda x: -x[3][reason])[:top]:
    print(f"{a - base:05x} {reason[6:]} {100.0 * st[reason] / max(tot[reason], 1):5.1f}%  all {100.0 * sum(st.values()) / alls:5.2f}%  {s}")
