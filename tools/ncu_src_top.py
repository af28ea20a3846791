#!/usr/bin/env python
"""Top stall lines of one kernel from an ncu --page source --csv export (development tool)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
blk = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = rows[starts[blk]:(starts[blk + 1] if blk + 1 < len(starts) else len(rows))]
print(rows[0][1][:120])
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = rows[1]
i_src, i_s, i_ex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
i_wf, i_wfi = h.index("L1 Wavefronts Shared"), h.index("L1 Wavefronts Shared Ideal")
data = [r for r in rows[2:] if len(r) == len(h)]
tot = sum(int(r[i_s]) for r in data)
print("total samples", tot)
agg = {}
for r in data:
    for i in stall_cols:
        agg[h[i]] = agg.get(h[i], 0) + int(r[i])
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
wf = sum(int(r[i_wf]) for r in data); wfi = sum(int(r[i_wfi]) for r in data)
print("smem wavefronts", wf, "ideal", wfi)
for n, r in sorted(enumerate(data), key=lambda nr: -int(nr[1][i_s]))[:top]:
    st = sorted(((int(r[i]), h[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{n:5d} {int(r[i_s]):6d} {100*int(r[i_s])/tot:5.1f}% ex={r[i_ex]:>8s} wf={r[i_wf]:>9s}/{r[i_wfi]:>9s} {r[i_src].strip()[:70]:70s} {st}")
