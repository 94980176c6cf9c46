#!/usr/bin/env python
"""Split `ncu --page source --csv --print-source sass` output into the code between barriers (BAR.SYNC)
and print per segment: warp instructions executed, thread-level lane efficiency, samples, shared-memory
wavefronts, main stall reasons.  usage: sass_phases.py file.csv"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
body = []
for r in rows[hi + 1:]:
    if r and r[0] in ("Kernel Name", "Address"):
        break
    if len(r) == len(hdr):
        body.append(r)
col = {h: i for i, h in enumerate(hdr)}
num = lambda s: float(s) if s not in ("", "-") else 0.0  # noqa: E731
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "(Not Issued)" not in h]
tot_i = sum(num(r[col["Instructions Executed"]]) for r in body)
tot_s = sum(num(r[col["# Samples"]]) for r in body) or 1
seg, segs = [], []
for i, r in enumerate(body):
    seg.append(i)
    if "BAR.SYNC" in r[col["Source"]] or "EXIT" in r[col["Source"]].split()[-2:][0]:
        segs.append(seg)
        seg = []
if seg:
    segs.append(seg)
print(f"total warp-inst {tot_i:.0f} samples {tot_s:.0f}")
print(" seg  lines         inst%  lane-eff  samples%  smem-wavefronts(excess)  top stalls ... first/last instruction")
for k, sg in enumerate(segs):
    ins = sum(num(body[i][col["Instructions Executed"]]) for i in sg)
    thr = sum(num(body[i][col["Thread Instructions Executed"]]) for i in sg)
    smp = sum(num(body[i][col["# Samples"]]) for i in sg)
    wf = sum(num(body[i][col["L1 Wavefronts Shared"]]) for i in sg)
    wfx = sum(num(body[i][col["L1 Wavefronts Shared Excessive"]]) for i in sg)
    if ins / tot_i < 0.002 and smp / tot_s < 0.002:
        continue
    st = Counter()
    for i in sg:
        for c in stall_cols:
            st[hdr[c][6:]] += num(body[i][c])
    tops = ", ".join(f"{n} {v / max(smp, 1) * 100:.0f}%" for n, v in st.most_common(3))
    print(f"{k:4d} {sg[0]:5d}-{sg[-1]:5d} {ins / tot_i * 100:6.1f}% {thr / max(ins, 1) / 32 * 100:7.0f}% {smp / tot_s * 100:8.1f}%"
          f" {wf / 1e6:9.1f}M ({wfx / 1e6:.1f}M)  {tops}   | {body[sg[0]][col['Source']].strip()[:40]}")
