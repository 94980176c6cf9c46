#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source sass` output: opcode mix, stall reasons,
hottest instructions.  usage: sass_profile.py file.csv [top_n]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
body = []
for r in rows[hi + 1:]:
    if r and r[0] in ("Kernel Name", "Address"):
        break  # next launch section
    if len(r) == len(hdr):
        body.append(r)
isrc, ins, samp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
num = lambda s: int(float(s)) if s not in ("", "-") else 0  # noqa: E731
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "(Not Issued)" not in h]
tot_inst = sum(num(r[ins]) for r in body)
tot_s = sum(num(r[samp]) for r in body) or 1
print("total warp-inst", tot_inst, "samples", tot_s, "sass lines", len(body))
mix, smp = Counter(), Counter()
for r in body:
    parts = r[isrc].split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    op = op.split(".")[0]
    mix[op] += num(r[ins])
    smp[op] += num(r[samp])
for op, c in mix.most_common(top):
    print(f"{op:12s} inst {c / tot_inst * 100:5.1f}%  samples {smp[op] / tot_s * 100:5.1f}%")
st = Counter()
for r in body:
    for i in stall_cols:
        st[hdr[i]] += num(r[i])
print({k: round(v / tot_s * 100, 1) for k, v in st.most_common(10)})
print("hottest instructions by samples:")
order = sorted(range(len(body)), key=lambda i: -num(body[i][samp]))[:top]
for i in sorted(order):
    r = body[i]
    tops = sorted(((num(r[c]), hdr[c]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {num(r[samp]) / tot_s * 100:5.2f}% inst={num(r[ins]):9d} {r[isrc][:70]:70s} {tops}")
