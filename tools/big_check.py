#!/usr/bin/env python
"""Full-size property checks on one GPU (configs 3 and 5 of BASELINE.json):
counts sum to the number of windows, distinct keys strictly ascending, sum(key*count) equals the
sum over all extracted keys (mod 2^64).  usage: big_check.py [--n 3100000000] [--k 31] [--rc]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kman_b200 import fasta  # noqa: E402
from kman_b200.engine import get_engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=3_100_000_000)
ap.add_argument("--k", type=int, default=31)
ap.add_argument("--records", type=int, default=24)
args = ap.parse_args()
eng = get_engine(0)
n, k = args.n, args.k
t0 = time.time()
lut = np.frombuffer(b"ACGT", np.uint8)
parts, starts, pos = [], [0], 0
per = n // args.records
for r in range(args.records):
    m = per if r + 1 < args.records else n - per * (args.records - 1)
    rng = np.random.default_rng(1234 + r)
    x = np.empty(m + 1, np.uint8)
    step = 1 << 27
    for s in range(0, m, step):
        e = min(m, s + step)
        x[s:e] = lut[rng.integers(0, 4, size=e - s, dtype=np.uint8)]
    x[m] = 10
    parts.append(x)
    pos += m + 1
    starts.append(pos)
bases = np.concatenate(parts)
del parts
flat = fasta.FlatInput(bases, np.array(starts, np.uint64), ["chr%d" % (i + 1) for i in range(args.records)], ["chr%d" % (i + 1) for i in range(args.records)])
n_win = flat.n_windows(k)
print(f"generated {n} bases in {args.records} records, {n_win} windows, {time.time()-t0:.1f}s", flush=True)
d = eng.upload(flat, alphabet="ACGT", with_names=False)
del bases
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
a = eng.extract(d, k, False, val_bytes=0, want_hist=True)
e1.record(); torch.cuda.synchronize(); t_ex = e0.elapsed_time(e1)
assert a.n == n_win, (a.n, n_win)
kb = a.key_bytes
words = a.keys[: a.n * kb].view(torch.int64)
ksum = int(words.sum().item()) & (2**64 - 1) if kb == 8 else None
n_keys = a.n
e0.record()
tab = eng.sort_count(a)
e1.record(); torch.cuda.synchronize(); t_sort = e0.elapsed_time(e1)
t_rle = 0.0
print("sort passes", eng.lib.kmg_get_stat(b"sort_passes"), "hybrid path", eng.lib.kmg_get_stat(b"hybrid_path"),
      "irregular tiles", eng.lib.kmg_get_stat(b"hybrid_irregular"), flush=True)
counts = tab.counts[: tab.n * 4].view(torch.int32)
total = int(counts.sum(dtype=torch.int64).item())
print(f"extract {t_ex:.1f} ms, sort+count {t_sort:.1f} ms ({n_keys/t_sort/1e6:.2f} G keys/s); distinct {tab.n}", flush=True)
assert total == n_win, (total, n_win)
if kb == 8:
    keys = tab.keys[: tab.n * 8].view(torch.int64)
    ok = True
    step = 1 << 28
    for s in range(0, tab.n - 1, step):
        e = min(tab.n, s + step + 1)
        ok = ok and bool((keys[s + 1 : e] > keys[s : e - 1]).all())
    assert ok, "distinct keys not strictly ascending"
    acc = 0
    for s in range(0, tab.n, step):
        e = min(tab.n, s + step)
        acc += int((keys[s:e] * counts[s:e].to(torch.int64)).sum().item())
    assert acc & (2**64 - 1) == ksum, "checksum of checksums mismatch"
else:
    kk = tab.keys[: tab.n * 16].view(torch.int64).view(-1, 2)
    lo, hi = kk[:, 0], kk[:, 1]
    # (hi, lo) strictly ascending; lo compared as unsigned
    lo_u = lo ^ (-(2**63))
    asc = (hi[1:] > hi[:-1]) | ((hi[1:] == hi[:-1]) & (lo_u[1:] > lo_u[:-1]))
    assert bool(asc.all()), "distinct 128-bit keys not strictly ascending"
# steady state: the first calls above include allocations and module loading
del a, tab, counts
torch.cuda.synchronize()
ts = []
for _ in range(3):
    e0.record()
    tab = eng.sort_count(eng.extract(d, k, False, val_bytes=0, reuse="big_", want_hist=True), reuse="big_")
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"BIG_CHECK_OK n={n} k={k} windows={n_win} first pass {n_win/(t_ex+t_sort+t_rle)/1e6:.2f} G k-mers/s, "
      f"steady state {min(ts):.1f} ms per pass = {n_win/min(ts)/1e6:.2f} G k-mers/s")
