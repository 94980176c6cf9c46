#!/usr/bin/env python
"""Count / uniq timings on a genome-LIKE synthetic input (not uniform random): 41 % GC, 10-15 % of
the sequence made of ~300 bp interspersed repeats (10 % diverged copies with poly-A tails),
microsatellites, soft-masked lower case and N gaps.  Shows what the hybrid sort's skew handling
(24-bit prefix, re-sorted irregular tiles, adaptive oversize pre-check) costs on such data.
usage: python tools/genome_like_bench.py [--n 100000000] [--k 31]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kman_b200 import fasta  # noqa: E402
from kman_b200.engine import get_engine  # noqa: E402


def genome_like(n, seed=7):
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    seq = acgt[rng.choice(4, size=n, p=[0.295, 0.205, 0.205, 0.295])].copy()
    # interspersed repeats: 4 families of ~300 bp, copies diverged by 10 %, each followed by a poly-A tail
    fams = [acgt[rng.integers(0, 4, size=300)] for _ in range(4)]
    n_copies = n // 2500
    starts = rng.integers(0, n - 400, size=n_copies)
    for s0 in starts:
        f = fams[rng.integers(0, 4)].copy()
        mut = rng.random(300) < 0.10
        f[mut] = acgt[rng.integers(0, 4, size=int(mut.sum()))]
        seq[s0 : s0 + 300] = f | 0x20  # soft-masked
        tail = int(rng.integers(10, 40))
        seq[s0 + 300 : s0 + 300 + tail] = ord("a")
    # microsatellites
    for s0 in rng.integers(0, n - 200, size=n // 50000):
        unit = acgt[rng.integers(0, 4, size=int(rng.integers(1, 5)))]
        reps = int(rng.integers(10, 40))
        run = np.tile(unit, reps)
        seq[s0 : s0 + run.size] = run
    # N gaps
    for s0 in rng.integers(0, n - 20000, size=max(1, n // 10_000_000)):
        seq[s0 : s0 + int(rng.integers(1000, 20000))] = ord("N")
    return seq


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000_000)
    ap.add_argument("--k", type=int, default=31)
    args = ap.parse_args()
    eng = get_engine(0)
    n, k = args.n, args.k
    per = n // 4
    recs_b, starts, pos = [], [0], 0
    for r in range(4):
        x = genome_like(per, seed=7 + r)
        recs_b += [x, np.array([10], np.uint8)]
        pos += per + 1
        starts.append(pos)
    flat = fasta.FlatInput(np.concatenate(recs_b), np.array(starts, np.uint64), ["chr%d" % i for i in range(4)], ["chr%d" % i for i in range(4)])
    d = eng.upload(flat, alphabet="ACGT", with_names=False)

    def count():
        return eng.sort_count(eng.extract(d, k, False, val_bytes=0, reuse="g_", want_hist=True), reuse="g_")

    def uniq():
        return eng.sort_uniq(eng.extract(d, k, False, val_bytes=4, reuse="gu_", want_hist=True))

    res = {}
    for hy in (0, 1, 2, 3):
        eng.lib.kmg_set_option(b"hybrid", min(hy, 1))  # (also resets the back-off)
        # 2 / 3: prefix width forced to 24 / 16 bits (a forced width also disables the back-off)
        eng.lib.kmg_set_option(b"hybrid_pb", {2: 24, 3: 16}.get(hy, 0))
        eng.lib.kmg_set_option(b"time_passes", 1)
        tab = count()
        torch.cuda.synchronize()
        print(f"   [one count call: onesweep launches {eng.lib.kmg_get_stat(b'sort_pass_count')} = {eng.lib.kmg_get_stat(b'sort_pass_ns')/1e6:.3f} ms, "
              f"local sort launches {eng.lib.kmg_get_stat(b'local_sort_count')} = {eng.lib.kmg_get_stat(b'local_sort_ns')/1e6:.3f} ms]")
        eng.lib.kmg_set_option(b"time_passes", 0)
        nk = int(tab.counts[: tab.n * 4].view(torch.int32).sum(dtype=torch.int64))
        t_c = timed(count)
        st_c = (eng.lib.kmg_get_stat(b"sort_passes"), eng.lib.kmg_get_stat(b"hybrid_path"), eng.lib.kmg_get_stat(b"hybrid_irregular"))
        u = uniq()
        t_u = timed(uniq)
        st_u = (eng.lib.kmg_get_stat(b"sort_passes"), eng.lib.kmg_get_stat(b"hybrid_path"), eng.lib.kmg_get_stat(b"hybrid_irregular"))
        res[hy] = (tab.n, u.n)
        print(f"hybrid {hy}: {nk} k-mers, {tab.n} distinct, {u.n} occur once")
        print(f"   count {t_c:7.3f} ms = {nk/t_c/1e6:6.2f} G k-mers/s  (passes {st_c[0]}, path {st_c[1]}, irregular tiles {st_c[2]})")
        print(f"   uniq  {t_u:7.3f} ms = {nk/t_u/1e6:6.2f} G k-mers/s  (passes {st_u[0]}, path {st_u[1]}, irregular tiles {st_u[2]})")
        print(f"   runs sorted by the block in the last hybrid sort {eng.lib.kmg_get_stat(b'hybrid_big_runs')}")
    eng.lib.kmg_set_option(b"hybrid_pb", 0)
    assert res[0] == res[1] == res[2] == res[3], "hybrid and plain sorts disagree"
    print("GENOME_LIKE_OK")


if __name__ == "__main__":
    main()
