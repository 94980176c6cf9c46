#!/usr/bin/env python
"""Randomised cross-check of the hybrid sort family (kmg_radix_sort key-only, kmg_sort_count,
kmg_sort_uniq; 8- and 16-byte keys) against numpy on adversarial key distributions: clusters of
distinct keys sharing long prefixes (every size class: thread / warp / block / oversize), exact
duplicates, skewed top bytes, constant high bits, forced prefix widths and tile widths.
usage: python tools/fuzz_sort.py [--seconds 120] [--seed 1] [--trials N]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kman_b200.engine import KeyArray, get_engine  # noqa: E402


def make_keys(rng, n, bits):
    """u64 keys in [0, 2^bits) with clusters; returns (hi, lo) limbs for bits > 64."""
    wide = bits > 64
    hb = bits - 64 if wide else bits
    top = rng.integers(0, 2**63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    if hb < 64:
        top &= np.uint64((1 << hb) - 1)
    low = rng.integers(0, 2**63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    pos = 0
    n_clusters = int(rng.integers(0, 60))
    for _ in range(n_clusters):
        size = int(rng.choice([20, 30, 100, 128, 129, 700, 3000, 4097, 6000, 9000]))
        if pos + size > n // 2:
            break
        share = int(rng.integers(min(28, hb), hb + 1))  # bits of the top limb the cluster shares
        base = top[pos] >> np.uint64(hb - share) << np.uint64(hb - share) if share < hb else top[pos]
        mask = np.uint64((1 << (hb - share)) - 1) if share < hb else np.uint64(0)
        top[pos : pos + size] = base | (top[pos : pos + size] & mask)
        if rng.random() < 0.3:  # some exact duplicates inside
            k = max(1, size // 10)
            top[pos : pos + k] = top[pos]
            low[pos : pos + k] = low[pos]
        pos += size
    if rng.random() < 0.3:  # global duplicates
        idx = rng.integers(0, n, size=n // 4)
        top[idx] = top[(idx * 7 + 1) % n]
        low[idx] = low[(idx * 7 + 1) % n]
    if rng.random() < 0.2:  # skewed top byte
        sel = rng.random(n) < 0.4
        top[sel] >>= np.uint64(3)
    perm = rng.permutation(n)
    return (top[perm], low[perm]) if wide else (top[perm], None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--trials", type=int, default=0, help="stop after this many trials (0: run for --seconds)")
    args = ap.parse_args()
    eng = get_engine(0)
    lib = eng.lib
    rng = np.random.default_rng(args.seed)
    dev = eng.device
    t_end = time.time() + args.seconds
    trials, paths = 0, {}
    while time.time() < t_end and (args.trials == 0 or trials < args.trials):
        n = int(rng.integers(1 << 20, 3_500_000))
        wide = rng.random() < 0.3
        bits = int(rng.choice([66, 90, 126, 128])) if wide else int(rng.choice([32, 40, 50, 62, 64]))
        mode = str(rng.choice(["sort", "count", "uniq"]))
        hi, lo = make_keys(rng, n, bits)
        lib.kmg_set_option(b"hybrid", 1)
        lib.kmg_set_option(b"hybrid_pb", int(rng.choice([0, 0, 16, 24])))
        lib.kmg_set_option(b"local_tile", int(rng.choice([7936, 7936, 4096, 2500])))
        lib.kmg_set_option(b"count_fused", int(rng.random() < 0.8))
        raw = np.stack([lo, hi], axis=1) if wide else hi
        kb = 16 if wide else 8
        t = lambda x: torch.from_numpy(np.ascontiguousarray(x).view(np.uint8).reshape(-1)).to(dev)  # noqa: E731
        z = lambda b: torch.zeros(max(b, 16), dtype=torch.uint8, device=dev)  # noqa: E731
        if wide:
            order = np.lexsort((lo, hi))
            srt = raw[order]
            head = np.ones(n, bool)
            head[1:] = (srt[1:] != srt[:-1]).any(axis=1)
        else:
            order = np.argsort(raw, kind="stable")
            srt = raw[order]
            head = np.ones(n, bool)
            head[1:] = srt[1:] != srt[:-1]
        idx = np.flatnonzero(head)
        desc = f"trial {trials}: n={n} bits={bits} mode={mode} path={lib.kmg_get_stat(b'hybrid_path')}"
        if mode == "sort":
            a = KeyArray(t(raw), z(n * kb), None, None, n, kb, 0, bits // 2, False)
            a = eng.sort(a, 0, bits)
            got = a.keys_host()
            assert np.array_equal(got, srt), desc
        elif mode == "count":
            a = KeyArray(t(raw), z(n * kb), None, None, n, kb, 0, bits // 2, False)
            tab = eng.sort_count(a, bits)
            keys = tab.keys[: tab.n * kb].cpu().numpy().view(np.uint64)
            keys = keys.reshape(-1, 2) if wide else keys
            counts = tab.counts[: tab.n * 4].cpu().numpy().view(np.uint32)
            assert np.array_equal(keys, srt[idx]), desc
            assert np.array_equal(counts.astype(np.int64), np.diff(np.append(idx, n))), desc
        else:
            vdt = np.uint32 if rng.random() < 0.5 else np.uint64
            vals = np.arange(n, dtype=vdt)
            a = KeyArray(t(raw), z(n * kb), t(vals), z(n * vals.itemsize), n, kb, vals.itemsize, bits // 2, False)
            r = eng.sort_uniq(a, bits)
            keys = r.keys[: r.n * kb].cpu().numpy().view(np.uint64)
            keys = keys.reshape(-1, 2) if wide else keys
            got_v = r.vals[: r.n * vals.itemsize].cpu().numpy().view(vdt)
            runlen = np.diff(np.append(idx, n))
            one = idx[runlen == 1]
            assert np.array_equal(keys, srt[one]), desc
            assert np.array_equal(got_v.astype(np.int64), vals[order][one].astype(np.int64)), desc
        eng._status(eng._last_sort_ws)
        pth = int(lib.kmg_get_stat(b"hybrid_path"))
        paths[(mode, kb, pth)] = paths.get((mode, kb, pth), 0) + 1
        trials += 1
    lib.kmg_set_option(b"hybrid_pb", 0)
    lib.kmg_set_option(b"local_tile", 7936)
    lib.kmg_set_option(b"count_fused", 1)
    print("FUZZ_OK", trials, "trials; (mode, key bytes, hybrid path) counts:", dict(sorted(paths.items())))


if __name__ == "__main__":
    main()
