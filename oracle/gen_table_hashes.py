#!/usr/bin/env python
"""Write tests/golden/table_hashes.json: sha256 of the canonical binary result tables of BASELINE.json's
configurations 2, 3 (scaled), 4 and 5 at sizes the reference itself cannot finish, computed with the
oracle's numpy tier (oracle/kmer_oracle.py `extract_np` + stable sort + run-length grouping, i.e.
kmermaid/seq.py:284-328, batch.py:156-168, join.py:95-130,243-285 restated; that tier is pinned
against the unmodified reference on the small goldens by tests/test_oracle_golden.py).

TEST INFRASTRUCTURE.  Runs in the build container (about 10 minutes, < 30 GB of host memory); the
`-m gpu` tests rebuild the same inputs with oracle/synth_configs.py and hash what the CUDA path returns.

usage: python oracle/gen_table_hashes.py [--only NAME[,NAME...]]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import kmer_oracle as ko  # noqa: E402
import synth_configs as sc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "table_hashes.json")

# name -> (records factory, k, rc, alphabet, mode)
CASES = {
    "cfg2_100mbp_k31_count": (lambda: sc.cfg2(100_000_000), 31, False, "ACGT", "count"),
    "cfg2dup_100mbp_k31_count": (lambda: sc.cfg2(100_000_000, dup=True), 31, False, "ACGT", "count"),
    "cfg2_100mbp_k31_uniq": (lambda: sc.cfg2(100_000_000), 31, False, "ACGT", "uniq"),
    "cfg3_24mbp_24rec_k31_count": (lambda: sc.cfg3(24_000_000, 24), 31, False, "ACGT", "count"),
    "cfg3_24mbp_24rec_k31_uniq": (lambda: sc.cfg3(24_000_000, 24), 31, False, "ACGT", "uniq"),
    "cfg4_10mbp_k25_uniq_iupac": (lambda: sc.cfg4(10_000_000), 25, False, "IUPAC", "uniq"),
    "cfg4_10mbp_k25_uniq_acgt": (lambda: sc.cfg4(10_000_000), 25, False, "ACGT", "uniq"),
    "cfg4_10mbp_k25_uniq_rc_iupac": (lambda: sc.cfg4(10_000_000), 25, True, "IUPAC", "uniq"),
    "cfg4_10mbp_k25_count_iupac": (lambda: sc.cfg4(10_000_000), 25, False, "IUPAC", "count"),
    "cfg4_100mbp_k25_uniq_iupac": (lambda: sc.cfg4(100_000_000), 25, False, "IUPAC", "uniq"),
    "cfg4_10mbp_k45_uniq_iupac": (lambda: sc.cfg4(10_000_000), 45, False, "IUPAC", "uniq"),
    "cfg5_20mbp_k63_count": (lambda: sc.cfg5(20_000_000), 63, False, "ACGT", "count"),
    "cfg5dup_20mbp_k63_count": (lambda: sc.cfg5(20_000_000, dup=True), 63, False, "ACGT", "count"),
    "cfg5_20mbp_k63_uniq_rc": (lambda: sc.cfg5(20_000_000), 63, True, "ACGT", "uniq"),
}


def digests(recs, k, rc, alphabet, mode):
    """Per stream: sort (stable, batch.py:156-168), group (join.py:95-130), keep all groups with their
    sizes (count, join.py:265-285) or the groups of size one with their coordinates (uniq, :243-263)."""
    ex = ko.extract_np(recs, k, rc, alphabet)
    out = {"n_windows": int(ex["n_windows"])}
    for name in ("narrow", "wide"):
        st = ex[name]
        order = ko._lexsort_limbs(st["keys"])
        srt = [l[order] for l in st["keys"]]
        heads, lens = ko._rle(srt)
        rows = sc.key_rows if name == "narrow" else sc.widen_rows
        if mode == "count":
            out[name] = sc.count_digest(rows([l[heads] for l in srt]), lens)
        else:
            sel = heads[lens == 1]
            vals = (st["pos"][order][sel].astype(np.uint64) << np.uint64(1)) | st["strand"][order][sel].astype(np.uint64)
            out[name] = sc.uniq_digest(rows([l[sel] for l in srt]), vals)
        out[name]["keys_in"] = int(st["pos"].shape[0])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    only = [x for x in args.only.split(",") if x]
    res = {}
    if os.path.exists(OUT):
        res = json.load(open(OUT))
    for name, (make, k, rc, alphabet, mode) in CASES.items():
        if only and name not in only:
            continue
        t0 = time.time()
        d = digests(make(), k, rc, alphabet, mode)
        d.update({"k": k, "rc": rc, "alphabet": alphabet, "mode": mode})
        res[name] = d
        print(f"{name}: {time.time() - t0:.1f} s  narrow rows {d['narrow']['rows']} wide rows {d['wide']['rows']}", flush=True)
        with open(OUT, "w") as fh:
            json.dump(res, fh, indent=1, sort_keys=True)
            fh.write("\n")


if __name__ == "__main__":
    main()
