#!/usr/bin/env python
"""count / uniq TEXT of a genome-like input (tools/genome_like_bench.py's generator, a few Mbp)
against the CPU oracle, byte for byte.  usage: python tools/genome_like_parity.py [--n 4000000]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import kmer_oracle as ko  # noqa: E402
from genome_like_bench import genome_like  # noqa: E402
from kman_b200 import fasta  # noqa: E402
from kman_b200.engine import get_engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4_000_000)
args = ap.parse_args()
eng = get_engine(0)
recs = [("chrA like a genome", genome_like(args.n // 2, 3).tobytes().decode("latin-1")),
        ("chrB", genome_like(args.n // 2, 4).tobytes().decode("latin-1"))]
d = eng.upload(fasta.from_records(recs), alphabet="ACGT")
for k, rc in ((31, False), (21, True), (45, False)):
    got = eng.count_text(d, k, rc)
    path_c = eng.lib.kmg_get_stat(b"hybrid_path")
    assert got == ko.count_text_np(recs, k, rc, "ACGT"), ("count", k, rc)
    got = eng.uniq_text(d, k, rc)
    path_u = eng.lib.kmg_get_stat(b"hybrid_path")
    assert got == ko.uniq_text_np(recs, k, rc, "ACGT"), ("uniq", k, rc)
    print(f"k={k} rc={rc}: count and uniq text equal the oracle's ({len(got)} bytes of uniq text; hybrid paths {path_c}/{path_u})", flush=True)
print("GENOME_LIKE_PARITY_OK")
