#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun on the GPU box, 2..8 ranks): distributed count / uniq
against the CPU oracle; rank-order concatenation of the per-rank tables must equal the global table.

Inputs are MULTI-RECORD flat buffers WITH their separators (chunk boundaries fall anywhere, also on
and next to a separator; kmermaid/seq.py:361-383 chunking, batcher.py:387-388 records never share a
k-mer), with duplicated stretches (counts > 1), N runs + isolated IUPAC symbols (the wide stream
under the default alphabet) and soft-masked lower case; consecutive calls on ONE counter use inputs
whose lengths differ (also by one window), so that ranks see different chunk sizes from call to
call.  Used by tests/test_gpu_dist.py (pytest -m gpu, when the box has >= 2 GPUs) and by hand."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import kmer_oracle as ko  # noqa: E402
from kman_b200 import fasta  # noqa: E402
from kman_b200.dist import DistributedCounter  # noqa: E402
from kman_b200.engine import get_engine  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
# KMG_TEST_ONE_GPU=1: every rank uses cuda:0 and the plumbing runs over gloo (NCCL refuses two ranks on
# one device; the peer buffers are CUDA-IPC mappings either way) -- the topology the driver's
# single-GPU test box can run
one_gpu = os.environ.get("KMG_TEST_ONE_GPU") == "1"
if one_gpu:
    lr = 0
torch.cuda.set_device(lr)
if one_gpu:
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
eng = get_engine(lr)
dc = DistributedCounter(eng)
if rank == 0:
    print("p2p path:", dc.p2p, "shared cursors:", dc.shared, "backend:", dist.get_backend(), flush=True)


def genome(seed, n, n_rec, with_other):
    rng = np.random.default_rng(seed)
    b = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=n, dtype=np.uint8)].copy()
    b[n // 2 : n // 2 + n // 4] = b[: n // 4]  # duplicated stretch: counts > 1, non-singletons
    if with_other:
        for _ in range(6):
            s, ln = int(rng.integers(0, n - 3000)), int(rng.integers(1, 2500))
            b[s : s + ln] = ord("N")
        iso = rng.integers(0, n, size=40)
        b[iso] = np.frombuffer(b"RYKMSW", np.uint8)[rng.integers(0, 6, size=iso.size)]
        b[n // 3 : n // 3 + n // 9] |= 0x20  # soft-masked
    cuts = sorted(int(x) for x in rng.integers(1, n - 1, size=n_rec - 1))
    # one record boundary exactly where a rank's chunk starts (2 ranks: the middle)
    cuts[0] = n // 2
    cuts = sorted(set(cuts))
    recs, prev = [], 0
    for i, c in enumerate(cuts + [n]):
        recs.append(("chr%d" % (i + 1), b[prev:c].tobytes().decode()))
        prev = c
    return recs


def rows(limbs):
    return limbs[0] if len(limbs) == 1 else np.stack(limbs[::-1], axis=1)


def wide_rows(limbs):
    """the device keeps 4-bit keys in 128 bits up to k = 32 and in 256 bits up to k = 64"""
    want = 2 if len(limbs) <= 2 else 4
    return rows([np.zeros_like(limbs[0])] * (want - len(limbs)) + list(limbs))


ok = True
only = os.environ.get("KMG_TEST_CASES")  # e.g. "0,4": a subset (the pytest wrapper keeps its run short)
CASES = [
    # k, rc, n_bases, records, alphabet, non-ACGT symbols
    (31, False, 3_000_000, 5, "ACGT", False),
    (31, False, 2_999_999, 5, "ACGT", False),   # one base shorter on the same counter (dist.py cap key)
    (21, True, 2_500_001, 3, "ACGT", False),
    (45, False, 2_400_000, 4, "ACGT", False),
    (25, False, 2_600_000, 4, None, True),      # default IUPAC alphabet: N / R / Y windows are the wide stream
    (25, True, 1_300_000, 3, None, True),
    (31, False, 3_400_001, 2, "ACGT", True),    # ACGT-only alphabet: windows with N are skipped
    (45, True, 1_200_000, 3, None, True),       # k > 32 with IUPAC symbols: 256-bit wide keys
]
if only:
    CASES = [CASES[int(i)] for i in only.split(",")]
for k, rc, n, n_rec, alphabet, other in CASES:
    recs = genome(1000 + k + n % 7, n, n_rec, other)
    flat = fasta.from_records(recs)
    d = dc.shard(flat, k, alphabet=alphabet)
    tabs = dc.count_streams(d, k, rc)
    got_c = [(t.keys_host().copy(), t.counts_host().copy()) for t in tabs]
    sing = dc.uniq_streams(d, k, rc)
    got_u = [(s.keys_host().copy(), s.vals_host().copy()) for s in sing]
    out = [None] * world
    dist.all_gather_object(out, (got_c, got_u))
    if rank == 0:
        ab = alphabet or ko.DEFAULT_ALPHABET
        _, _, det = ko.count_np(recs, k, rc, ab)
        *_, du = ko.uniq_np(recs, k, rc, ab)
        msgs = []
        for si, name in enumerate(("narrow", "wide")):
            wk = rows(det[name]["keys"]) if name == "narrow" else wide_rows(det[name]["keys"])
            if name == "wide":
                if wk.shape[0] == 0:
                    assert all(len(o[0]) == 1 for o in out), "no wide windows, yet a rank returned a wide table"
                    continue
            gk = np.concatenate([o[0][si][0].reshape((-1,) + wk.shape[1:]) for o in out])
            gc = np.concatenate([o[0][si][1] for o in out])
            good = gk.shape == wk.shape and (gk == wk).all() and (gc == det[name]["counts"]).all()
            uk = rows(du[name]["keys"]) if name == "narrow" else wide_rows(du[name]["keys"])
            gs = np.concatenate([o[1][si][0].reshape((-1,) + uk.shape[1:]) for o in out])
            gv = np.concatenate([o[1][si][1] for o in out])
            wv = (du[name]["pos"].astype(np.uint64) << np.uint64(1)) | du[name]["strand"].astype(np.uint64)
            good2 = gs.shape == uk.shape and (gs == uk).all() and (gv.astype(np.uint64) == wv).all()
            msgs.append(f"{name}: count {'OK' if good else 'MISMATCH'} ({gk.shape[0]} distinct), "
                        f"uniq {'OK' if good2 else 'MISMATCH'} ({gs.shape[0]})")
            ok = ok and good and good2
        print(f"k={k} rc={rc} n={n} records={n_rec} alphabet={ab} world={world}: " + "; ".join(msgs), flush=True)
dc.close()
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("DIST_GPU_OK" if ok else "DIST_GPU_FAIL")
    sys.exit(0 if ok else 1)
