#!/usr/bin/env python
"""Per-stage wall-clock of the distributed count step (synchronising between stages; run under
torchrun with KMG_DIST_TIMING=1).  A development tool: the sum is larger than the pipelined step."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["KMG_DIST_TIMING"] = "1"
from kman_b200 import fasta  # noqa: E402
from kman_b200.dist import DistributedCounter  # noqa: E402
from kman_b200.engine import get_engine  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
eng = get_engine(lr)
dc = DistributedCounter(eng)
n, k = 100_000_000, 31
rng = np.random.default_rng(1234 + rank)
chunk = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=n, dtype=np.uint8)]
flat = fasta.FlatInput(chunk, np.array([0, n + 1], np.uint64), ["chr1"], ["chr1"])
d = eng.upload(flat, alphabet="ACGT", with_names=False)
d.pos_offset = rank * n
for it in range(8):
    if it == 3:
        dc._timing.clear()
        dc._t_last = None
    dc.count(d, k, False)
if rank == 0:
    tot = sum(v for kk, v in dc._timing.items() if kk != "(outside)")
    for kk, v in dc._timing.items():
        print(f"{kk:36s} {v / 5 * 1e3:8.3f} ms")
    print(f"{'sum (serialised)':36s} {tot / 5 * 1e3:8.3f} ms")
dist.barrier()
dist.destroy_process_group()
