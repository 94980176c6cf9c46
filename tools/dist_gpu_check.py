#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun on the GPU box): distributed count/uniq of a small
duplicated genome against the CPU oracle; rank-order concatenation must equal the global table."""
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
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
eng = get_engine(lr)
dc = DistributedCounter(eng)
print("p2p path:", dc.p2p, flush=True) if rank == 0 else None
ok = True
for k, rc in ((31, False), (21, True), (45, False)):
    seq = ko.synth_bases(3_000_000, 77).decode()
    seq = seq[:2_000_000] + seq[:1_000_000]
    recs = [("chr1", seq)]
    flat = fasta.from_records(recs)
    # single flat record + separator: drop the separator so chunking sees only bases
    flat = fasta.FlatInput(flat.bases[:-1], flat.rec_starts, flat.names, flat.titles)
    d = dc.shard(flat, k, alphabet="ACGT")
    tab = dc.count(d, k, rc)
    ku, cu = tab.keys_host().copy(), tab.counts_host().copy()
    s = dc.uniq(d, k, rc)
    su, sv = s.keys_host().copy(), s.vals_host().copy()
    out = [None] * world
    dist.all_gather_object(out, (ku, cu, su, sv))
    if rank == 0:
        _, _, det = ko.count_np(recs, k, rc, "ACGT")
        gk = np.concatenate([o[0] for o in out])
        gc = np.concatenate([o[1] for o in out])
        want = det["narrow"]["keys"]
        wk = want[0] if len(want) == 1 else np.stack([want[1], want[0]], axis=1)
        good = gk.shape == wk.shape and (gk == wk).all() and (gc == det["narrow"]["counts"]).all()
        *_, du = ko.uniq_np(recs, k, rc, "ACGT")
        want = du["narrow"]["keys"]
        wk = want[0] if len(want) == 1 else np.stack([want[1], want[0]], axis=1)
        gs = np.concatenate([o[2] for o in out])
        gv = np.concatenate([o[3] for o in out])
        wv = (du["narrow"]["pos"].astype(np.uint64) << np.uint64(1)) | du["narrow"]["strand"].astype(np.uint64)
        good2 = gs.shape == wk.shape and (gs == wk).all() and (gv == wv).all()
        print(f"k={k} rc={rc} world={world}: count {'OK' if good else 'MISMATCH'} ({gk.shape[0]} distinct), uniq {'OK' if good2 else 'MISMATCH'} ({gs.shape[0]})", flush=True)
        ok = ok and good and good2
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("DIST_GPU_OK" if ok else "DIST_GPU_FAIL")
    sys.exit(0 if ok else 1)
