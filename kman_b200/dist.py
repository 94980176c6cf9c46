"""Multi-GPU path: one process per GPU, torch.distributed for the plumbing.

There is nothing like this in the reference (it merges batch files on one host,
kmermaid/join.py:63-93).  The two partitioning rules are:

  * INPUT: the flat base buffer is cut into `world` contiguous chunks of window starts; a
    rank's chunk of bases carries k-1 extra bases so that every window lies in exactly one
    chunk -- the reference's own chunking rule, Sequence.batcher (kmermaid/seq.py:361-383,
    known answer tests/test_seq.py:136-138).
  * KEYS: range partition by the top key bits, part = ((key >> (bits-16)) * world) >> 16
    (libkmg's kmg_range_partition), so rank r ends up with key range r: equal k-mers meet on
    one GPU (counts are final, no second reduction) and the rank-order concatenation of the
    per-rank sorted outputs is globally sorted.

The only collective on the data path is ONE all-to-all of the partitioned keys (and payload
in uniq mode), preceded by an all-gather of the world x world count matrix.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


# ---- pure host logic (unit-tested on CPU) ------------------------------------------------------
def chunk_windows(n_bases: int, k: int, world: int) -> List[Tuple[int, int]]:
    """Split the window starts [0, n_bases-k+1) into `world` contiguous ranges [b, e).

    Rank r needs bases [b, e + k - 1): consecutive chunks overlap by k-1 bases, exactly like
    Sequence.batcher's `start += batchSize - k + 1` (seq.py:379-383)."""
    n_win = max(0, n_bases - k + 1)
    per = (n_win + world - 1) // world if world else 0
    out = []
    for r in range(world):
        b = min(r * per, n_win)
        e = min(b + per, n_win)
        out.append((b, e))
    return out


def chunk_bases(n_bases: int, k: int, world: int) -> List[Tuple[int, int]]:
    """Base ranges [b, e) each rank must hold (window range + k-1 overlap)."""
    return [(b, min(n_bases, e + k - 1) if e > b else b) for b, e in chunk_windows(n_bases, k, world)]


def part_of_keys(keys_hi16: np.ndarray, world: int) -> np.ndarray:
    """numpy mirror of libkmg's RangeDigit on the top 16 key bits (used by the CPU tests)."""
    return (keys_hi16.astype(np.uint64) * np.uint64(world)) >> np.uint64(16)


def sort_bits_after_partition(key_bits: int, world: int) -> int:
    """Bits the local sort still has to cover.  When `world` is a power of two, the top
    log2(world) bits are constant inside a rank after the range partition."""
    if world & (world - 1) == 0:
        return max(0, key_bits - (world.bit_length() - 1))
    return key_bits


# ---- exchange ------------------------------------------------------------------------------------
def gather_count_matrix(send_counts: np.ndarray, device: torch.device, group=None) -> np.ndarray:
    """All-gather of every rank's per-destination counts -> (world, world) int64 matrix
    M[src, dst]."""
    world = dist.get_world_size(group)
    mine = torch.from_numpy(np.ascontiguousarray(send_counts, dtype=np.int64)).to(device)
    allc = torch.empty(world * world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allc, mine, group=group)
    return allc.cpu().numpy().reshape(world, world)


def exchange(send: torch.Tensor, send_counts: np.ndarray, elems_per_item: int = 1, group=None,
             matrix: Optional[np.ndarray] = None) -> Tuple[torch.Tensor, np.ndarray]:
    """One all-to-all(v).  `send` holds the items grouped by destination rank (int64 view,
    `elems_per_item` int64 per item); returns (recv tensor, per-source item counts)."""
    rank = dist.get_rank(group)
    if matrix is None:
        matrix = gather_count_matrix(send_counts, send.device, group)
    recv_counts = matrix[:, rank].copy()
    recv = torch.empty(int(recv_counts.sum()) * elems_per_item, dtype=send.dtype, device=send.device)
    dist.all_to_all_single(
        recv,
        send[: int(np.sum(send_counts)) * elems_per_item],
        output_split_sizes=[int(c) * elems_per_item for c in recv_counts],
        input_split_sizes=[int(c) * elems_per_item for c in send_counts],
        group=group,
    )
    return recv, recv_counts


@dataclass
class RankShard:
    """What one rank holds of a distributed input."""

    device_input: object  # engine.DeviceInput of the rank's chunk (+ k-1 overlap)
    win_begin: int  # local window range inside the chunk (always 0 .. n)
    win_end: int
    n_windows_global: int


class DistributedCounter:
    """`kmer count` / `kmer uniq` across the GPUs of one box."""

    def __init__(self, engine, group=None):
        self.eng = engine
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def shard(self, flat, k: int, alphabet: Optional[str] = None, natype=None):
        """Upload this rank's chunk of a host-resident flat input (k-1 overlap)."""
        from kman_b200 import alphabet as ab
        from kman_b200.fasta import FlatInput

        natype = natype or ab.NATYPES.DNA
        n_bases = int(flat.bases.shape[0])
        b, e = chunk_bases(n_bases, k, self.world)[self.rank]
        sub = FlatInput(flat.bases[b:e], flat.rec_starts, flat.names, flat.titles)
        d = self.eng.upload(sub, alphabet, natype, with_names=False)
        d.pos_offset = b
        return d

    def _partition_exchange(self, a, with_vals: bool):
        """range partition -> count matrix -> all-to-all; returns the received KeyArray."""
        from kman_b200.engine import KeyArray

        a, pc = self.eng.range_partition(a, self.world)
        per = a.key_bytes // 8
        send = a.keys[: max(a.n, 1) * a.key_bytes].view(torch.int64)
        matrix = gather_count_matrix(pc, a.keys.device, self.group)
        recv, recv_counts = exchange(send, pc, per, self.group, matrix)
        n = int(recv_counts.sum())
        rvals = None
        if with_vals:
            vper = 1
            v64 = a.vals[: max(a.n, 1) * a.val_bytes]
            if a.val_bytes == 4:
                # widen the payload for the exchange so one int64 all-to-all suffices
                v64 = v64.view(torch.int32).to(torch.int64).contiguous().view(torch.uint8)
            rv, _ = exchange(v64.view(torch.int64), pc, vper, self.group, matrix)
            rvals = rv.view(torch.uint8)
        kbuf = recv.view(torch.uint8)
        alt = torch.empty(max(kbuf.numel(), 16), dtype=torch.uint8, device=kbuf.device)
        valt = torch.empty(max(rvals.numel(), 16), dtype=torch.uint8, device=kbuf.device) if with_vals else None
        if kbuf.numel() == 0:
            kbuf = torch.empty(16, dtype=torch.uint8, device=alt.device)
        return KeyArray(kbuf, alt, rvals, valt, n, a.key_bytes, 8 if with_vals else 0, a.k, a.wide)

    def count(self, d, k: int, rc: bool = False):
        """This rank's slice (key range `rank`) of the global count table, narrow stream."""
        a = self.eng.extract(d, k, rc, wide=False, val_bytes=0)
        n_other = torch.tensor([a.n_other], dtype=torch.int64, device=a.keys.device)
        dist.all_reduce(n_other, group=self.group)
        if int(n_other.item()):
            raise ValueError("distributed path handles the narrow (plain ACGT) stream only in this build; "
                             f"input holds {int(n_other.item())} windows with other alphabet symbols")
        r = self._partition_exchange(a, with_vals=False)
        r = self.eng.sort(r, 0, sort_bits_after_partition(r.key_bits, self.world))
        return self.eng.rle_count(r)

    def uniq(self, d, k: int, rc: bool = False):
        """This rank's slice of the global singleton list (keys + (pos<<1|strand) payload)."""
        a = self.eng.extract(d, k, rc, wide=False, val_bytes=8)
        n_other = torch.tensor([a.n_other], dtype=torch.int64, device=a.keys.device)
        dist.all_reduce(n_other, group=self.group)
        if int(n_other.item()):
            raise ValueError("distributed path handles the narrow (plain ACGT) stream only in this build")
        r = self._partition_exchange(a, with_vals=True)
        r = self.eng.sort(r, 0, sort_bits_after_partition(r.key_bits, self.world))
        return self.eng.singletons(r)
