"""Deterministic synthetic inputs of BASELINE.json's configurations and the canonical
binary form / hashes of their results.

TEST INFRASTRUCTURE (like everything under oracle/): imported by oracle/gen_table_hashes.py, which
runs the oracle's numpy tier in the build container and commits the hashes under tests/golden/, and
by the `-m gpu` tests, which rebuild the same inputs on the GPU box and hash what the CUDA path
returns.  Nothing under kman_b200/ imports it.
"""
from __future__ import annotations

import hashlib
from typing import Dict, List, Tuple

import numpy as np

_ACGT = np.frombuffer(b"ACGT", np.uint8)


def random_acgt(n: int, seed: int) -> np.ndarray:
    """SURVEY §8d generator: default_rng(seed).integers(0, 4, n, uint8) -> b"ACGT"[...]."""
    rng = np.random.default_rng(seed)
    out = np.empty(n, np.uint8)
    step = 1 << 26  # (same chunking as bench.py: the stream of one seed does not depend on it)
    for s in range(0, n, step):
        m = min(step, n - s)
        out[s : s + m] = _ACGT[rng.integers(0, 4, size=m, dtype=np.uint8)]
    return out


def split_records(bases: np.ndarray, cuts: List[int], prefix: str = "chr") -> List[Tuple[str, str]]:
    recs, prev = [], 0
    for i, c in enumerate(list(cuts) + [bases.size]):
        recs.append((f"{prefix}{i + 1} synthetic", bases[prev:c].tobytes().decode("latin-1")))
        prev = c
    return recs


def cfg2(n: int = 100_000_000, dup: bool = False) -> List[Tuple[str, str]]:
    """config 2: n bp, ONE record, seed 1234 (k=31, count).  dup: second half = copy of the first."""
    b = random_acgt(n, 1234)
    if dup:
        b = b.copy()
        b[n // 2 : 2 * (n // 2)] = b[: n // 2]
    return [("chr1 synthetic seed=1234", b.tobytes().decode("latin-1"))]


def cfg3(n: int, n_rec: int = 24) -> List[Tuple[str, str]]:
    """config 3 (scaled to n bp): n_rec records with lengths proportional to the human chromosomes,
    record i from seed 1234+i (k=31, count)."""
    human = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51,
             156, 57]
    w = np.array((human * ((n_rec + 23) // 24))[:n_rec], np.float64)
    lens = np.maximum(1, np.floor(w / w.sum() * n)).astype(np.int64)
    lens[0] += n - int(lens.sum())
    return [(f"chr{i + 1} synthetic seed={1234 + i}", random_acgt(int(L), 1234 + i).tobytes().decode("latin-1"))
            for i, L in enumerate(lens)]


def cfg4(n: int = 10_000_000, seed: int = 4321) -> List[Tuple[str, str]]:
    """config 4: 3 records; 50 N runs (lengths log-uniform 1..1e5), 200 soft-masked intervals (1e2..1e5),
    10 isolated IUPAC symbols, N runs touching a record start and a record end (k=25, uniq)."""
    b = random_acgt(n, seed).copy()
    rng = np.random.default_rng(seed + 1)
    scale = min(1.0, n / 10_000_000)
    for _ in range(50):
        ln = int(np.exp(rng.uniform(0, np.log(1e5 * scale + 1))))
        s = int(rng.integers(0, max(1, n - ln)))
        b[s : s + ln] = ord("N")
    for _ in range(200):
        ln = int(np.exp(rng.uniform(np.log(1e2), np.log(1e5 * scale + 1e2))))
        s = int(rng.integers(0, max(1, n - ln)))
        b[s : s + ln] |= 0x20  # lower case (n stays n)
    iso = rng.integers(0, n, size=10)
    b[iso] = np.frombuffer(b"RYKMSW", np.uint8)[rng.integers(0, 6, size=10)]
    cuts = [n // 3, (2 * n) // 3 + 7]
    b[:37] = ord("N")  # run touching the first record's start
    b[cuts[0] - 53 : cuts[0]] = ord("n")  # ... the first record's end
    b[cuts[0] : cuts[0] + 11] = ord("N")  # ... and the second record's start
    return split_records(b, cuts)


def cfg5(n: int = 10_000_000, n_rec: int = 8, dup: bool = False) -> List[Tuple[str, str]]:
    """config 5 (scaled to n bp): n_rec equal records, seed 1234 (k=63, count, 128-bit keys)."""
    b = random_acgt(n, 1234)
    if dup:
        b = b.copy()
        b[n // 2 : 2 * (n // 2)] = b[: n // 2]
    cuts = [(i * n) // n_rec for i in range(1, n_rec)]
    return split_records(b, cuts)


# ---- canonical binary form ---------------------------------------------------------------------
def key_rows(limbs) -> np.ndarray:
    """oracle limb list [hi, ..., lo] -> the device layout: u64[n], or rows of 2 / 4 limbs, least significant first."""
    if len(limbs) == 1:
        return np.ascontiguousarray(limbs[0], dtype=np.uint64)
    return np.ascontiguousarray(np.stack(limbs[::-1], axis=1), dtype=np.uint64)


def widen_rows(limbs) -> np.ndarray:
    """4-bit keys live in 128-bit device keys up to k = 32, in 256-bit ones up to k = 64."""
    want = 2 if len(limbs) <= 2 else 4
    return key_rows([np.zeros_like(limbs[0])] * (want - len(limbs)) + list(limbs))


def sha(*arrays: np.ndarray) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).view(np.uint8).reshape(-1).data)
    return h.hexdigest()


def count_digest(keys: np.ndarray, counts: np.ndarray) -> Dict[str, object]:
    """(distinct keys ascending, multiplicities) -> what the goldens store."""
    return {"rows": int(keys.shape[0]), "total": int(counts.astype(np.int64).sum()),
            "keys_sha256": sha(keys.astype(np.uint64)), "counts_sha256": sha(counts.astype(np.uint32))}


def uniq_digest(keys: np.ndarray, vals: np.ndarray) -> Dict[str, object]:
    """(singleton keys ascending, (global position << 1) | strand) -> what the goldens store."""
    return {"rows": int(keys.shape[0]), "keys_sha256": sha(keys.astype(np.uint64)),
            "vals_sha256": sha(vals.astype(np.uint64))}
