"""Helpers shared by the GPU parity tests (imported only under -m gpu)."""
import numpy as np


def first_diff(a: np.ndarray, b: np.ndarray) -> str:
    if a.shape != b.shape:
        return f"shape {a.shape} != {b.shape}"
    d = np.flatnonzero((a != b).reshape(a.shape[0], -1).any(axis=1)) if a.size else np.zeros(0, int)
    if d.size == 0:
        return "equal"
    i = int(d[0])
    return f"{d.size} rows differ, first at {i}: got {a[i]!r} want {b[i]!r}"


def limbs_to_rows(limbs) -> np.ndarray:
    """oracle limb list [hi, ..., lo] -> array comparable with Engine host views:
    8-byte keys -> (n,) uint64 ; 16- / 32-byte keys -> (n, 2) / (n, 4) uint64, least significant limb first."""
    if len(limbs) == 1:
        return limbs[0]
    assert len(limbs) in (2, 4)
    return np.stack(limbs[::-1], axis=1)


def widen(limbs):
    """4-bit oracle keys come with as many limbs as they need; the device uses 128-bit wide keys up to
    k = 32 and 256-bit ones up to k = 64."""
    want = 2 if len(limbs) <= 2 else 4
    return [np.zeros_like(limbs[0])] * (want - len(limbs)) + list(limbs)
