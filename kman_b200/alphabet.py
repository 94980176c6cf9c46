"""Alphabets as data.

The reference takes its alphabet from `oligo_melting.AB_NA[natype]` (two rows: allowed
symbols, their complements) and uses it at kmermaid/seq.py:318 (window filter) and
seq.py:279 (reverse complement).  That constant is not in the reference's repository
("parity unpinned" for non-ACGT symbols), so both plausible values ship as
data and the choice is a parameter; IUPAC -- upstream oligo_melting's best-known value -- is
the default.  Environment override: KMG_ALPHABET=ACGT|IUPAC.
"""
from __future__ import annotations

import os
from enum import Enum
from functools import lru_cache
from typing import Tuple

import numpy as np


class NATYPES(Enum):
    """Nucleic acid types (stand-in for oligo_melting.NATYPES used at batcher.py:45)."""

    DNA = 1
    RNA = 2


AB_TABLE = {
    "IUPAC": {
        NATYPES.DNA: ("ACGTRYKMSWBDHVN", "TGCAYRMKSWVHDBN"),
        NATYPES.RNA: ("ACGURYKMSWBDHVN", "UGCAYRMKSWVHDBN"),
    },
    "ACGT": {NATYPES.DNA: ("ACGT", "TGCA"), NATYPES.RNA: ("ACGU", "UGCA")},
}

SYMBOLS16 = "ABCDGHKMNRSTUVWY"


def default_alphabet() -> str:
    name = os.environ.get("KMG_ALPHABET", "IUPAC").upper()
    if name not in AB_TABLE:
        raise AssertionError(f"unknown alphabet {name!r}; expected one of {sorted(AB_TABLE)}")
    return name


def rows(alphabet: str, natype: NATYPES) -> Tuple[str, str]:
    if alphabet not in AB_TABLE:
        raise AssertionError(f"unknown alphabet {alphabet!r}")
    if natype not in NATYPES:
        raise AssertionError("sequence type must be from NATYPES")
    return AB_TABLE[alphabet][natype]


@lru_cache(maxsize=None)
def lut_tables(alphabet: str, natype: NATYPES) -> Tuple[np.ndarray, np.ndarray]:
    """(lut[256] uint8, comp16[16] uint8) built by libkmg's own kmg_build_lut."""
    from kman_b200 import _lib

    lib = _lib.load()
    sym, comp = rows(alphabet, natype)
    lut = np.zeros(256, np.uint8)
    c16 = np.zeros(16, np.uint8)
    _lib.check(
        lib.kmg_build_lut(sym.encode(), comp.encode(), lut.ctypes.data_as(_lib.u8p), c16.ctypes.data_as(_lib.u8p))
    )
    lut.setflags(write=False)
    c16.setflags(write=False)
    return lut, c16
