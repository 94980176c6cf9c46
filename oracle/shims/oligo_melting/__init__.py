"""Stand-in for the un-vendored third-party dependency `oligo_melting`
(2.0.1.post3 @ git 301b2c84, pinned in the reference's pyproject.toml:34 /
poetry.lock:291-310).  TEST INFRASTRUCTURE ONLY: it exists so that the
UNMODIFIED reference package under /root/reference can be imported in the build
container to generate golden vectors (oracle/gen_golden.py).  Nothing in the
product (kman_b200/) imports it.

Only the members the reference touches are provided (SURVEY.md §8c call sites:
kmermaid/seq.py:130,147,149,279,318,509; batcher.py:45,103; join.py:26).

The alphabet is the one thing the reference's sources and tests do not pin
("parity unpinned" for non-ACGT symbols).  It is therefore switchable with the
environment variable KMG_ORACLE_ALPHABET:
  IUPAC (default) -> upstream oligo_melting's best-known constants
                     AB_DNA = ["ACGTRYKMSWBDHVN", "TGCAYRMKSWVHDBN"]
  ACGT            -> strict four-letter alphabet
"""
import os
from enum import Enum


class NATYPES(Enum):
    DNA = 1
    RNA = 2


if os.environ.get("KMG_ORACLE_ALPHABET", "IUPAC").upper() == "ACGT":
    AB_DNA = ["ACGT", "TGCA"]
    AB_RNA = ["ACGU", "UGCA"]
else:
    AB_DNA = ["ACGTRYKMSWBDHVN", "TGCAYRMKSWVHDBN"]
    AB_RNA = ["ACGURYKMSWBDHVN", "UGCAYRMKSWVHDBN"]
AB_NA = {NATYPES.DNA: AB_DNA, NATYPES.RNA: AB_RNA}


class Sequence:
    def __init__(self, seq, t, name=None):
        self._text = seq.upper()
        self._len = len(self._text)
        self._natype = t
        self._ab = AB_NA[t]
        self._name = "%d-mer" % self._len if name is None else name

    @property
    def text(self):
        return self._text

    @property
    def len(self):
        return self._len

    @property
    def name(self):
        return self._name

    @property
    def natype(self):
        return self._natype

    @property
    def ab(self):
        return self._ab

    def __len__(self):
        return self._len

    def __eq__(self, other):
        return self.text == other.text and self.natype == other.natype

    __hash__ = None

    @staticmethod
    def check_ab(seq, ab):
        return all(x in ab[0] for x in set(seq))

    @staticmethod
    def mkrc(na, t):
        ab = AB_NA[t]
        table = {a: b for a, b in zip(ab[0], ab[1])}
        return "".join(table[c] for c in reversed(na.upper()))
