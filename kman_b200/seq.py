"""Sequence / k-mer record types with the reference's interface (kmermaid/seq.py).

Same names, argument meaning and error behaviour as the reference so that its callers and
tests read the same; the per-window Python loop of `Sequence.yield_kmers`
(kmermaid/seq.py:284-328) is replaced by the GPU extraction kernel (K1+K2 behind
`kmg_extract`), and records are decoded from packed keys only when a caller iterates.

  SequenceCoords  kmermaid/seq.py:15-127    coordinates, text form "ref:start-end:strand"
  Sequence        kmermaid/seq.py:130-412   kmers / batches / kmerator / batcher
  KMer            kmermaid/seq.py:415-509   k-mer record, FASTA text form
  SequenceCount   kmermaid/seq.py:512-565   (sequence, [headers]) record, TSV text form
"""
from __future__ import annotations

import re
from enum import Enum, unique
from typing import Iterator, List, Optional, Tuple

import numpy as np

from kman_b200 import alphabet as ab
from kman_b200.alphabet import NATYPES


class SequenceCoords:
    """Window coordinates on a reference record; always on the PLUS strand, 0-based, half-open."""

    @unique
    class STRAND(Enum):
        PLUS = 0
        MINUS = 1

        @property
        def label(self) -> str:
            return "+-"[int(self.value)]

    # same grammar as the reference's pattern (seq.py:44-48): greedy ref, then start-end:strand
    regexp = re.compile(r"^(?P<ref>.+):(?P<start>[0-9]+)-(?P<end>[0-9]+):(?P<strand>[\+-])$")

    def __init__(self, ref: str, start: int, end: int, strand: "SequenceCoords.STRAND" = STRAND.PLUS):
        if start < 0 or end < 0:
            raise AssertionError
        if not isinstance(strand, SequenceCoords.STRAND):
            raise AssertionError
        self._ref, self._start, self._end, self._strand = ref, start, end, strand

    ref = property(lambda self: self._ref)
    start = property(lambda self: self._start)
    end = property(lambda self: self._end)
    strand = property(lambda self: self._strand)

    def __eq__(self, other) -> bool:
        return (
            isinstance(other, SequenceCoords)
            and self.ref == other.ref
            and self.start == other.start
            and self.end == other.end
            and self.strand == other.strand
        )

    __hash__ = None

    @staticmethod
    def rev(strand: "SequenceCoords.STRAND") -> "SequenceCoords.STRAND":
        return SequenceCoords.STRAND.MINUS if strand == SequenceCoords.STRAND.PLUS else SequenceCoords.STRAND.PLUS

    def __repr__(self) -> str:
        return "%s:%d-%d:%s" % (self.ref, self.start, self.end, self.strand.label)

    @staticmethod
    def from_str(s: str) -> "SequenceCoords":
        m = SequenceCoords.regexp.search(s)
        if m is None:
            raise AssertionError(f"incompatible string: {s}")
        strand = SequenceCoords.STRAND.PLUS if m.group("strand") == "+" else SequenceCoords.STRAND.MINUS
        return SequenceCoords(m.group("ref"), int(m.group("start")), int(m.group("end")), strand)


class _NucleicAcid:
    """The members of oligo_melting.Sequence the reference relies on."""

    def __init__(self, seq: str, t: NATYPES, name: Optional[str] = None):
        self._text = seq.upper()
        self._natype = t
        self._name = "%d-mer" % len(self._text) if name is None else name

    text = property(lambda self: self._text)
    natype = property(lambda self: self._natype)
    name = property(lambda self: self._name)
    len = property(lambda self: len(self._text))

    @property
    def ab(self) -> Tuple[str, str]:
        return ab.rows(ab.default_alphabet(), self._natype)

    def __len__(self) -> int:
        return len(self._text)

    def __eq__(self, other) -> bool:
        return self.text == other.text and self.natype == other.natype

    __hash__ = None

    @staticmethod
    def check_ab(seq: str, alphabet_rows: Tuple[str, str]) -> bool:
        return all(c in alphabet_rows[0] for c in set(seq))

    @staticmethod
    def mkrc(na: str, t: NATYPES) -> str:
        sym, comp = ab.rows(ab.default_alphabet(), t)
        table = dict(zip(sym, comp))
        return "".join(table[c] for c in reversed(na.upper()))


def decode_keys(keys: np.ndarray, k: int, wide: bool, natype: NATYPES = NATYPES.DNA) -> np.ndarray:
    """Packed keys (host view from the engine) -> (n, k) uint8 ASCII matrix."""
    bits = 4 if wide else 2
    if wide:
        sym = np.frombuffer(ab.SYMBOLS16.encode(), np.uint8)
    else:
        sym = np.frombuffer(ab.rows("ACGT", natype)[0].encode(), np.uint8)
    n = keys.shape[0]
    out = np.empty((n, k), np.uint8)
    per = 64 // bits
    two = keys.ndim == 2
    for j in range(k):
        from_end = k - 1 - j
        limb = keys[:, from_end // per] if two else keys
        sh = np.uint64(bits * (from_end % per))
        out[:, j] = sym[((limb >> sh) & np.uint64((1 << bits) - 1)).astype(np.intp)]
    return out


class Sequence(_NucleicAcid):
    """Nucleic acid sequence with k-mer generators (kmermaid/seq.py:130-412)."""

    doReverseComplement = False

    def __init__(self, seq, t, name=None):
        if not isinstance(t, NATYPES):
            raise AssertionError("sequence type must be from om.NATYPES")
        super().__init__(seq, t, name)

    def kmers(self, k: int) -> Iterator["KMer"]:
        return self.kmerator(self.text, k, self.natype, self.name, rc=self.doReverseComplement)

    def batches(self, k: int, batchSize: int) -> Iterator[Tuple[str, int]]:
        return self.batcher(self.text, k, batchSize)

    def kmers_batched(self, k: int, batchSize: int = 1) -> Iterator[Iterator["KMer"]]:
        if batchSize < 1:
            raise AssertionError
        if batchSize == 1:
            yield self.kmers(k)
        else:
            yield from self.kmerator_batched(self.text, k, self.natype, batchSize, self.name, rc=self.doReverseComplement)

    @staticmethod
    def yield_kmers(seq: str, prefix: str, k: int, t: NATYPES, offset: int, strand: SequenceCoords.STRAND,
                    rc: bool) -> Iterator["KMer"]:
        """Replaces the window loop of seq.py:284-328: one GPU extraction of both key streams,
        then records are decoded lazily in the reference's order (position, '+' before '-')."""
        from kman_b200 import fasta
        from kman_b200.engine import get_engine

        eng = get_engine()
        d = eng.upload(fasta.from_records([(prefix, seq)]), natype=t, with_names=False)
        streams = []
        a = eng.extract(d, k, rc, wide=False, val_bytes=8)
        streams.append(a)
        if a.n_other:
            streams.append(eng.extract(d, k, rc, wide=True, val_bytes=8))
        vals = np.concatenate([s.vals_host() for s in streams]) if streams else np.zeros(0, np.uint64)
        txt = np.concatenate([decode_keys(s.keys_host(), k, s.wide, t) for s in streams])
        order = np.argsort(vals, kind="stable")  # (pos << 1 | strand): emission order
        for i in order:
            pos, minus = int(vals[i] >> np.uint64(1)), int(vals[i] & np.uint64(1))
            st = SequenceCoords.rev(strand) if minus else strand
            yield KMer(prefix, pos + offset, pos + offset + k, txt[i].tobytes().decode("latin-1"), t, strand=st)

    @staticmethod
    def kmerator(seq: str, k: int, t: NATYPES, prefix: str = "ref", offset: int = 0,
                 strand: SequenceCoords.STRAND = SequenceCoords.STRAND.PLUS, rc=False) -> Iterator["KMer"]:
        return iter(Sequence.yield_kmers(seq, prefix, k, t, offset, strand, rc))

    @staticmethod
    def batcher(seq: str, k: int, batchSize: int) -> Iterator[Tuple[str, int]]:
        """Chunks with k-1 overlap (seq.py:361-383); kman_b200.dist.chunk_bases is the same rule
        in index form and is what the multi-GPU path uses."""
        start = 0
        while start < len(seq) - k + 1:
            yield (seq[start : min(len(seq), start + batchSize)], start)
            start += batchSize - k + 1

    @staticmethod
    def kmerator_batched(seq: str, k: int, t: NATYPES, batchSize: int = 1, prefix="ref",
                         rc=False) -> Iterator[Iterator["KMer"]]:
        if batchSize < 1:
            raise AssertionError
        if batchSize == 1:
            yield Sequence.kmerator(seq, k, t, prefix, rc=rc)
        for chunk, i in Sequence.batcher(seq, k, batchSize):
            yield Sequence.kmerator(chunk, k, t, prefix, offset=i, rc=rc)


class KMer(Sequence):
    """K-mer record (kmermaid/seq.py:415-509)."""

    def __init__(self, chrom: str, start: int, end: int, seq: str, t: NATYPES = NATYPES.DNA,
                 strand: SequenceCoords.STRAND = SequenceCoords.STRAND.PLUS):
        if len(seq) != end - start:
            raise AssertionError
        super().__init__(seq, t)
        self._coords = SequenceCoords(chrom, start, end, strand)

    coords = property(lambda self: self._coords)
    header = property(lambda self: str(self._coords))
    seq = property(lambda self: self.text)

    def __eq__(self, other) -> bool:
        if not self.coords == other.coords:
            return False
        return super().__eq__(other)

    __hash__ = None

    @staticmethod
    def from_fasta(record: Tuple[str, str], t: NATYPES = NATYPES.DNA) -> "KMer":
        c = SequenceCoords.from_str(record[0])
        return KMer(c.ref, c.start, c.end, record[1], t, strand=c.strand)

    from_file = from_fasta

    def as_fasta(self) -> str:
        return ">%s\n%s\n" % (self.header, self.seq)

    def __repr__(self) -> str:
        return "%s\t%s" % (self.header, self.seq)

    def is_ab_checked(self) -> bool:
        return all(c in self.ab[0] for c in set(self.text))


class SequenceCount(Sequence):
    """A sequence with the headers of all the places it occurs (kmermaid/seq.py:512-565)."""

    def __init__(self, seq, headers, t=NATYPES.DNA):
        super().__init__(seq, t)
        if not all(isinstance(h, str) for h in headers):
            raise AssertionError
        self.__headers: List[str] = headers

    @property
    def header(self) -> List[str]:
        return self.__headers.copy()

    seq = property(lambda self: self.text)

    @staticmethod
    def from_text(line: str, t: NATYPES = NATYPES.DNA) -> "SequenceCount":
        seq, headers = line.strip().split("\t")
        return SequenceCount(seq, headers.split(" "), t)

    from_file = from_text

    def __repr__(self) -> str:
        return "%s\t%s" % (self.seq, " ".join(self.header))

    def as_text(self) -> str:
        return str(self) + "\n"
