"""Host-side FASTA loader -> flat base buffer for the device path.

Text rules follow the reference's SmartFastaParser (kmermaid/parsers.py:53-128) and the
record-naming rule of FastaRecordBatcher.do (kmermaid/batcher.py:551):
  * everything before the first line starting with '>' is skipped; no such line ->
    AssertionError("premature end of file or empty file") (parsers.py:100-102)
  * title = header line without '>' and without trailing whitespace; name = title up to the
    first SPACE
  * sequence = the following lines, each right-stripped, concatenated, with ' ' and '\\r'
    removed; k-mers therefore span line breaks
  * `.gz` input is read through gzip (batcher.py:480)
Deliberate differences: the file is opened read-only, and an
empty record yields a zero-length record instead of hanging.

Flat layout handed to the GPU: record bytes, then one '\\n' separator after EVERY record.
The separator is never an alphabet symbol, so no window spans two records (the reference
extracts each record separately, batcher.py:387-388).
"""
from __future__ import annotations

import gzip
import os
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

SEP = 10  # '\n'
_RSTRIP_ONLY = np.array([9, 11, 12, 28, 29, 30, 31], np.uint8)  # stripped at line ends only


@dataclass
class FlatInput:
    bases: np.ndarray  # uint8, records joined by SEP (one after every record)
    rec_starts: np.ndarray  # uint64[n_rec + 1]; record r = bases[rec_starts[r] : rec_starts[r+1]-1]
    names: List[str]
    titles: List[str]

    @property
    def n_rec(self) -> int:
        return len(self.names)

    def rec_len(self, r: int) -> int:
        return int(self.rec_starts[r + 1] - self.rec_starts[r] - 1)

    def n_windows(self, k: int) -> int:
        """Number of length-k windows, valid or not (the unit of the k-mers/s metric)."""
        lens = (self.rec_starts[1:] - self.rec_starts[:-1]).astype(np.int64) - 1
        return int(np.maximum(lens - k + 1, 0).sum())


def from_records(records: Sequence[Tuple[str, str]]) -> FlatInput:
    """Build a FlatInput from (title, sequence) pairs as the reference's parser yields them."""
    titles = [t for t, _ in records]
    names = [t.split(" ")[0] for t in titles]
    starts = np.zeros(len(records) + 1, np.uint64)
    parts = []
    pos = 0
    for i, (_, s) in enumerate(records):
        starts[i] = pos
        b = s.encode("latin-1") if isinstance(s, str) else bytes(s)
        parts.append(b)
        parts.append(b"\n")
        pos += len(b) + 1
    starts[len(records)] = pos
    bases = np.frombuffer(b"".join(parts), np.uint8) if parts else np.zeros(0, np.uint8)
    return FlatInput(bases, starts, names, titles)


def _parse_slow(raw: bytes) -> FlatInput:
    """Line-by-line path for files with tabs / control characters (rstrip matters there)."""
    text = raw.decode("latin-1").replace("\r\n", "\n").replace("\r", "\n")
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    i = 0
    while i < len(lines) and not lines[i].startswith(">"):
        i += 1
    if i == len(lines):
        raise AssertionError("premature end of file or empty file")
    recs = []
    while i < len(lines):
        title = lines[i][1:].rstrip()
        i += 1
        seq = []
        while i < len(lines) and not lines[i].startswith(">"):
            seq.append(lines[i].rstrip())
            i += 1
        recs.append((title, "".join(seq).replace(" ", "").replace("\r", "")))
    return from_records(recs)


def parse_bytes(raw: bytes) -> FlatInput:
    data = np.frombuffer(raw, np.uint8)
    if data.size == 0:
        raise AssertionError("premature end of file or empty file")
    if np.isin(data, _RSTRIP_ONLY).any() or data.max() >= 0x80:
        return _parse_slow(raw)
    # '\r' ends a line exactly like '\n' does ("\r\n" merely adds an empty line)
    is_term = (data == 10) | (data == 13)
    term = np.flatnonzero(is_term)
    line_starts = np.concatenate(([0], term + 1))
    line_starts = line_starts[line_starts < data.size]
    hdr_starts = line_starts[data[line_starts] == ord(">")]
    if hdr_starts.size == 0:
        raise AssertionError("premature end of file or empty file")
    # end of each header line = first terminator at or after its start
    idx = np.searchsorted(term, hdr_starts)
    if term.size:
        hdr_ends = np.where(idx < term.size, term[np.minimum(idx, term.size - 1)], data.size)
    else:
        hdr_ends = np.full(hdr_starts.size, data.size)
    keep = ~(is_term | (data == 32))
    titles, names, parts = [], [], []
    starts = np.zeros(hdr_starts.size + 1, np.uint64)
    pos = 0
    for r in range(hdr_starts.size):
        h0, h1 = int(hdr_starts[r]), int(hdr_ends[r])
        title = raw[h0 + 1 : h1].decode("latin-1").rstrip()
        titles.append(title)
        names.append(title.split(" ")[0])
        s0 = min(h1 + 1, data.size)
        s1 = int(hdr_starts[r + 1]) if r + 1 < hdr_starts.size else data.size
        seq = data[s0:s1][keep[s0:s1]]
        starts[r] = pos
        parts.append(seq)
        parts.append(np.array([SEP], np.uint8))
        pos += seq.size + 1
    starts[hdr_starts.size] = pos
    return FlatInput(np.concatenate(parts), starts, names, titles)


def read_fasta(path: str) -> FlatInput:
    if not os.path.isfile(path):
        raise AssertionError(f"input file not found: {path}")  # batcher.py:475-476
    if path.endswith(".gz"):
        with gzip.open(path, "rb") as fh:
            raw = fh.read()
    else:
        with open(path, "rb") as fh:
            raw = fh.read()
    return parse_bytes(raw)
