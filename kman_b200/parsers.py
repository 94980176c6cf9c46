"""FASTA record iterator with the reference's interface (kmermaid/parsers.py:12-139).

`SmartFastaParser(FH).parse()` yields (title, sequence) string pairs under the reference's
text rules; the implementation is the loader of kman_b200.fasta (read-only open, terminates
on empty records -- SURVEY.md Appendix A6/A9)."""
from __future__ import annotations

import io
from typing import Iterator, Tuple

from kman_b200 import fasta


class SmartFastaParser:
    def __init__(self, FH):
        if isinstance(FH, str):
            self._path = FH
        elif isinstance(FH, io.TextIOWrapper):
            self._path = FH.name
        else:
            raise AssertionError("type error.")

    def parse(self) -> Iterator[Tuple[str, str]]:
        flat = fasta.read_fasta(self._path)
        for r, title in enumerate(flat.titles):
            b, e = int(flat.rec_starts[r]), int(flat.rec_starts[r + 1]) - 1
            yield title, flat.bases[b:e].tobytes().decode("latin-1")

    @staticmethod
    def parse_file(path: str) -> Iterator[Tuple[str, str]]:
        return SmartFastaParser(path).parse()
