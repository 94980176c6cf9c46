"""Host file plumbing with the reference's interface (kmermaid/io.py:18-71)."""
from __future__ import annotations

import gzip
import os
import shutil
import tempfile
from typing import List


def set_tempdir(path: str, create: bool = True) -> None:
    if not os.path.isdir(path):
        if not create:
            raise AssertionError(f"folder not found: {path}")
        os.makedirs(path, exist_ok=True)
    tempfile.tempdir = path


def input_file_exists(path: str) -> None:
    if not os.path.isfile(path):
        raise AssertionError(f"input file not found: {path}")


def copy_batches(batches: List, output_path: str, compress: bool = False) -> None:
    """Copy (optionally gzip) every batch's temporary file to `output_path` (io.py:46-71)."""
    for b in batches:
        if not os.path.isfile(b.tmp):
            continue
        if compress:
            dst = os.path.join(output_path, os.path.basename(b.tmp) + ".gz")
            with gzip.open(dst, "wb") as oh, open(b.tmp, "rb") as ih:
                shutil.copyfileobj(ih, oh)
        else:
            shutil.copy(b.tmp, output_path)
