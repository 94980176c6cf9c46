"""kman_b200 -- a B200-native engine for kmermaid's extract -> sort -> uniq/count hot path.

The product path is hand-written CUDA (sm_100a) behind the C ABI declared in include/kmg.h
(built in-tree as kman_b200/libkmg.so); Python is the host language and mirrors the
reference's `FastaBatcher` / `Batch` / `KJoinerThreading` API and `kmer` CLI.  There is no
CPU fallback: importing the package works anywhere, but every compute call requires the
shared library and a CUDA device and fails loudly otherwise.
"""
__version__ = "0.1.0"

from kman_b200 import alphabet  # noqa: F401
