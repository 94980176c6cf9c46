"""Device pipeline: drives libkmg's stage API with torch-owned buffers.

torch is plumbing only (device memory, streams, torch.distributed); every byte of the hot
path is produced by the CUDA kernels behind include/kmg.h.  There is no CPU fallback: the
constructor raises if the shared library or a CUDA device is missing.

Stage -> reference mapping:
  extract()        seq.py:284-328 (+ rc :245-282)           K1+K2
  sort()           batch.py:156-168 + join.py:63-93         K3
  rle_count()      join.py:95-130 + :265-285                K4
  singletons()     join.py:95-130 + :243-263                K4
  count_text()/uniq_text()  join.py:262,284; seq.py:103-104 K6 (+ narrow/wide rank merge)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from kman_b200 import _lib, alphabet as ab
from kman_b200.fasta import FlatInput


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _on_device(fn):
    """Run an Engine method with the engine's GPU current: libkmg launches on the current device
    (kernel attributes, workspaces and streams are per device), and one process may own several
    engines (get_engine(device))."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        if torch.cuda.current_device() == self.device.index:
            return fn(self, *args, **kwargs)
        with torch.cuda.device(self.device):
            return fn(self, *args, **kwargs)

    return wrapper


@dataclass
class DeviceInput:
    """Flat base buffer resident in HBM."""

    bases: torch.Tensor  # uint8[n_bases (+pad)]
    n_bases: int
    flat: FlatInput
    alphabet: str
    natype: ab.NATYPES
    lut: torch.Tensor
    comp16: torch.Tensor
    rec_starts: Optional[torch.Tensor] = None  # uint64 as int64 [n_rec+1]
    names_buf: Optional[torch.Tensor] = None
    name_offs: Optional[torch.Tensor] = None
    pos_offset: int = 0  # global position of bases[0] (multi-GPU chunks)

    @property
    def rna(self) -> int:
        return 1 if self.natype == ab.NATYPES.RNA else 0


@dataclass
class KeyArray:
    """n keys (+ optional payload) in `keys`/`vals`; `keys_alt`/`vals_alt` are same-sized scratch."""

    keys: torch.Tensor  # uint8 bytes
    keys_alt: Optional[torch.Tensor]
    vals: Optional[torch.Tensor]
    vals_alt: Optional[torch.Tensor]
    n: int
    key_bytes: int
    val_bytes: int
    k: int
    wide: bool
    n_other: int = 0  # narrow extraction: windows that belong to the wide stream
    is_sorted: bool = False
    hist: Optional[torch.Tensor] = None  # digit histograms of exactly these keys (from kmg_extract)

    @property
    def key_bits(self) -> int:
        return self.k * (4 if self.wide else 2)

    def keys_host(self) -> np.ndarray:
        raw = self.keys[: self.n * self.key_bytes].cpu().numpy()
        return raw.view(np.uint64) if self.key_bytes == 8 else raw.view(np.uint64).reshape(-1, self.key_bytes // 8)

    def vals_host(self) -> Optional[np.ndarray]:
        if self.vals is None:
            return None
        raw = self.vals[: self.n * self.val_bytes].cpu().numpy()
        return raw.view(np.uint32 if self.val_bytes == 4 else np.uint64)


@dataclass
class CountTable:
    keys: torch.Tensor
    counts: torch.Tensor  # uint32 as bytes
    n: int
    key_bytes: int
    k: int
    wide: bool

    def keys_host(self) -> np.ndarray:
        raw = self.keys[: self.n * self.key_bytes].cpu().numpy()
        return raw.view(np.uint64) if self.key_bytes == 8 else raw.view(np.uint64).reshape(-1, self.key_bytes // 8)

    def counts_host(self) -> np.ndarray:
        return self.counts[: self.n * 4].cpu().numpy().view(np.uint32)


class Engine:
    """One engine per (process, GPU)."""

    def __init__(self, device: Optional[int] = None):
        self.lib = _lib.load()  # ImportError if libkmg.so is missing
        if not torch.cuda.is_available():
            raise _lib.KmgError("no CUDA device visible: kman_b200 has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self._scratch: Dict[str, torch.Tensor] = {}
        self._small = torch.zeros(8, dtype=torch.int64, device=self.device)
        self._luts: Dict[Tuple[str, ab.NATYPES], Tuple[torch.Tensor, torch.Tensor]] = {}

    # ---- plumbing ---------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _buf(self, name: str, nbytes: int) -> torch.Tensor:
        """Named scratch buffer, grown geometrically and reused across calls."""
        t = self._scratch.get(name)
        if t is None or t.numel() < nbytes:
            self._scratch.pop(name, None)
            t = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
            self._scratch[name] = t
        return t

    def _new(self, nbytes: int) -> torch.Tensor:
        return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=self.device)

    def _status(self, ws: torch.Tensor) -> None:
        _lib.check(self.lib.kmg_ws_status(ws.data_ptr(), self._stream()))

    def lut_tensors(self, alphabet: str, natype: ab.NATYPES) -> Tuple[torch.Tensor, torch.Tensor]:
        key = (alphabet, natype)
        if key not in self._luts:
            lut, c16 = ab.lut_tables(alphabet, natype)
            self._luts[key] = (
                torch.from_numpy(lut.copy()).to(self.device),
                torch.from_numpy(c16.copy()).to(self.device),
            )
        return self._luts[key]

    # ---- upload -----------------------------------------------------------------------------
    @_on_device
    def upload(self, flat: FlatInput, alphabet: Optional[str] = None, natype: ab.NATYPES = ab.NATYPES.DNA,
               with_names: bool = True) -> DeviceInput:
        alphabet = alphabet or ab.default_alphabet()
        lut, c16 = self.lut_tensors(alphabet, natype)
        n = int(flat.bases.shape[0])
        bases = torch.empty(((n + 15) // 16 + 1) * 16, dtype=torch.uint8, device=self.device)
        if n:
            src = flat.bases if flat.bases.flags.writeable and flat.bases.flags.c_contiguous else np.array(flat.bases)
            bases[:n].copy_(torch.from_numpy(src), non_blocking=False)
        d = DeviceInput(bases, n, flat, alphabet, natype, lut, c16)
        if with_names:
            self._upload_names(d)
        return d

    @_on_device
    def load_fasta(self, path: str, alphabet: Optional[str] = None, natype: ab.NATYPES = ab.NATYPES.DNA,
                   with_names: bool = True) -> DeviceInput:
        """Read a FASTA file and flatten it ON THE GPU (kmg_fasta_flatten): the raw bytes go to
        the device once, line structure / header lines / blanks are resolved by three kernels.
        Files with TAB / VT / FF / control separators or non-ASCII bytes take the host loader,
        which implements the reference's rstrip() rule for them."""
        import gzip
        import os

        from kman_b200 import fasta

        if not os.path.isfile(path):
            raise AssertionError(f"input file not found: {path}")  # batcher.py:475-476
        opener = gzip.open if path.endswith(".gz") else open
        with opener(path, "rb") as fh:
            raw = fh.read()
        if len(raw) == 0:
            raise AssertionError("premature end of file or empty file")
        n_raw = len(raw)
        d_raw = torch.empty(((n_raw + 15) // 16 + 1) * 16, dtype=torch.uint8, device=self.device)
        import warnings

        with warnings.catch_warnings():  # (a read-only view of the bytes object: no second host copy)
            warnings.simplefilter("ignore", UserWarning)
            d_raw[:n_raw].copy_(torch.from_numpy(np.frombuffer(raw, np.uint8)))
        bases = torch.empty(((n_raw + 15) // 16 + 2) * 16, dtype=torch.uint8, device=self.device)
        max_rec = 1 << 16
        while True:
            rec_starts = torch.zeros(max_rec + 1, dtype=torch.int64, device=self.device)
            hdr_begin = torch.zeros(max_rec, dtype=torch.int64, device=self.device)
            ws_bytes = self.lib.kmg_fasta_workspace_bytes(n_raw)
            ws = self._buf("ws_fasta", ws_bytes)
            _lib.check(self.lib.kmg_fasta_flatten(d_raw.data_ptr(), n_raw, bases.data_ptr(), rec_starts.data_ptr(),
                                                  hdr_begin.data_ptr(), max_rec, self._small[5:].data_ptr(),
                                                  ws.data_ptr(), ws_bytes, self._stream()))
            n_out, n_rec, special = (int(x) for x in self._small[5:8].cpu().numpy().view(np.uint64))
            if special:
                return self.upload(fasta.parse_bytes(raw), alphabet, natype, with_names)
            if n_rec <= max_rec:
                break
            max_rec = n_rec
        if n_rec == 0:
            raise AssertionError("premature end of file or empty file")  # parsers.py:100-102
        hb = hdr_begin[:n_rec].cpu().numpy()
        titles = []
        for b in hb:
            b = int(b)
            e1, e2 = raw.find(b"\n", b), raw.find(b"\r", b)
            e = min(x for x in (e1, e2, n_raw) if x >= 0)
            titles.append(raw[b:e].decode("latin-1").rstrip())
        names = [t.split(" ")[0] for t in titles]
        flat = FlatInput(bases[:n_out].cpu().numpy(), rec_starts[: n_rec + 1].cpu().numpy().astype(np.uint64), names, titles)
        alphabet = alphabet or ab.default_alphabet()
        lut, c16 = self.lut_tensors(alphabet, natype)
        d = DeviceInput(bases, n_out, flat, alphabet, natype, lut, c16)
        if with_names:
            self._upload_names(d)
        return d

    def _upload_names(self, d: DeviceInput) -> None:
        flat = d.flat
        d.rec_starts = torch.from_numpy(flat.rec_starts.astype(np.int64)).to(self.device)
        nb = [nm.encode("latin-1") for nm in flat.names]
        offs = np.zeros(len(nb) + 1, np.int64)
        offs[1:] = np.cumsum([len(b) for b in nb])
        d.name_offs = torch.from_numpy(offs).to(self.device)
        d.names_buf = torch.from_numpy(np.frombuffer(b"".join(nb) + b"\0", np.uint8).copy()).to(self.device)

    @_on_device
    def count_windows(self, d: DeviceInput, k: int, rc: bool = False) -> Tuple[int, int]:
        """(narrow keys, wide-stream windows) of an input WITHOUT building a key: for k >= 12 the 4-mer
        histogram pre-pass of the bases knows both (kmg_extract_dest_counts with one destination);
        smaller k runs the extraction."""
        n_win = max(0, d.n_bases - k + 1)
        if n_win == 0:
            return 0, 0
        if k < 12:
            a = self.extract(d, k, rc, wide=False, val_bytes=0)
            return a.n, a.n_other
        counts = torch.zeros(2, dtype=torch.int64, device=self.device)
        ws_bytes = self.lib.kmg_dest_counts_workspace_bytes()
        ws = self._buf("ws_dest_counts", ws_bytes)
        _lib.check(self.lib.kmg_extract_dest_counts(d.bases.data_ptr(), d.n_bases, 0, n_win, k, int(rc), d.lut.data_ptr(), 1,
                                                    counts.data_ptr(), ws.data_ptr(), ws_bytes, self._stream()))
        n, n_other = (int(x) for x in counts.cpu())
        return n, n_other

    # ---- K1+K2 ------------------------------------------------------------------------------
    @_on_device
    def extract(self, d: DeviceInput, k: int, rc: bool = False, wide: bool = False, val_bytes: int = 0,
                win_begin: int = 0, win_end: Optional[int] = None, reuse: Optional[str] = None,
                want_hist: bool = False) -> KeyArray:
        """Keys (and payload) of the windows starting in [win_begin, win_end), emission order.

        `reuse`: name prefix of engine-owned scratch buffers to place the output in (the
        benchmark loop) instead of fresh allocations."""
        if k <= 1:
            raise AssertionError(f"k must be >= 1, got {k} instead.")  # batcher.py:477-478
        if k > 64:
            raise ValueError(f"k={k}: this build supports k <= 64 (no CPU fallback)")
        n_win_total = max(0, d.n_bases - k + 1)
        win_end = n_win_total if win_end is None else min(win_end, n_win_total)
        win_begin = min(win_begin, win_end)
        n_win = win_end - win_begin
        # narrow: 2-bit codes, 8 / 16 bytes; wide: 4-bit codes, 16 bytes up to k = 32, 32 bytes up to k = 64
        kb = (16 if k <= 32 else 32) if wide else (8 if k <= 32 else 16)
        cap = max(n_win * (2 if rc else 1), 1)
        mk = (lambda nm, nb: self._buf(reuse + nm, nb)) if reuse else (lambda nm, nb: self._new(nb))
        keys = mk("keys", cap * kb)
        keys_alt = mk("keys_alt", cap * kb)
        vals = mk("vals", cap * val_bytes) if val_bytes else None
        vals_alt = mk("vals_alt", cap * val_bytes) if val_bytes else None
        ws_bytes = self.lib.kmg_extract_workspace_bytes(n_win)
        ws = self._buf("ws_extract", ws_bytes)
        # fused digit histograms for the sort that follows (narrow stream only)
        hist = None
        if want_hist and not wide and k >= 4:
            hist = self._buf((reuse or "") + "hist", 16 * 256 * 8) if reuse else self._new(16 * 256 * 8)
        _lib.check(
            self.lib.kmg_extract(
                d.bases.data_ptr(), d.n_bases, win_begin, win_end, k, int(rc), int(wide), d.lut.data_ptr(),
                d.comp16.data_ptr(), keys.data_ptr(), kb, _ptr(vals), val_bytes, d.pos_offset,
                self._small.data_ptr(), _ptr(hist), ws.data_ptr(), ws_bytes, self._stream(),
            )
        )
        self._status(ws)
        cnt = self._small[:2].cpu().numpy().view(np.uint64)
        return KeyArray(keys, keys_alt, vals, vals_alt, int(cnt[0]), kb, val_bytes, k, wide, n_other=int(cnt[1]),
                        hist=hist)

    # ---- K3 -----------------------------------------------------------------------------------
    @_on_device
    def sort(self, a: KeyArray, begin_bit: int = 0, end_bit: Optional[int] = None) -> KeyArray:
        end_bit = a.key_bits if end_bit is None else end_bit
        if a.key_bytes == 32:
            return self._sort256(a, end_bit)
        if a.n > 1 and end_bit > begin_bit:
            ws_bytes = self.lib.kmg_radix_sort_workspace_bytes(a.n, a.key_bytes, a.val_bytes, begin_bit, end_bit)
            ws = self._buf("ws_sort", ws_bytes)
            sel = C.c_int(0)
            hist = a.hist if (begin_bit == 0 and end_bit == a.key_bits) else None
            _lib.check(
                self.lib.kmg_radix_sort(
                    a.keys.data_ptr(), a.keys_alt.data_ptr(), _ptr(a.vals), _ptr(a.vals_alt), a.n, a.key_bytes,
                    a.val_bytes, begin_bit, end_bit, _ptr(hist), C.byref(sel), ws.data_ptr(), ws_bytes, self._stream(),
                )
            )
            a.hist = None
            self._last_sort_ws = ws
            if sel.value:
                a.keys, a.keys_alt = a.keys_alt, a.keys
                a.vals, a.vals_alt = a.vals_alt, a.vals
        a.is_sorted = True
        return a

    def _sort256(self, a: KeyArray, end_bit: int) -> KeyArray:
        """256-bit keys (wide stream, 33 <= k <= 64): kmg_sort256 sorts out of place into the alt buffers."""
        if a.n > 1 and end_bit > 0:
            ws_bytes = self.lib.kmg_sort256_workspace_bytes(a.n)
            ws = self._buf("ws_sort256", ws_bytes)
            _lib.check(self.lib.kmg_sort256(a.keys.data_ptr(), a.keys_alt.data_ptr(), _ptr(a.vals), _ptr(a.vals_alt), a.n,
                                            a.val_bytes, end_bit, ws.data_ptr(), ws_bytes, self._stream()))
            a.keys, a.keys_alt = a.keys_alt, a.keys
            a.vals, a.vals_alt = a.vals_alt, a.vals
        a.is_sorted = True
        return a

    @_on_device
    def sort_count(self, a: KeyArray, end_bit: Optional[int] = None, reuse: Optional[str] = None) -> CountTable:
        """sort() + rle_count() in one native call (kmg_sort_count): when the hybrid finish applies,
        its local sort emits the (k-mer, count) table directly.  Consumes `a` (both key buffers)."""
        end_bit = a.key_bits if end_bit is None else end_bit
        assert a.val_bytes == 0
        if a.key_bytes == 32:
            return self.rle_count(self.sort(a, 0, end_bit), reuse=reuse)
        counts = self._buf(reuse + "counts", a.n * 4) if reuse else self._new(a.n * 4)
        if a.n == 0:
            return CountTable(a.keys_alt, counts, 0, a.key_bytes, a.k, a.wide)
        ws_bytes = self.lib.kmg_sort_count_workspace_bytes(a.n, a.key_bytes, end_bit)
        ws = self._buf("ws_sort", ws_bytes)
        sel = C.c_int(0)
        hist = a.hist if end_bit == a.key_bits else None
        _lib.check(
            self.lib.kmg_sort_count(a.keys.data_ptr(), a.keys_alt.data_ptr(), a.n, a.key_bytes, end_bit, _ptr(hist),
                                    counts.data_ptr(), self._small[2:].data_ptr(), C.byref(sel), ws.data_ptr(), ws_bytes,
                                    self._stream())
        )
        a.hist = None
        self._last_sort_ws = ws
        n_out = self._sort_result_rows(ws)
        return CountTable(a.keys_alt if sel.value else a.keys, counts, n_out, a.key_bytes, a.k, a.wide)

    def _sort_result_rows(self, ws: torch.Tensor) -> int:
        """Rows of the table kmg_sort_count / kmg_sort_uniq just produced.  When the hybrid finish ran,
        the call already synchronised and read the count and the status word back together
        (kmg_get_stat "n_out" / "ws_err"): no second round trip."""
        n_out, err = int(self.lib.kmg_get_stat(b"n_out")), int(self.lib.kmg_get_stat(b"ws_err"))
        if n_out >= 0 and err == 0:
            return n_out
        self._status(ws)
        return int(self._small[2:3].cpu().numpy().view(np.uint64)[0])

    # ---- K4 -----------------------------------------------------------------------------------
    @_on_device
    def rle_count(self, a: KeyArray, reuse: Optional[str] = None) -> CountTable:
        """Distinct keys + counts.  The distinct keys are written into `a.keys_alt`."""
        assert a.is_sorted
        counts = self._buf(reuse + "counts", a.n * 4) if reuse else self._new(a.n * 4)
        if a.n == 0:
            return CountTable(a.keys_alt, counts, 0, a.key_bytes, a.k, a.wide)
        ws_bytes = self.lib.kmg_rle_workspace_bytes(a.n)
        ws = self._buf("ws_rle", ws_bytes)
        _lib.check(
            self.lib.kmg_rle_count(a.keys.data_ptr(), a.n, a.key_bytes, a.keys_alt.data_ptr(), counts.data_ptr(),
                                   self._small[2:].data_ptr(), ws.data_ptr(), ws_bytes, self._stream())
        )
        self._status(ws)
        n_out = int(self._small[2:3].cpu().numpy().view(np.uint64)[0])
        return CountTable(a.keys_alt, counts, n_out, a.key_bytes, a.k, a.wide)

    @_on_device
    def sort_uniq(self, a: KeyArray, end_bit: Optional[int] = None) -> KeyArray:
        """sort() + singletons() in one native call (kmg_sort_uniq).  Repeated keys are dropped, so
        their order is irrelevant and the keys take the hybrid finish with the payload.  Consumes `a`."""
        end_bit = a.key_bits if end_bit is None else end_bit
        assert a.val_bytes in (4, 8)
        if a.key_bytes == 32:
            return self.singletons(self.sort(a, 0, end_bit))
        if a.n == 0:
            return KeyArray(a.keys_alt, None, a.vals_alt, None, 0, a.key_bytes, a.val_bytes, a.k, a.wide, is_sorted=True)
        ws_bytes = self.lib.kmg_sort_uniq_workspace_bytes(a.n, a.key_bytes, a.val_bytes, end_bit)
        ws = self._buf("ws_sort", ws_bytes)
        sel = C.c_int(0)
        hist = a.hist if end_bit == a.key_bits else None
        _lib.check(
            self.lib.kmg_sort_uniq(a.keys.data_ptr(), a.keys_alt.data_ptr(), a.vals.data_ptr(), a.vals_alt.data_ptr(), a.n,
                                   a.key_bytes, a.val_bytes, end_bit, _ptr(hist), self._small[2:].data_ptr(), C.byref(sel),
                                   ws.data_ptr(), ws_bytes, self._stream())
        )
        a.hist = None
        self._last_sort_ws = ws
        n_out = self._sort_result_rows(ws)
        k_, v_ = (a.keys_alt, a.vals_alt) if sel.value else (a.keys, a.vals)
        return KeyArray(k_, None, v_, None, n_out, a.key_bytes, a.val_bytes, a.k, a.wide, is_sorted=True)

    @_on_device
    def singletons(self, a: KeyArray) -> KeyArray:
        """Keys (+payload) that occur exactly once, ascending (written into the alt buffers)."""
        assert a.is_sorted
        if a.n == 0:
            return KeyArray(a.keys_alt, None, a.vals_alt, None, 0, a.key_bytes, a.val_bytes, a.k, a.wide, is_sorted=True)
        ws_bytes = self.lib.kmg_rle_workspace_bytes(a.n)
        ws = self._buf("ws_rle", ws_bytes)
        _lib.check(
            self.lib.kmg_select_singletons(a.keys.data_ptr(), _ptr(a.vals), a.n, a.key_bytes, a.val_bytes,
                                           a.keys_alt.data_ptr(), _ptr(a.vals_alt), self._small[2:].data_ptr(),
                                           ws.data_ptr(), ws_bytes, self._stream())
        )
        self._status(ws)
        n_out = int(self._small[2:3].cpu().numpy().view(np.uint64)[0])
        return KeyArray(a.keys_alt, None, a.vals_alt, None, n_out, a.key_bytes, a.val_bytes, a.k, a.wide, is_sorted=True)

    # ---- K5 -----------------------------------------------------------------------------------
    @_on_device
    def range_partition(self, a: KeyArray, n_parts: int) -> Tuple[KeyArray, np.ndarray]:
        """Stable split by the top key bits into n_parts regions (into the alt buffers)."""
        ws_bytes = self.lib.kmg_partition_workspace_bytes(max(a.n, 1), a.key_bytes, a.val_bytes)
        ws = self._buf("ws_sort", ws_bytes)
        pc = torch.zeros(n_parts, dtype=torch.int64, device=self.device)
        _lib.check(
            self.lib.kmg_range_partition(a.keys.data_ptr(), _ptr(a.vals), a.n, a.key_bytes, a.val_bytes, a.key_bits,
                                         n_parts, a.keys_alt.data_ptr(), _ptr(a.vals_alt), pc.data_ptr(),
                                         ws.data_ptr(), ws_bytes, self._stream())
        )
        if a.n:
            self._status(ws)
        a.keys, a.keys_alt = a.keys_alt, a.keys
        a.vals, a.vals_alt = a.vals_alt, a.vals
        return a, pc.cpu().numpy().astype(np.int64)

    # ---- K6 -----------------------------------------------------------------------------------
    @_on_device
    def format_counts(self, t: CountTable, rna: int = 0) -> torch.Tensor:
        """Device text "SEQ\\tCOUNT\\n" for one stream (uint8 tensor, exact length)."""
        if t.n == 0:
            return torch.empty(0, dtype=torch.uint8, device=self.device)
        text = self._new(t.n * (t.k + 12))
        ws_bytes = self.lib.kmg_format_workspace_bytes(t.n)
        ws = self._buf("ws_fmt", ws_bytes)
        _lib.check(
            self.lib.kmg_format_counts(t.keys.data_ptr(), t.counts.data_ptr(), t.n, t.key_bytes, t.k, int(t.wide), rna,
                                       text.data_ptr(), self._small[4:].data_ptr(), ws.data_ptr(), ws_bytes,
                                       self._stream())
        )
        self._status(ws)
        nbytes = int(self._small[4:5].cpu().numpy().view(np.uint64)[0])
        return text[:nbytes]

    @_on_device
    def format_uniq(self, s: KeyArray, d: DeviceInput) -> torch.Tensor:
        """Device text ">NAME:START-END:STRAND\\nSEQ\\n" for one stream."""
        if s.n == 0:
            return torch.empty(0, dtype=torch.uint8, device=self.device)
        if d.rec_starts is None:
            self._upload_names(d)
        max_name = max((len(nm) for nm in d.flat.names), default=0)
        text = self._new(s.n * (s.k + 48 + max_name))
        ws_bytes = self.lib.kmg_format_workspace_bytes(s.n)
        ws = self._buf("ws_fmt", ws_bytes)
        _lib.check(
            self.lib.kmg_format_uniq(s.keys.data_ptr(), s.vals.data_ptr(), s.n, s.key_bytes, s.val_bytes, s.k,
                                     int(s.wide), d.rna, d.rec_starts.data_ptr(), d.flat.n_rec, d.names_buf.data_ptr(),
                                     d.name_offs.data_ptr(), text.data_ptr(), self._small[4:].data_ptr(),
                                     ws.data_ptr(), ws_bytes, self._stream())
        )
        self._status(ws)
        nbytes = int(self._small[4:5].cpu().numpy().view(np.uint64)[0])
        return text[:nbytes]

    @_on_device
    def merge_ranks(self, narrow_keys: torch.Tensor, n_narrow: int, wide_keys: torch.Tensor, n_wide: int, k: int,
                    rna: int) -> Tuple[torch.Tensor, torch.Tensor]:
        rn = torch.zeros(max(n_narrow, 1), dtype=torch.int64, device=self.device)
        rw = torch.zeros(max(n_wide, 1), dtype=torch.int64, device=self.device)
        if k <= 32:
            _lib.check(
                self.lib.kmg_merge_ranks(narrow_keys.data_ptr(), n_narrow, 8, wide_keys.data_ptr(), n_wide, 16, k, rna,
                                         rn.data_ptr(), rw.data_ptr(), self._stream())
            )
        else:  # 16-byte narrow keys, 32-byte wide keys
            tmp = self._new(max(n_wide, 1) * 16)
            _lib.check(
                self.lib.kmg_merge_ranks_wide(narrow_keys.data_ptr(), n_narrow, wide_keys.data_ptr(), n_wide, k, rna,
                                              rn.data_ptr(), rw.data_ptr(), tmp.data_ptr(), self._stream())
            )
        return rn[:n_narrow], rw[:n_wide]

    # ---- whole path on one GPU ------------------------------------------------------------------
    def sorted_streams(self, d: DeviceInput, k: int, rc: bool, val_bytes: int) -> List[KeyArray]:
        """[narrow] or [narrow, wide]: extracted and sorted key arrays of the input."""
        narrow = self.sort(self.extract(d, k, rc, wide=False, val_bytes=val_bytes, want_hist=True))
        out = [narrow]
        if narrow.n_other:
            out.append(self.sort(self.extract(d, k, rc, wide=True, val_bytes=val_bytes)))
        return out

    def _pipeline_buffers(self, d: DeviceInput, k: int, rc: bool, val_bytes: int, reuse: Optional[str],
                          win_begin: int, win_end: Optional[int]):
        if k <= 1:
            raise AssertionError(f"k must be >= 1, got {k} instead.")  # batcher.py:477-478
        if k > 64:
            raise ValueError(f"k={k}: this build supports k <= 64 (no CPU fallback)")
        n_win_total = max(0, d.n_bases - k + 1)
        win_end = n_win_total if win_end is None else min(win_end, n_win_total)
        win_begin = min(win_begin, win_end)
        kb = 16 if k > 32 else 8
        cap = max((win_end - win_begin) * (2 if rc else 1), 1)
        mk = (lambda nm, nb: self._buf(reuse + nm, nb)) if reuse else (lambda nm, nb: self._new(nb))
        keys, keys_alt = mk("keys", cap * kb), mk("keys_alt", cap * kb)
        vals = mk("vals", cap * val_bytes) if val_bytes else None
        vals_alt = mk("vals_alt", cap * val_bytes) if val_bytes else None
        ws_bytes = self.lib.kmg_pipeline_workspace_bytes(win_end - win_begin, k, int(rc), val_bytes)
        ws = self._buf("ws_pipeline", ws_bytes)
        return win_begin, win_end, kb, cap, keys, keys_alt, vals, vals_alt, ws, ws_bytes

    @_on_device
    def count_narrow(self, d: DeviceInput, k: int, rc: bool = False, reuse: Optional[str] = None, win_begin: int = 0,
                     win_end: Optional[int] = None) -> Tuple[CountTable, int]:
        """(k-mer, count) table of the narrow stream in ONE native call (kmg_extract_sort_count: the
        extraction kernel is the sort's first prefix pass), and the number of windows that belong
        to the wide stream.  seq.py:284-328 + batch.py:156-168 + join.py:63-130,265-285."""
        win_begin, win_end, kb, cap, keys, keys_alt, _, _, ws, ws_bytes = self._pipeline_buffers(
            d, k, rc, 0, reuse, win_begin, win_end)
        counts = self._buf(reuse + "counts", cap * 4) if reuse else self._new(cap * 4)
        res = (C.c_uint64 * 4)()
        _lib.check(self.lib.kmg_extract_sort_count(d.bases.data_ptr(), d.n_bases, win_begin, win_end, k, int(rc),
                                                   d.lut.data_ptr(), keys.data_ptr(), keys_alt.data_ptr(),
                                                   counts.data_ptr(), res, ws.data_ptr(), ws_bytes, self._stream()))
        return CountTable(keys_alt if res[3] else keys, counts, int(res[0]), kb, k, False), int(res[2])

    @_on_device
    def uniq_narrow(self, d: DeviceInput, k: int, rc: bool = False, reuse: Optional[str] = None, win_begin: int = 0,
                    win_end: Optional[int] = None, val_bytes: Optional[int] = None) -> Tuple[KeyArray, int]:
        """Singleton keys with their (position << 1 | strand) payload of the narrow stream in ONE native
        call (kmg_extract_sort_uniq), and the number of wide-stream windows.  join.py:243-263."""
        vb = val_bytes or (4 if ((d.pos_offset + d.n_bases) << 1) < (1 << 32) else 8)
        win_begin, win_end, kb, cap, keys, keys_alt, vals, vals_alt, ws, ws_bytes = self._pipeline_buffers(
            d, k, rc, vb, reuse, win_begin, win_end)
        res = (C.c_uint64 * 4)()
        _lib.check(self.lib.kmg_extract_sort_uniq(d.bases.data_ptr(), d.n_bases, win_begin, win_end, k, int(rc),
                                                  d.lut.data_ptr(), keys.data_ptr(), keys_alt.data_ptr(), vals.data_ptr(),
                                                  vals_alt.data_ptr(), vb, d.pos_offset, res, ws.data_ptr(), ws_bytes,
                                                  self._stream()))
        k_, v_ = (keys_alt, vals_alt) if res[3] else (keys, vals)
        return KeyArray(k_, None, v_, None, int(res[0]), kb, vb, k, False, is_sorted=True), int(res[2])

    def count(self, d: DeviceInput, k: int, rc: bool = False) -> List[CountTable]:
        """`kmer count` (SEQ_COUNT) on device: one CountTable per stream."""
        tab, n_other = self.count_narrow(d, k, rc)
        out = [tab]
        if n_other:
            out.append(self.sort_count(self.extract(d, k, rc, wide=True, val_bytes=0)))
        return out

    def uniq(self, d: DeviceInput, k: int, rc: bool = False) -> List[KeyArray]:
        """`kmer uniq` on device: singleton keys with payload, one KeyArray per stream."""
        vb = 4 if ((d.pos_offset + d.n_bases) << 1) < (1 << 32) else 8
        sing, n_other = self.uniq_narrow(d, k, rc, val_bytes=vb)
        out = [sing]
        if n_other:
            out.append(self.sort_uniq(self.extract(d, k, rc, wide=True, val_bytes=vb)))
        return out

    # ---- text (interleaves the narrow and the wide stream in ASCII order) -----------------------
    def _interleave(self, texts: List[torch.Tensor], keys: List[torch.Tensor], ns: List[int], k: int, rna: int,
                    line_lens: Optional[List[torch.Tensor]]) -> bytes:
        if len(texts) == 1 or ns[1] == 0:
            return texts[0].cpu().numpy().tobytes()
        if ns[0] == 0:
            return texts[1].cpu().numpy().tobytes()
        rn, rw = self.merge_ranks(keys[0], ns[0], keys[1], ns[1], k, rna)
        # merged position of every line, then a stable host-side interleave of the two texts
        pos_n = (torch.arange(ns[0], device=self.device) + rn).cpu().numpy()
        pos_w = (torch.arange(ns[1], device=self.device) + rw).cpu().numpy()
        tn, tw = texts[0].cpu().numpy(), texts[1].cpu().numpy()

        def line_starts(t: np.ndarray, per_rec: int) -> np.ndarray:
            nl = np.flatnonzero(t == 10)
            ends = nl[per_rec - 1 :: per_rec] + 1
            return np.concatenate(([0], ends))

        per = 1 if line_lens is None else 2
        sn, sw = line_starts(tn, per), line_starts(tw, per)
        total = ns[0] + ns[1]
        lens = np.empty(total, np.int64)
        lens[pos_n] = np.diff(sn)
        lens[pos_w] = np.diff(sw)
        offs = np.concatenate(([0], np.cumsum(lens)))
        out = np.empty(int(offs[-1]), np.uint8)
        # wide lines are few: copy the narrow text in runs between them
        src_is_wide = np.zeros(total, bool)
        src_is_wide[pos_w] = True
        # narrow runs
        idx_n = 0
        prev = 0
        for j, pw in enumerate(pos_w):
            run = pw - prev  # narrow lines before this wide line
            if run:
                out[offs[prev] : offs[pw]] = tn[sn[idx_n] : sn[idx_n + run]]
                idx_n += run
            out[offs[pw] : offs[pw + 1]] = tw[sw[j] : sw[j + 1]]
            prev = pw + 1
        if prev < total:
            out[offs[prev] :] = tn[sn[idx_n] :]
        return out.tobytes()

    @_on_device
    def abundance(self, d: DeviceInput, k: int, rc: bool = False, masked: bool = False) -> Tuple[np.ndarray, np.ndarray]:
        """VEC_COUNT / VEC_COUNT_MASKED (join.py:287-335): per flat position the number of occurrences of the
        k-mer that starts there ('+' vector) / of its reverse complement ('-' vector); masked: occurrences
        in other records only.  Both streams are sorted stably with their payload, then every element
        scatters its group's size (kmg_abundance_scatter).  Returns host uint32 arrays over the flat buffer."""
        vb = 4 if ((d.pos_offset + d.n_bases) << 1) < (1 << 32) else 8
        if d.rec_starts is None:
            self._upload_names(d)
        out = torch.zeros(2 * (d.n_bases + 1), dtype=torch.int32, device=self.device)
        plus, minus = out[: d.n_bases + 1], out[d.n_bases + 1:]
        err = torch.zeros(1, dtype=torch.int32, device=self.device)
        for s in self.sorted_streams(d, k, rc, vb):
            if s.n == 0:
                continue
            _lib.check(self.lib.kmg_abundance_scatter(s.keys.data_ptr(), s.vals.data_ptr(), s.n, s.key_bytes, s.val_bytes,
                                                      d.rec_starts.data_ptr(), d.flat.n_rec, int(masked), d.pos_offset,
                                                      plus.data_ptr(), minus.data_ptr(), err.data_ptr(), self._stream()))
        if int(err.item()):
            raise ValueError("a k-mer occurs more than 2^32-1 times: count does not fit uint32")
        host = out.cpu().numpy().view(np.uint32)
        return host[: d.n_bases + 1], host[d.n_bases + 1:]

    def batch_text_chunks(self, d: DeviceInput, k: int, rc: bool = False, chunk: int = 8_000_000):
        """Text of the reference's batch files (batch.py:281-296: every k-mer as ">ref:start-end:strand\nSEQ\n",
        stably sorted by sequence, batch.py:156-168), formatted on the GPU and yielded as bytes in chunks of
        `chunk` records.  How the records are split over files is not a contract of the reference
        (SURVEY.md f-3); concatenating the chunks gives one sorted batch."""
        vb = 4 if ((d.pos_offset + d.n_bases) << 1) < (1 << 32) else 8
        streams = self.sorted_streams(d, k, rc, vb)  # stable payload sorts: equal sequences keep emission order
        if len(streams) == 2 and streams[1].n:
            texts = [self.format_uniq(s, d) for s in streams]
            yield self._interleave(texts, [s.keys for s in streams], [s.n for s in streams], k, d.rna, [])
            return
        s = streams[0]
        for lo in range(0, s.n, chunk):
            m = min(chunk, s.n - lo)
            part = KeyArray(s.keys[lo * s.key_bytes:], None, s.vals[lo * s.val_bytes:], None, m, s.key_bytes, s.val_bytes,
                            k, False, is_sorted=True)
            yield self.format_uniq(part, d).cpu().numpy().tobytes()

    def count_text(self, d: DeviceInput, k: int, rc: bool = False) -> bytes:
        """Bytes of the reference's `kmer count` output file (join.py:284)."""
        tabs = self.count(d, k, rc)
        texts = [self.format_counts(t, d.rna) for t in tabs]
        return self._interleave(texts, [t.keys for t in tabs], [t.n for t in tabs], k, d.rna, None)

    def uniq_text(self, d: DeviceInput, k: int, rc: bool = False) -> bytes:
        """Bytes of the reference's `kmer uniq` output file (join.py:262)."""
        sing = self.uniq(d, k, rc)
        texts = [self.format_uniq(s, d) for s in sing]
        return self._interleave(texts, [s.keys for s in sing], [s.n for s in sing], k, d.rna, [])


_engines: Dict[int, Engine] = {}


def get_engine(device: Optional[int] = None) -> Engine:
    """Process-wide engine for a device (created on first use)."""
    if not torch.cuda.is_available():
        _lib.load()
        raise _lib.KmgError("no CUDA device visible: kman_b200 has no CPU fallback")
    idx = torch.cuda.current_device() if device is None else device
    if idx not in _engines:
        _engines[idx] = Engine(idx)
    return _engines[idx]
