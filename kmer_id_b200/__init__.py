"""kmer_id_b200 - B200 (sm_100a) implementation of kmer_id's nk10 read-classification path.

This package is a thin ctypes binding over ``libkmerid_b200.so`` (C-ABI in ``include/kmer_id.h``).
It exists for the tests and ``bench.py``; the product host is the C++ ``nk10`` drop-in under
``kmer_id_b200/host``.  There is no CPU fallback: importing works without a GPU (so the symbol
table can be checked), but every compute call fails with :class:`KidError` when no CUDA device is
usable, and the import itself fails loudly when the shared library has not been built.

Reference seams (``/root/reference/newkmer_10nx.cpp``):
  Database  <- Hashtable (:158-265) + Tree1 (:93-154)
  Sample    <- gcount/ucount/kmer_seen globals (:61-64) and main():1017-1043
  Sample.classify_* <- process_qual (:714-760) + process_read (:452-617)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KID_LIB_PATH") or os.path.join(_HERE, "libkmerid_b200.so")  # (KID_LIB_PATH: A/B builds in experiments)
KSIZE = 30

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make lib` (or __graft_entry__.build()). "
        "kmer_id_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

_vp, _i, _u, _sz, _u64 = C.c_void_p, C.c_int, C.c_uint, C.c_size_t, C.c_uint64
_SIGS = {
    "kid_last_error": (C.c_char_p, []),
    "kid_version": (C.c_char_p, []),
    "kid_device_count": (_i, [C.POINTER(_i)]),
    "kid_device_init": (_i, [_i]),
    "kid_kernel_launches": (C.c_ulonglong, []),
    "kid_host_alloc": (_i, [C.POINTER(_vp), _sz]),
    "kid_host_free": (None, [_vp]),
    "kid_db_build": (_i, [_vp, _vp, _sz, _i, _vp, _i, _i, _u, _i, _vp, C.POINTER(_vp)]),
    "kid_db_free": (None, [_vp]),
    "kid_db_n_taxa": (_i, [_vp]),
    "kid_db_device": (_i, [_vp]),
    "kid_db_stats": (_i, [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)]),
    "kid_db_lookup": (_i, [_vp, _vp, _sz, _vp]),
    "kid_db_msca": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "kid_db_table_device": (_i, [_vp, C.POINTER(_vp), C.POINTER(_u64)]),
    "kid_sample_create": (_i, [_vp, C.POINTER(_vp)]),
    "kid_sample_free": (None, [_vp]),
    "kid_sample_begin": (_i, [_vp, _vp]),
    "kid_sample_db": (_vp, [_vp]),
    "kid_classify_device": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "kid_classify_host": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "kid_pack_bound": (_sz, [_sz, _u64]),
    "kid_pack_reads": (_i, [_vp, _vp, _vp, _sz, _u, C.c_uint32, _vp, _sz, _vp, _vp, C.POINTER(_sz)]),
    "kid_pack_device": (_i, [_vp, _vp, _vp, _vp, _u64, _sz, _vp, _sz, _vp, _vp, _vp]),
    "kid_dense_bound": (_sz, [_u64]),
    "kid_pack_reads_dense": (_i, [_vp, _vp, _vp, _sz, _u, C.c_uint32, _vp, _sz, _vp, _vp, _sz, _vp, _sz,
                                  C.POINTER(_sz), _vp, C.POINTER(C.c_uint32)]),
    "kid_classify_dense_host": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _sz, _vp]),
    "kid_classify_dense_async": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _sz, _sz, _vp]),
    "kid_classify_packed_device": (_i, [_vp, _vp, _vp, _sz, _vp, _vp]),
    "kid_classify_packed_host": (_i, [_vp, _vp, C.c_uint32, _vp, _sz, _vp]),
    "kid_classify_async": (_i, [_vp, _i, _vp, _vp, _vp, _sz, _vp, _vp]),
    "kid_classify_packed_async": (_i, [_vp, _i, _vp, C.c_uint32, _vp, _sz, _vp]),
    "kid_wait": (_i, [_vp, _i]),
    "kid_sample_set_chunk_reads": (_i, [_vp, _sz]),
    "kid_sample_transfer_bytes": (_i, [_vp, C.POINTER(_u64), C.POINTER(_u64)]),
    "kid_sample_counts": (_i, [_vp, _vp, _vp, _vp]),
    "kid_samples_counts": (_i, [C.POINTER(_vp), _i, _vp, _vp, _vp]),
    "kid_sample_counters": (_i, [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64), _vp]),
    "kid_sample_gcount_device": (_i, [_vp, C.POINTER(_vp)]),
    "kid_sample_seen_device": (_i, [_vp, C.POINTER(_vp), C.POINTER(_u64)]),
    "kid_seen_or_device": (_i, [_vp, _vp, C.POINTER(_vp), _i, _u64, _u64, _vp]),
    "kid_ucount_or_range_device": (_i, [_vp, C.POINTER(_vp), _i, _u64, _u64, _vp, _vp]),
    "kid_sample_use_seen_buffer": (_i, [_vp, _vp, _u64]),
    "kid_peer_enable": (_i, [C.POINTER(_i), _i]),
    "kid_sample_ucount_partial": (_i, [_vp, C.POINTER(_vp), _i, _i, _i, _vp]),
    "kid_sample_ucount_device": (_i, [_vp, C.POINTER(_vp)]),
    "kid_sample_read_counts": (_i, [_vp, _vp, _vp, _vp]),
    "kid_device_sync": (_i, [_i]),
    "kid_ucount_range_device": (_i, [_vp, _vp, _u64, _u64, _vp, _vp]),
    "kid_fastq_create": (_i, [_vp, C.POINTER(_vp)]),
    "kid_fastq_free": (None, [_vp]),
    "kid_fastq_load_gz_file": (_i, [_vp, C.c_char_p, C.POINTER(_sz)]),
    "kid_fastq_prefetch_gz_file": (_i, [_vp, C.c_char_p]),
    "kid_fastq_classify": (_i, [_vp, _vp, _vp]),
    "kid_fastq_fetch": (_i, [_vp, _vp, _sz, C.POINTER(_vp), C.POINTER(_vp)]),
    "kid_fastq_stats": (_i, [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64),
                            C.POINTER(C.c_double), _i]),
}
for _name, (_res, _args) in _SIGS.items():
    _f = getattr(lib, _name)  # AttributeError here = the .so does not match include/kmer_id.h
    _f.restype = _res
    _f.argtypes = _args

KID_DB_ACCEPT_U = 1
KID_DB_LAYOUT_KEYHASH = 2
KID_PK_INVALID = 0x80000000
KID_PACK_IMPL_BYTES = 0x100
KID_PACK_IMPL_SWAR = 0x200
KID_MAX_SLOTS = 4
ERROR_NAMES = {-1: "KID_EINVAL", -2: "KID_ECUDA", -3: "KID_ENOMEM", -4: "KID_ERANGE",
               -5: "KID_ETREE", -6: "KID_EFULL", -7: "KID_EUNSUPPORTED"}
KID_EUNSUPPORTED = -7


class KidError(RuntimeError):
    def __init__(self, code: int):
        self.code = code
        msg = lib.kid_last_error().decode("utf-8", "replace")
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {msg}")


def _check(rc: int) -> None:
    if rc != 0:
        raise KidError(rc)


def device_count() -> int:
    n = _i(0)
    rc = lib.kid_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def kernel_launches() -> int:
    return int(lib.kid_kernel_launches())


def _np_ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def _as_ptr(x) -> Optional[int]:
    """numpy array -> host pointer, torch tensor -> data_ptr(), int -> itself, None -> NULL"""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()  # torch.Tensor


def pack_reads(seq: np.ndarray, qual: Optional[np.ndarray], off: np.ndarray, flags: int = 0,
               word0: int = 0, want_span: bool = False, words: Optional[np.ndarray] = None,
               meta: Optional[np.ndarray] = None):
    """kid_pack_reads on numpy arrays (host only, no GPU needed): text batch -> (words, meta[, span]).
    `words` / `meta` may be preallocated (e.g. views of pinned memory); words is returned cut to size."""
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    n = off.size - 1
    if qual is not None:
        qual = np.ascontiguousarray(qual, dtype=np.uint8)
    cap = int(lib.kid_pack_bound(n, int(off[-1] - off[0])))
    if words is None:
        words = np.empty(cap, dtype=np.uint32)
    if meta is None:
        meta = np.empty(2 * (n + 1), dtype=np.uint32)
    assert meta.size >= 2 * (n + 1) and words.dtype == np.uint32 and meta.dtype == np.uint32
    span = np.zeros((n, 2), dtype=np.uint32) if want_span else None
    nw = _sz(0)
    _check(lib.kid_pack_reads(_np_ptr(seq), _np_ptr(qual) if qual is not None else None, _np_ptr(off), n, flags,
                              word0, _np_ptr(words), words.size, _np_ptr(meta),
                              _np_ptr(span) if span is not None else None, C.byref(nw)))
    out = (words[:nw.value], meta[:2 * (n + 1)])
    return out + (span,) if want_span else out


class DenseBatch:
    """A dense read batch in host memory (include/kmer_id.h): filled by append(), handed to
    Sample.classify_dense_host / classify_dense_async.  Arrays may be preallocated (pinned views)."""

    def __init__(self, max_reads: int, max_bases: int, codes=None, boff=None, flagbits=None, inv=None,
                 max_inv: Optional[int] = None, flags: int = 0):
        self.flags = flags
        self.codes = codes if codes is not None else np.zeros(int(lib.kid_dense_bound(max_bases)), np.uint32)
        self.boff = boff if boff is not None else np.zeros(max_reads + 1, np.uint32)
        self.flagbits = flagbits if flagbits is not None else np.zeros(max_reads // 32 + 2, np.uint32)
        self.inv = inv if inv is not None else np.zeros(max_inv if max_inv is not None else max(1024, max_bases // 64), np.uint32)
        self.n = 0
        self.n_bases = 0
        self.n_inv = 0
        self.boff[0] = 0

    def append(self, seq: np.ndarray, qual: Optional[np.ndarray], off: np.ndarray, want_span: bool = False):
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        n = off.size - 1
        if qual is not None:
            qual = np.ascontiguousarray(qual, dtype=np.uint8)
        assert self.n + n + 1 <= self.boff.size
        span = np.zeros((n, 2), dtype=np.uint32) if want_span else None
        ni, nb = _sz(0), C.c_uint32(0)
        w0 = self.n_bases >> 4
        _check(lib.kid_pack_reads_dense(_np_ptr(seq), _np_ptr(qual) if qual is not None else None, _np_ptr(off), n,
                                        self.flags, self.n_bases, self.codes[w0:].ctypes.data, self.codes.size - w0,
                                        self.boff[self.n:].ctypes.data, _np_ptr(self.flagbits), self.n,
                                        self.inv[self.n_inv:].ctypes.data, self.inv.size - self.n_inv, C.byref(ni),
                                        _np_ptr(span) if span is not None else None, C.byref(nb)))
        self.n += n
        self.n_bases = nb.value
        self.n_inv += ni.value
        return span

    @property
    def wire_bytes(self) -> int:
        return 4 * (((self.n_bases + 15) >> 4) + self.n + 1 + (self.n + 31) // 32 + self.n_inv)


class Database:
    """GPU-resident probe table + taxonomy (Hashtable + Tree1 of the reference)."""

    def __init__(self, keys, taxa, parent: np.ndarray, device: int = 0, flags: int = 0,
                 log2_sectors: int = 0, stream: int = 0):
        parent = np.ascontiguousarray(parent, dtype=np.int32)
        on_device = not isinstance(keys, np.ndarray)
        if on_device:
            n = int(keys.numel())
            assert keys.dtype.itemsize == 8 and taxa.dtype.itemsize == 4 and int(taxa.numel()) == n
        else:
            keys = np.ascontiguousarray(keys, dtype=np.uint64)
            taxa = np.ascontiguousarray(taxa, dtype=np.uint32)
            n = keys.size
            assert taxa.size == n
        self._h = _vp()
        _check(lib.kid_db_build(_as_ptr(keys) if n else None, _as_ptr(taxa) if n else None, n,
                                int(on_device), _np_ptr(parent), parent.size, device, flags,
                                log2_sectors, stream or None, C.byref(self._h)))
        self.n_taxa = parent.size
        self.device = device

    def close(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.kid_db_free(self._h)
        self._h = None

    __del__ = close

    def stats(self) -> dict:
        a, b, c, d = _u64(), _u64(), _u64(), _u64()
        _check(lib.kid_db_stats(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"n_distinct": a.value, "n_sectors": b.value, "table_bytes": c.value,
                "n_displaced": d.value}

    def pack_device(self, seq, qual, off, total_bases: int, n_reads: int, words, words_cap: int, meta,
                    out_span=None, stream: int = 0):
        """kid_pack_device: text batch on the device -> packed batch on the device (raw addresses / tensors)."""
        _check(lib.kid_pack_device(self._h, _as_ptr(seq), _as_ptr(qual), _as_ptr(off), total_bases, n_reads,
                                   _as_ptr(words), words_cap, _as_ptr(meta), _as_ptr(out_span), stream or None))

    def table_device(self) -> Tuple[int, int]:
        p, n = _vp(), _u64()
        _check(lib.kid_db_table_device(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def lookup(self, keys: np.ndarray) -> np.ndarray:
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        out = np.zeros(keys.size, dtype=np.uint32)
        _check(lib.kid_db_lookup(self._h, _np_ptr(keys), keys.size, _np_ptr(out)))
        return out

    def msca(self, x: np.ndarray, y: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.int32)
        y = np.ascontiguousarray(y, dtype=np.int32)
        out = np.zeros(x.size, dtype=np.int32)
        _check(lib.kid_db_msca(self._h, _np_ptr(x), _np_ptr(y), x.size, _np_ptr(out)))
        return out


class Sample:
    """Per-sample accumulators (gcount / seen flags -> ucount) and the classify calls."""

    def __init__(self, db: Database):
        self.db = db
        self._h = _vp()
        _check(lib.kid_sample_create(db._h, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.kid_sample_free(self._h)
        self._h = None

    __del__ = close

    def begin(self, stream: int = 0):
        _check(lib.kid_sample_begin(self._h, stream or None))

    def set_chunk_reads(self, n: int):
        _check(lib.kid_sample_set_chunk_reads(self._h, n))

    def classify_device(self, seq, qual, off, n_reads: int, out_taxon=None, out_span=None,
                        stream: int = 0):
        """All arguments are device buffers (torch tensors or raw addresses)."""
        _check(lib.kid_classify_device(self._h, _as_ptr(seq), _as_ptr(qual), _as_ptr(off), n_reads,
                                       _as_ptr(out_taxon), _as_ptr(out_span), stream or None))

    def classify_host(self, seq, qual, off, n_reads: int, out_taxon=None, out_span=None):
        """Host buffers (numpy arrays or pinned torch tensors); returns when outputs are complete."""
        _check(lib.kid_classify_host(self._h, _as_ptr(seq), _as_ptr(qual), _as_ptr(off), n_reads,
                                     _as_ptr(out_taxon), _as_ptr(out_span)))

    def classify_packed_device(self, words, meta, n_reads: int, out_taxon=None, stream: int = 0):
        """Packed batch already on the device (torch tensors or raw addresses)."""
        _check(lib.kid_classify_packed_device(self._h, _as_ptr(words), _as_ptr(meta), n_reads,
                                              _as_ptr(out_taxon), stream or None))

    def classify_packed_host(self, words, meta, n_reads: int, out_taxon=None, word0: int = 0):
        """Packed batch in host memory; returns when out_taxon is complete."""
        _check(lib.kid_classify_packed_host(self._h, _as_ptr(words), word0, _as_ptr(meta), n_reads,
                                            _as_ptr(out_taxon)))

    def classify_dense_host(self, batch: "DenseBatch", out_taxon=None):
        _check(lib.kid_classify_dense_host(self._h, _as_ptr(batch.codes), _as_ptr(batch.boff), _as_ptr(batch.flagbits),
                                           _as_ptr(batch.inv), batch.n_inv, batch.n, _as_ptr(out_taxon)))

    def classify_dense_async(self, slot: int, batch: "DenseBatch", out_taxon=None):
        _check(lib.kid_classify_dense_async(self._h, slot, _as_ptr(batch.codes), _as_ptr(batch.boff),
                                            _as_ptr(batch.flagbits), _as_ptr(batch.inv), batch.n_inv, batch.n,
                                            _as_ptr(out_taxon)))

    def classify_async(self, slot: int, seq, qual, off, n_reads: int, out_taxon=None, out_span=None):
        _check(lib.kid_classify_async(self._h, slot, _as_ptr(seq), _as_ptr(qual), _as_ptr(off), n_reads,
                                      _as_ptr(out_taxon), _as_ptr(out_span)))

    def classify_packed_async(self, slot: int, words, meta, n_reads: int, out_taxon=None, word0: int = 0):
        _check(lib.kid_classify_packed_async(self._h, slot, _as_ptr(words), word0, _as_ptr(meta), n_reads,
                                             _as_ptr(out_taxon)))

    def wait(self, slot: int):
        _check(lib.kid_wait(self._h, slot))

    def classify(self, seq: np.ndarray, qual: Optional[np.ndarray], off: np.ndarray,
                 want_span: bool = False):
        """Convenience for tests: numpy in, numpy out."""
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        n = off.size - 1
        if qual is not None:
            qual = np.ascontiguousarray(qual, dtype=np.uint8)
            assert qual.size >= int(off[-1])
        out = np.full(n, -2, dtype=np.int32)
        span = np.zeros((n, 2), dtype=np.uint32) if want_span else None
        self.classify_host(seq, qual, off, n, out, span)
        return (out, span) if want_span else out

    def counts(self, stream: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        g = np.zeros(self.db.n_taxa, dtype=np.int32)
        u = np.zeros(self.db.n_taxa, dtype=np.int32)
        _check(lib.kid_sample_counts(self._h, _np_ptr(g), _np_ptr(u), stream or None))
        return g, u

    def counters(self, stream: int = 0) -> dict:
        a, b, c = _u64(), _u64(), _u64()
        _check(lib.kid_sample_counters(self._h, C.byref(a), C.byref(b), C.byref(c), stream or None))
        return {"lookups": a.value, "hits": b.value, "reads": c.value}

    def transfer_bytes(self) -> Tuple[int, int]:
        a, b = _u64(), _u64()
        _check(lib.kid_sample_transfer_bytes(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def gcount_device(self) -> int:
        p = _vp()
        _check(lib.kid_sample_gcount_device(self._h, C.byref(p)))
        return p.value

    def seen_device(self) -> Tuple[int, int]:
        p, n = _vp(), _u64()
        _check(lib.kid_sample_seen_device(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def seen_or(self, dst: int, srcs, word0: int, n_words: int, stream: int = 0):
        arr = (_vp * len(srcs))(*srcs)
        _check(lib.kid_seen_or_device(self.db._h, dst, arr, len(srcs), word0, n_words, stream or None))

    def use_seen_buffer(self, ptr: int, n_words: int):
        _check(lib.kid_sample_use_seen_buffer(self._h, ptr, n_words))

    def ucount_or_range(self, srcs, word0: int, n_words: int, ucount_partial, stream: int = 0):
        arr = (_vp * len(srcs))(*srcs)
        _check(lib.kid_ucount_or_range_device(self.db._h, arr, len(srcs), word0, n_words,
                                              _as_ptr(ucount_partial), stream or None))

    def ucount_range(self, seen: int, word0: int, n_words: int, ucount_partial, stream: int = 0):
        _check(lib.kid_ucount_range_device(self.db._h, seen, word0, n_words, _as_ptr(ucount_partial),
                                           stream or None))


class GzFastq:
    """A gzip FASTQ file inflated, framed and classified on the device (kid_fastq_* of include/kmer_id.h;
    process_fqgz, newkmer_10nx.cpp:762-816).  load() returns False for a file the host reader has to take."""

    PHASES = ("read", "find", "inflate", "chain", "resolve", "frame", "classify", "fetch")

    def __init__(self, db: Database):
        self.db = db
        self._h = _vp()
        _check(lib.kid_fastq_create(db._h, C.byref(self._h)))
        self.n_reads = 0
        self.why = ""

    def close(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.kid_fastq_free(self._h)
        self._h = None

    __del__ = close

    def prefetch(self, path: str) -> None:
        _check(lib.kid_fastq_prefetch_gz_file(self._h, os.fsencode(path)))

    def load(self, path: str) -> bool:
        n = _sz(0)
        rc = lib.kid_fastq_load_gz_file(self._h, os.fsencode(path), C.byref(n))
        if rc == KID_EUNSUPPORTED:
            self.why = lib.kid_last_error().decode("utf-8", "replace")
            self.n_reads = 0
            return False
        _check(rc)
        self.n_reads = n.value
        return True

    def classify(self, sample: "Sample") -> np.ndarray:
        out = np.full(self.n_reads, -2, dtype=np.int32)
        _check(lib.kid_fastq_classify(self._h, sample._h, _np_ptr(out) if self.n_reads else None))
        return out

    def fetch(self, reads) -> list:
        """[(header, trimmed bases)] as bytes for the given record indices"""
        idx = np.ascontiguousarray(reads, dtype=np.uint32)
        data, lens = _vp(), _vp()
        _check(lib.kid_fastq_fetch(self._h, _np_ptr(idx) if idx.size else None, idx.size, C.byref(data), C.byref(lens)))
        if idx.size == 0:
            return []
        ln = np.ctypeslib.as_array(C.cast(lens, C.POINTER(C.c_uint32)), shape=(2 * idx.size,)).copy()
        total = int(ln.sum())
        raw = C.string_at(data, total) if total else b""
        out, at = [], 0
        for i in range(idx.size):
            h, b = int(ln[2 * i]), int(ln[2 * i + 1])
            out.append((raw[at:at + h], raw[at + h:at + h + b]))
            at += h + b
        return out

    def stats(self) -> dict:
        a, b, c, d = _u64(), _u64(), _u64(), _u64()
        ph = (C.c_double * 8)()
        _check(lib.kid_fastq_stats(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d), ph, 8))
        return {"text_bytes": a.value, "pieces": b.value, "inflated_again": c.value, "members": d.value,
                "seconds": dict(zip(self.PHASES, [float(x) for x in ph]))}
