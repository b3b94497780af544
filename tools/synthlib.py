"""ctypes binding of tools/libkidsynth.so - the synthetic workload of tools/synth/kid_synth.h.
BENCH / TEST TOOLING, not part of the product."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libkidsynth.so")
lib = C.CDLL(LIB_PATH)


class KsConfig(C.Structure):
    _fields_ = [("seed_db", C.c_uint64), ("seed_reads", C.c_uint64), ("n_probes", C.c_uint64),
                ("n_taxa", C.c_uint32), ("read_len", C.c_uint32), ("stride", C.c_uint32),
                ("on_target_pct", C.c_uint32), ("sub_per_10k", C.c_uint32), ("n_per_10k", C.c_uint32),
                ("bad_tail_pct", C.c_uint32), ("reserved", C.c_uint32)]


_vp, _u64, _i = C.c_void_p, C.c_uint64, C.c_int
lib.ks_create.restype = _vp
lib.ks_create.argtypes = [C.POINTER(KsConfig), _vp, _vp]
lib.ks_free.argtypes = [_vp]
lib.ks_db_host.argtypes = [_vp, _u64, _u64, _vp, _vp]
lib.ks_db_device.argtypes = [_vp, _i, _u64, _u64, _vp, _vp, _vp]
lib.ks_reads_host.argtypes = [_vp, _u64, _u64, _vp, _vp]
lib.ks_reads_device.argtypes = [_vp, _i, _u64, _u64, _vp, _vp, _vp]


def load_taxonomy(golden_dir: str, num: int = 1, den: int = 1, tree="btree_10.txt", refkey="refkey10.txt"):
    """(parent int32[n_taxa], prefix uint64[n_taxa+1]) exactly as tools/kid_synth computes them."""
    counts = []
    with open(os.path.join(golden_dir, refkey), "rb") as f:
        next(f)
        for line in f:
            t = line.rstrip(b"\r\n").split(b"\t")
            if len(t) >= 3:
                counts.append(int(t[2]))
    n = len(counts)
    parent = np.ones(n, dtype=np.int32)
    with open(os.path.join(golden_dir, tree), "rb") as f:
        for line in f:
            t = line.split()
            if len(t) >= 2 and 0 <= int(t[1]) < n:
                parent[int(t[1])] = int(t[0])
    c = np.array([(x * num // den) if i > 1 else 0 for i, x in enumerate(counts)], dtype=np.uint64)
    prefix = np.concatenate([[0], np.cumsum(c)]).astype(np.uint64)
    return parent, prefix


class Workload:
    def __init__(self, parent: np.ndarray, prefix: np.ndarray, read_len=150, stride=None, seed_db=10,
                 seed_reads=21, on_target_pct=70, sub_per_10k=50, n_per_10k=10, bad_tail_pct=20):
        self.parent = np.ascontiguousarray(parent, dtype=np.int32)
        self.prefix = np.ascontiguousarray(prefix, dtype=np.uint64)
        self.cfg = KsConfig(seed_db, seed_reads, int(self.prefix[-1]), self.parent.size, read_len,
                            stride or read_len, on_target_pct, sub_per_10k, n_per_10k, bad_tail_pct, 0)
        self.h = lib.ks_create(C.byref(self.cfg), self.prefix.ctypes.data, self.parent.ctypes.data)
        assert self.h, "ks_create failed"

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            lib.ks_free(self.h)
        self.h = None

    @property
    def n_probes(self) -> int:
        return int(self.prefix[-1])

    @property
    def n_taxa(self) -> int:
        return self.parent.size

    @property
    def stride(self) -> int:
        return self.cfg.stride

    def db_host(self, i0: int = 0, n: int | None = None):
        n = self.n_probes - i0 if n is None else n
        keys = np.empty(n, dtype=np.uint64)
        taxa = np.empty(n, dtype=np.uint32)
        assert lib.ks_db_host(self.h, i0, n, keys.ctypes.data, taxa.ctypes.data) == 0
        return keys, taxa

    def db_device(self, device: int, keys, taxa, i0: int = 0, n: int | None = None, stream: int = 0):
        n = self.n_probes - i0 if n is None else n
        rc = lib.ks_db_device(self.h, device, i0, n, keys.data_ptr(), taxa.data_ptr(), stream or None)
        assert rc == 0, f"ks_db_device rc={rc}"

    def reads_host(self, g0: int, n: int, seq=None, qual=None):
        """seq/qual: optional preallocated buffers (numpy or pinned torch uint8) of n*stride bytes"""
        if seq is None:
            seq = np.empty(n * self.stride + 16, dtype=np.uint8)
            qual = np.empty(n * self.stride + 16, dtype=np.uint8)
        sp = seq.ctypes.data if isinstance(seq, np.ndarray) else seq.data_ptr()
        qp = qual.ctypes.data if isinstance(qual, np.ndarray) else qual.data_ptr()
        assert lib.ks_reads_host(self.h, g0, n, sp, qp) == 0
        return seq, qual

    def reads_device(self, device: int, g0: int, n: int, seq, qual, stream: int = 0):
        rc = lib.ks_reads_device(self.h, device, g0, n, seq.data_ptr(), qual.data_ptr(), stream or None)
        assert rc == 0, f"ks_reads_device rc={rc}"

    def offsets(self, n: int) -> np.ndarray:
        assert self.cfg.stride == self.cfg.read_len, "offsets() describes tightly packed reads"
        return (np.arange(n + 1, dtype=np.uint64) * np.uint64(self.stride))
