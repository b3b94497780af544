"""ctypes binding of oracle/liboracle.so (kid_oracle.c) - TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Nothing under kmer_id_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
REF_DIR = os.path.join(_HERE, "_ref")


def build(quiet: bool = True) -> None:
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


if not os.path.exists(LIB_PATH):
    build()

lib = C.CDLL(LIB_PATH)
_vp, _i, _u, _sz, _u64, _cp = C.c_void_p, C.c_int, C.c_uint, C.c_size_t, C.c_uint64, C.c_char_p
for name, res, args in [
    ("kor_db_new", _vp, [_i, _u]),
    ("kor_db_free", None, [_vp]),
    ("kor_db_n_taxa", _i, [_vp]),
    ("kor_db_n_keys", _u64, [_vp]),
    ("kor_db_add_edge", _i, [_vp, _i, _i]),
    ("kor_db_load_tree", _i, [_vp, _cp]),
    ("kor_db_parent", _i, [_vp, _i]),
    ("kor_msca", _i, [_vp, _i, _i]),
    ("kor_db_add_key", None, [_vp, _u64, C.c_uint32]),
    ("kor_db_add_keys", None, [_vp, _vp, _vp, _sz]),
    ("kor_db_add_probe_seq", _i, [_vp, _cp, _sz, C.c_uint32]),
    ("kor_db_add_probe_line", _i, [_vp, _cp, _sz]),
    ("kor_db_load_probes_gz", C.c_longlong, [_vp, _cp]),
    ("kor_db_lookup", C.c_uint32, [_vp, _u64]),
    ("kor_trim", None, [_vp, _i, C.POINTER(_i), C.POINTER(_i)]),
    ("kor_canonical_key", _i, [_cp, _u, C.POINTER(_u64)]),
    ("kor_sample_new", _vp, [_vp]),
    ("kor_sample_free", None, [_vp]),
    ("kor_sample_reset", None, [_vp]),
    ("kor_sample_gcount", C.POINTER(C.c_int32), [_vp]),
    ("kor_sample_ucount", C.POINTER(C.c_int32), [_vp]),
    ("kor_sample_lookups", _u64, [_vp]),
    ("kor_sample_hits", _u64, [_vp]),
    ("kor_sample_tct", C.c_longlong, [_vp]),
    ("kor_classify_batch", None, [_vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    ("kor_run_fastq_gz", _i, [_vp, _cp, _vp]),
    ("kor_write_result", _i, [_vp, _cp]),
]:
    f = getattr(lib, name)
    f.restype = res
    f.argtypes = args

_libc = C.CDLL(None)
_libc.fopen.restype = _vp
_libc.fopen.argtypes = [_cp, _cp]
_libc.fclose.argtypes = [_vp]

FLAG_ACCEPT_U = 1


class OracleDB:
    def __init__(self, n_taxa: int, flags: int = 0):
        self.h = lib.kor_db_new(n_taxa, flags)
        self.n_taxa = n_taxa

    def __del__(self):
        if getattr(self, "h", None):
            lib.kor_db_free(self.h)
            self.h = None

    def set_parents(self, parent: np.ndarray):
        for child, p in enumerate(np.asarray(parent)):
            assert lib.kor_db_add_edge(self.h, int(p), child) == 0

    def add_edge(self, parent: int, child: int) -> int:
        return lib.kor_db_add_edge(self.h, parent, child)

    def load_tree(self, path: str) -> int:
        return lib.kor_db_load_tree(self.h, path.encode())

    def parents(self) -> np.ndarray:
        """raw Tree1::parent[] is not exported; get_parent() is what the hot path sees"""
        return np.array([lib.kor_db_parent(self.h, i) for i in range(self.n_taxa)], dtype=np.int32)

    def add_keys(self, keys: np.ndarray, taxa: np.ndarray):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        taxa = np.ascontiguousarray(taxa, dtype=np.uint32)
        assert keys.size == taxa.size
        lib.kor_db_add_keys(self.h, keys.ctypes.data, taxa.ctypes.data, keys.size)

    def add_probe_line(self, line: bytes) -> int:
        return lib.kor_db_add_probe_line(self.h, line, len(line))

    def load_probes_gz(self, path: str) -> int:
        return lib.kor_db_load_probes_gz(self.h, path.encode())

    def lookup(self, key: int) -> int:
        return lib.kor_db_lookup(self.h, key)

    def msca(self, x: int, y: int) -> int:
        return lib.kor_msca(self.h, x, y)

    @property
    def n_keys(self) -> int:
        return lib.kor_db_n_keys(self.h)


class OracleSample:
    def __init__(self, db: OracleDB):
        self.db = db
        self.h = lib.kor_sample_new(db.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib.kor_sample_free(self.h)
            self.h = None

    def reset(self):
        lib.kor_sample_reset(self.h)

    def classify(self, seq: np.ndarray, qual, off: np.ndarray):
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        n = off.size - 1
        fin = np.zeros(n, dtype=np.int32)
        st = np.zeros(n, dtype=np.int32)
        sp = np.zeros(n, dtype=np.int32)
        q = None
        if qual is not None:
            q = np.ascontiguousarray(qual, dtype=np.uint8)
        lib.kor_classify_batch(self.h, seq.ctypes.data, q.ctypes.data if q is not None else None,
                               off.ctypes.data, n, fin.ctypes.data, st.ctypes.data, sp.ctypes.data)
        return fin, np.stack([st, sp], axis=1)

    def run_fastq_gz(self, path: str, reads_path: str | None = None, append: bool = False) -> int:
        fp = None
        if reads_path:
            fp = _libc.fopen(reads_path.encode(), b"ab" if append else b"wb")
        rc = lib.kor_run_fastq_gz(self.h, path.encode(), fp)
        if fp:
            _libc.fclose(fp)
        return rc

    def write_result(self, path: str) -> int:
        return lib.kor_write_result(self.h, path.encode())

    @property
    def gcount(self) -> np.ndarray:
        return np.ctypeslib.as_array(lib.kor_sample_gcount(self.h), shape=(self.db.n_taxa,)).copy()

    @property
    def ucount(self) -> np.ndarray:
        return np.ctypeslib.as_array(lib.kor_sample_ucount(self.h), shape=(self.db.n_taxa,)).copy()

    @property
    def lookups(self) -> int:
        return lib.kor_sample_lookups(self.h)

    @property
    def hits(self) -> int:
        return lib.kor_sample_hits(self.h)

    @property
    def tct(self) -> int:
        return lib.kor_sample_tct(self.h)


def trim(qual: bytes, seqlen: int):
    a, b = _i(), _i()
    buf = C.create_string_buffer(qual, len(qual))
    lib.kor_trim(C.cast(buf, _vp), seqlen, C.byref(a), C.byref(b))
    return a.value, b.value


def canonical_key(seq30: bytes, flags: int = 0):
    k = _u64()
    ok = lib.kor_canonical_key(seq30, flags, C.byref(k))
    return k.value if ok else None
