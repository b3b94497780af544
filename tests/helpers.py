"""Shared test helpers: shipped taxonomy fixtures, small synthetic databases and reads, writers
for the reference's on-disk formats, and runners for the compiled reference binaries."""
from __future__ import annotations

import gzip
import os
import shutil
import subprocess
import sys
from dataclasses import dataclass

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
B10_NTAXA = 5982  # MAXTAR, newkmer_10nx.cpp:45
MASK60 = (1 << 60) - 1
_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def ref_binary(name: str = "nk10_small"):
    p = os.path.join(REF_DIR, name)
    return p if os.path.exists(p) else None


# ------------------------------------------------------------------------------- taxonomy
def load_tree(path: str, n_taxa: int) -> np.ndarray:
    """parent[] as Tree1 holds it after main():973-983 (well-formed files only)."""
    parent = np.ones(n_taxa, dtype=np.int32)
    with open(path, "rb") as f:
        for line in f:
            t = line.split()
            if len(t) >= 2:
                parent[int(t[1])] = int(t[0])
    return parent


def load_refkey_counts(path: str, col: int = 2) -> np.ndarray:
    counts = []
    with open(path, "rb") as f:
        next(f)
        for line in f:
            t = line.rstrip(b"\r\n").split(b"\t")
            if len(t) > col:
                counts.append(int(t[col]))
    return np.array(counts, dtype=np.int64)


def b10_parent() -> np.ndarray:
    return load_tree(os.path.join(GOLDEN, "b10", "btree_10.txt"), B10_NTAXA)


def b10_counts() -> np.ndarray:
    c = load_refkey_counts(os.path.join(GOLDEN, "b10", "refkey10.txt"))
    assert c.size == B10_NTAXA
    return c


def ancestors(parent: np.ndarray, t: int):
    out = []
    while t != 1 and t > 0:
        out.append(t)
        t = int(parent[t])
    return out


# ------------------------------------------------------------------------------- k-mers
def revcomp_key(k: np.ndarray) -> np.ndarray:
    k = np.asarray(k, dtype=np.uint64)
    out = np.zeros_like(k)
    for i in range(30):
        c = (k >> np.uint64(2 * i)) & np.uint64(3)
        out |= (np.uint64(3) - c) << np.uint64(2 * (29 - i))
    return out


def canonical(k: np.ndarray) -> np.ndarray:
    k = np.asarray(k, dtype=np.uint64)
    return np.minimum(k, revcomp_key(k))


def key_to_bases(k: int) -> np.ndarray:
    codes = [(k >> (2 * (29 - i))) & 3 for i in range(30)]
    return _BASES[np.array(codes)]


def bases_to_key(s: bytes) -> int:
    k = 0
    for ch in s:
        k = (k << 2) | b"ACGT".index(ch)
    return k


def revcomp_bases(b: np.ndarray) -> np.ndarray:
    lut = np.zeros(256, dtype=np.uint8)
    for a, c in zip(b"ACGTNacgtn", b"TGCANtgcan"):
        lut[a] = c
    return lut[b[::-1]]


@dataclass
class SynthDB:
    keys: np.ndarray    # uint64, file order (forward encodings as the loader sees them)
    taxa: np.ndarray    # uint32
    parent: np.ndarray  # int32[n_taxa]

    @property
    def n_taxa(self) -> int:
        return self.parent.size


def make_db(rng: np.random.Generator, n_probes: int, parent: np.ndarray | None = None,
            n_dup: int = 0, n_zero: int = 0, taxa_pool: np.ndarray | None = None) -> SynthDB:
    """Random canonical 30-mers tagged with taxa drawn from the shipped b10 refkey (leaves and
    internal nodes), plus optional duplicate keys with a different taxon (first must win) and
    taxon-0 lines (must be invisible)."""
    if parent is None:
        parent = b10_parent()
    if taxa_pool is None:
        counts = b10_counts() if parent.size == B10_NTAXA else np.ones(parent.size)
        taxa_pool = np.nonzero(counts > 0)[0]
        taxa_pool = taxa_pool[taxa_pool > 1]
    keys = canonical(rng.integers(0, 1 << 60, size=n_probes, dtype=np.uint64))
    taxa = rng.choice(taxa_pool, size=n_probes).astype(np.uint32)
    if n_dup:
        src = rng.integers(0, n_probes, size=n_dup)
        keys = np.concatenate([keys, keys[src]])
        taxa = np.concatenate([taxa, rng.choice(taxa_pool, size=n_dup).astype(np.uint32)])
    if n_zero:
        zk = canonical(rng.integers(0, 1 << 60, size=n_zero, dtype=np.uint64))
        # half of the zero-taxon lines shadow nothing, half precede a real line with the same key
        keys = np.concatenate([zk, keys, zk[: n_zero // 2]])
        taxa = np.concatenate([np.zeros(n_zero, np.uint32), taxa,
                               rng.choice(taxa_pool, size=n_zero // 2).astype(np.uint32)])
    return SynthDB(keys=keys, taxa=taxa, parent=parent.astype(np.int32))


@dataclass
class ReadBatch:
    seq: np.ndarray   # uint8, concatenated
    qual: np.ndarray  # uint8, same offsets
    off: np.ndarray   # uint64[n+1]
    names: list

    @property
    def n(self) -> int:
        return self.off.size - 1

    def padded(self, pad: int = 16):
        z = np.zeros(pad, dtype=np.uint8)
        return np.concatenate([self.seq, z]), np.concatenate([self.qual, z])


def make_reads(rng: np.random.Generator, db: SynthDB, n_reads: int, length=150, on_target=0.7,
               sub_rate=0.005, n_rate=0.001, lower_rate=0.0, bad_tail=0.2, ragged=False,
               name_prefix="S") -> ReadBatch:
    """Reads stitched from whole probes of one lineage (leaf + ancestors) with random spacers, so
    consecutive hits exercise msca; the rest are uniform random.  Mirrors SURVEY.md section 8(d)."""
    by_taxon: dict[int, np.ndarray] = {}
    order = np.argsort(db.taxa, kind="stable")
    st = db.taxa[order]
    bounds = np.flatnonzero(np.diff(np.concatenate([[-1], st.astype(np.int64), [1 << 40]])))
    for a, b in zip(bounds[:-1], bounds[1:]):
        if st[a] > 0:
            by_taxon[int(st[a])] = order[a:b]
    taxa_with = np.array(sorted(by_taxon))
    seqs, quals, names = [], [], []
    for r in range(n_reads):
        L = int(length if not ragged else rng.choice([1, 5, 29, 30, 31, 32, 33, 45, 60, 61, 62, 63, 64,
                                                      65, 100, 149, 150, 151, 250, 251, 300, 477,
                                                      478, 479, 480, 481, 511, 512, 513, 700, 1100]))
        if rng.random() < on_target and taxa_with.size:
            leaf = int(rng.choice(taxa_with))
            path = [t for t in ancestors(db.parent, leaf) if t in by_taxon] or [leaf]
            # now and then mix in a second lineage so the fold has to find a real LCA
            if rng.random() < 0.15:
                other = int(rng.choice(taxa_with))
                path = path + [t for t in ancestors(db.parent, other) if t in by_taxon][:1]
            parts, tot = [], 0
            skip = int(rng.integers(0, 30))
            while tot < L + 30:
                t = path[int(rng.integers(0, len(path)))]
                k = int(db.keys[int(rng.choice(by_taxon[t]))])
                b = key_to_bases(k)
                if rng.random() < 0.5:
                    b = revcomp_bases(b)
                parts.append(b)
                sp = _BASES[rng.integers(0, 4, size=int(rng.integers(0, 11)))]
                parts.append(sp)
                tot += 30 + sp.size
            s = np.concatenate(parts)[skip:skip + L].copy()
        else:
            s = _BASES[rng.integers(0, 4, size=L)].copy()
        if sub_rate:
            m = rng.random(L) < sub_rate
            s[m] = _BASES[rng.integers(0, 4, size=int(m.sum()))]
        if n_rate:
            s[rng.random(L) < n_rate] = ord("N")
        if lower_rate:
            m = rng.random(L) < lower_rate
            s[m] |= 0x20
        q = np.full(L, ord("I"), dtype=np.uint8)
        if bad_tail and rng.random() < bad_tail and L > 1:
            tl = min(L, int(rng.integers(1, 26)))
            ramp = np.linspace(ord("I"), ord("#"), tl).astype(np.uint8)
            mode = rng.integers(0, 4)
            if mode == 0:
                q[L - tl:] = ramp
            elif mode == 1:
                q[:tl] = ramp[::-1]
            elif mode == 2:
                q[L - tl:] = ramp
                q[:tl] = np.minimum(q[:tl], ramp[::-1])
            else:  # noisy quality everywhere, incl. bytes >= 0x80 (negative as signed char)
                q = rng.choice(np.array([35, 40, 48, 49, 50, 52, 53, 60, 73, 200], dtype=np.uint8), size=L)
        seqs.append(s)
        quals.append(q)
        names.append(f"@{name_prefix}.{r}/1".encode())
    off = np.concatenate([[0], np.cumsum([s.size for s in seqs])]).astype(np.uint64)
    seq_parts, qual_parts = seqs, quals
    return ReadBatch(seq=np.concatenate(seq_parts) if seq_parts else np.zeros(0, np.uint8),
                     qual=np.concatenate(qual_parts) if qual_parts else np.zeros(0, np.uint8),
                     off=off, names=names)


# ------------------------------------------------------------------------------- file writers
def write_probes_gz(path: str, db: SynthDB, extra_lines: list[bytes] | None = None):
    with gzip.open(path, "wb", compresslevel=1) as f:
        for i, (k, t) in enumerate(zip(db.keys.tolist(), db.taxa.tolist())):
            f.write(key_to_bases(k).tobytes() + b",%d,%d,%d,%s,1\n" % (t, i % 14791, i % 1000,
                                                                      b"F" if i & 1 else b"R"))
        for line in extra_lines or []:
            f.write(line)


def write_fastq_gz(path: str, batch: ReadBatch, crlf=False, final_newline=True, members=1):
    eol = b"\r\n" if crlf else b"\n"
    recs = []
    for r in range(batch.n):
        a, b = int(batch.off[r]), int(batch.off[r + 1])
        recs.append(batch.names[r] + eol + batch.seq[a:b].tobytes() + eol + b"+" + eol +
                    batch.qual[a:b].tobytes() + eol)
    data = b"".join(recs)
    if not final_newline and data.endswith(eol):
        data = data[: -len(eol)]
    with open(path, "wb") as f:
        step = max(1, (len(data) + members - 1) // members)
        for i in range(0, max(len(data), 1), step):
            f.write(gzip.compress(data[i:i + step], compresslevel=1))


def make_bact10_dir(workdir: str, db: SynthDB, extra_lines=None, tree_path=None):
    """cwd layout the reference expects: ./bact10/{bData10.txt,btree_10.txt,probes10.txt.gz}"""
    d = os.path.join(workdir, "bact10")
    os.makedirs(d, exist_ok=True)
    shutil.copy(os.path.join(GOLDEN, "b10", "bData10.txt"), d)
    shutil.copy(tree_path or os.path.join(GOLDEN, "b10", "btree_10.txt"), os.path.join(d, "btree_10.txt"))
    write_probes_gz(os.path.join(d, "probes10.txt.gz"), db, extra_lines)
    return d


def run_nk10(binary: str, workdir: str, reads_dir: str, timeout=600):
    """Run an nk10-compatible executable the way the README says: cwd holds ./bact10/, argv[1] is
    the FASTQ directory with a trailing slash."""
    if not reads_dir.endswith("/"):
        reads_dir += "/"
    return subprocess.run([binary, reads_dir], cwd=workdir, capture_output=True, timeout=timeout)


def read_result(path: str):
    a = np.loadtxt(path, delimiter=",", dtype=np.int64)
    return a[:, 1].astype(np.int32), a[:, 2].astype(np.int32)
