"""The C++ host `kmerread` (kmer_read_m3.cpp drop-in) against the compiled reference:
stdout and <wdir>result.txt byte-identical for every input kind, including the 16-probe lookup cap."""
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

OURS = os.path.join(H.ROOT, "kmer_id_b200", "bin", "kmerread")
MITO = os.path.join(H.GOLDEN, "mito")
MITO_NTAXA = 17227


def _mito_parent():
    return H.load_tree(os.path.join(MITO, "mitochondria_tree.txt"), MITO_NTAXA)


def _make_wdir(tmp, db, tree=True):
    w = os.path.join(str(tmp), "w") + "/"
    os.makedirs(w)
    shutil.copy(os.path.join(MITO, "mitochondria_data.txt"), w)
    if tree:
        shutil.copy(os.path.join(MITO, "mitochondria_tree.txt"), w)
    H.write_probes_gz(os.path.join(w, "mitochondria_probes.txt.gz"), db)
    return w


def _run(binary, wdir, f1, f2=None, env=None):
    args = [binary, "-wdir", wdir, "-f1", f1]
    if f2:
        args += ["-f2", f2]
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(args, capture_output=True, timeout=900, env=e)
    res = None
    p = os.path.join(wdir, "result.txt")
    if os.path.exists(p):
        res = open(p, "rb").read()
        os.remove(p)
    return r, res


def _both(ref, wdir, f1, f2=None, env=None):
    r_ref, res_ref = _run(ref, wdir, f1, f2)
    r_gpu, res_gpu = _run(OURS, wdir, f1, f2, env)
    assert r_ref.returncode == r_gpu.returncode == 0, (r_ref.stderr, r_gpu.stderr)
    assert r_gpu.stdout == r_ref.stdout
    assert res_ref is not None and res_gpu == res_ref
    return res_ref


def _write_fasta(path, batch, width=None, gz=False, crlf=False, blanks=False):
    """blanks: sprinkle empty lines.  The gz reader skips them; in the plain reader `linestream >> lseq`
    leaves lseq untouched on an empty line, so the reference appends the previous line AGAIN
    (kmer_read_m3.cpp:945-961) - the drop-in must do the same."""
    eol = b"\r\n" if crlf else b"\n"
    out = []
    for r in range(batch.n):
        if blanks and r % 3 == 1:
            out.append(eol)
        a, b = int(batch.off[r]), int(batch.off[r + 1])
        s = batch.seq[a:b].tobytes()
        out.append(b">" + batch.names[r][1:] + b" extra words" + eol)
        if width:
            out += [s[i:i + width] + eol for i in range(0, len(s), width)]
        else:
            out.append(s + eol)
    data = b"".join(out)
    with (gzip.open(path, "wb") if gz else open(path, "wb")) as f:
        f.write(data)


@pytest.mark.skipif(H.ref_binary("kmerread_small") is None, reason="oracle/_ref/kmerread_small not built")
def test_all_input_kinds_match_reference(tmp_path):
    rng = np.random.default_rng(41)
    parent = _mito_parent()
    pool = np.arange(2, MITO_NTAXA)
    db = H.make_db(rng, 20000, parent=parent, n_dup=300, n_zero=40, taxa_pool=pool)
    w = _make_wdir(tmp_path, db)
    a = H.make_reads(rng, db, 1500, lower_rate=0.01)
    b = H.make_reads(rng, db, 800, ragged=True, name_prefix="T")
    ref = H.ref_binary("kmerread_small")
    d = str(tmp_path)
    H.write_fastq_gz(os.path.join(d, "a.fastq.gz"), a, members=2)
    H.write_fastq_gz(os.path.join(d, "b.fastq.gz"), b, crlf=True)
    with gzip.open(os.path.join(d, "a.fastq.gz"), "rb") as f:
        open(os.path.join(d, "a.fastq"), "wb").write(f.read())
    _write_fasta(os.path.join(d, "a.fasta"), a, width=60)
    _write_fasta(os.path.join(d, "b.fasta"), b, crlf=True, blanks=True)
    _write_fasta(os.path.join(d, "a.fasta.gz"), a, width=70, gz=True, blanks=True)
    res = _both(ref, w, os.path.join(d, "a.fastq.gz"), os.path.join(d, "b.fastq.gz"))  # gz FASTQ: read on the device
    assert _both(ref, w, os.path.join(d, "a.fastq.gz"), os.path.join(d, "b.fastq.gz"), env={"KID_GPU_INGEST": "0"}) == res
    assert len(res.split(b"\n")) == MITO_NTAXA + 1
    g = np.array([int(l.split(b",")[1]) for l in res.split(b"\n")[:-1]])
    assert g[2:].sum() > 500
    _both(ref, w, os.path.join(d, "a.fastq"), "none")
    _both(ref, w, os.path.join(d, "a.fasta"), os.path.join(d, "b.fasta"))
    _both(ref, w, os.path.join(d, "a.fasta.gz"), os.path.join(d, "a.fastq"))
    _both(ref, w, os.path.join(H.GOLDEN, "1a.fasta"))  # the Galaxy fixture (mitokmer.xml:64-70)


@pytest.mark.skipif(H.ref_binary("kmerread_tiny") is None, reason="oracle/_ref/kmerread_tiny not built")
def test_sixteen_probe_cap_is_reproduced(tmp_path):
    """In a 4096-cell reference table 3300 keys probe deep: some sit beyond 16 probes and getHash
    (kmer_read_m3.cpp:232) never finds them.  Reads made of exactly those probes must miss."""
    rng = np.random.default_rng(42)
    parent = _mito_parent()
    db = H.make_db(rng, 3300, parent=parent, n_dup=100, taxa_pool=np.arange(2, 4000))
    w = _make_wdir(tmp_path, db)
    reads = H.make_reads(rng, db, 3000, on_target=1.0, sub_rate=0, n_rate=0, bad_tail=0)
    f1 = os.path.join(str(tmp_path), "r.fastq.gz")
    H.write_fastq_gz(f1, reads)
    res_cap = _both(H.ref_binary("kmerread_tiny"), w, f1, env={"KID_REF_LOG2_CELLS": "12"})
    r, res_nocap = _run(OURS, w, f1, env={"KID_REF_LOG2_CELLS": "30"})
    assert res_nocap != res_cap, "fixture too easy: the cap never mattered"
