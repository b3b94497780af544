"""Committed golden vectors from the UNMODIFIED reference (tests/golden/ref_case1, produced by
tests/golden/make_golden.py with oracle/_ref/nk10 = newkmer_10nx.cpp compiled as is).

The CPU test pins the oracle to them; the GPU test runs the shipped drop-in (kmer_id_b200/bin/nk10,
through the C-ABI) on the same inputs.  Neither needs /root/reference or oracle/_ref at run time."""
import os
import shutil
import subprocess

import pytest

import helpers as H
from oracle import kor

CASE = os.path.join(H.GOLDEN, "ref_case1")
SAMPLES = ("g1", "g2")


def _stage(tmp_path):
    work = str(tmp_path)
    os.makedirs(os.path.join(work, "bact10"))
    shutil.copy(os.path.join(CASE, "probes10.txt.gz"), os.path.join(work, "bact10"))
    for f in ("btree_10.txt", "bData10.txt"):
        shutil.copy(os.path.join(H.GOLDEN, "b10", f), os.path.join(work, "bact10"))
    fq = os.path.join(work, "fq")
    os.makedirs(fq)
    for f in os.listdir(os.path.join(CASE, "fq")):
        if f.endswith(".fastq.gz"):
            shutil.copy(os.path.join(CASE, "fq", f), fq)
    return work, fq


def _golden(name):
    with open(os.path.join(CASE, "fq", name), "rb") as f:
        return f.read()


def test_oracle_matches_golden(tmp_path):
    work, fq = _stage(tmp_path)
    db = kor.OracleDB(H.B10_NTAXA)
    db.load_tree(os.path.join(work, "bact10", "btree_10.txt"))
    n_lines = db.load_probes_gz(os.path.join(work, "bact10", "probes10.txt.gz"))
    lines = open(os.path.join(CASE, "stdout.txt")).read().split("\n")
    assert lines[1] == f"{n_lines} kmers loaded"
    s = kor.OracleSample(db)
    printed = []
    for name in SAMPLES:
        s.reset()
        reads_txt = os.path.join(fq, name + "_o_reads.txt")
        s.run_fastq_gz(os.path.join(fq, name + "_R1_tr.fastq.gz"), reads_txt)
        printed.append(f"{s.tct} reads loaded")
        s.run_fastq_gz(os.path.join(fq, name + "_R2_tr.fastq.gz"), reads_txt, append=True)
        printed.append(f"{s.tct} reads loaded")
        res = os.path.join(fq, name + "_o_result.txt")
        s.write_result(res)
        assert open(res, "rb").read() == _golden(name + "_result.txt")
        assert open(reads_txt, "rb").read() == _golden(name + "_reads.txt")
    assert printed == [lines[4], lines[5], lines[7], lines[8]]


@pytest.mark.gpu
def test_dropin_matches_golden(tmp_path):
    ours = os.path.join(H.ROOT, "kmer_id_b200", "bin", "nk10")
    assert os.path.exists(ours), "run `make host` first"
    work, fq = _stage(tmp_path)
    for env in ({}, {"KID_GZ_MIN_BYTES": "0", "KID_GZ_PIECE_BYTES": "4096", "KID_NO_CACHE": "1"}):  # zlib / multi-threaded inflate
        r = subprocess.run([ours, fq + "/"], cwd=work, capture_output=True, timeout=600, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr.decode()
        want = open(os.path.join(CASE, "stdout.txt")).read().replace("<DIR>/", fq + "/")
        assert r.stdout.decode() == want
        for name in SAMPLES:
            for suffix in ("_result.txt", "_reads.txt"):
                p = os.path.join(fq, name + suffix)
                assert open(p, "rb").read() == _golden(name + suffix), (name + suffix, env)
                os.remove(p)
