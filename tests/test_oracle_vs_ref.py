"""Pins the CPU oracle (oracle/kid_oracle.c) against the COMPILED reference (oracle/_ref/nk10*).

The reference ships no golden vectors, so the pin is differential: same files in,
byte-identical _result.txt / _reads.txt out.  Needs no GPU.  The binaries are built from
/root/reference by oracle/Makefile and travel with the repo snapshot."""
import os

import numpy as np
import pytest

import helpers as H
from oracle import kor

NK_SMALL = H.ref_binary("nk10_small")
needs_ref = pytest.mark.skipif(NK_SMALL is None, reason="oracle/_ref/nk10_small not built")


def oracle_run_dir(workdir, reads_dir, samples):
    """What main() does (newkmer_10nx.cpp:915-1054), through the oracle's file-level entry points."""
    db = kor.OracleDB(H.B10_NTAXA)
    db.load_tree(os.path.join(workdir, "bact10", "btree_10.txt"))
    n_lines = db.load_probes_gz(os.path.join(workdir, "bact10", "probes10.txt.gz"))
    s = kor.OracleSample(db)
    out = {}
    for name in samples:
        s.reset()
        reads_txt = os.path.join(reads_dir, name + "_oracle_reads.txt")
        rc1 = s.run_fastq_gz(os.path.join(reads_dir, name + "_R1_tr.fastq.gz"), reads_txt)
        rc2 = s.run_fastq_gz(os.path.join(reads_dir, name + "_R2_tr.fastq.gz"), reads_txt, append=True)
        res = os.path.join(reads_dir, name + "_oracle_result.txt")
        s.write_result(res)
        out[name] = (rc1, rc2, res, reads_txt, s.tct)
    return n_lines, out


def _compare(reads_dir, name, res_o, reads_o):
    with open(os.path.join(reads_dir, name + "_result.txt"), "rb") as f:
        ref_res = f.read()
    with open(res_o, "rb") as f:
        assert f.read() == ref_res, "oracle _result.txt differs from the reference's"
    with open(os.path.join(reads_dir, name + "_reads.txt"), "rb") as f:
        ref_reads = f.read()
    with open(reads_o, "rb") as f:
        assert f.read() == ref_reads, "oracle _reads.txt differs from the reference's"


@needs_ref
def test_oracle_matches_reference_end_to_end(tmp_path):
    rng = np.random.default_rng(1001)
    db = H.make_db(rng, 6000, n_dup=200, n_zero=60)
    work = str(tmp_path)
    reads_dir = os.path.join(work, "fq")
    os.makedirs(reads_dir)
    extra = [
        b"acgtacgtacgtacgtacgtacgtacgtac,7,0,0,F,1\n",            # lower case probe: never inserted
        b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT,9,1,2,F,1\n",  # 40-mer: every forward window
        b"ACGTACGTACGTNCGTACGTACGTACGTACGTACGTACGTAAAAAAAAAAAAAAAAAAAAAAAAAA,11,1,2,R,1\n",
        b"GGGGGGGGGGGGGGGGGGGGGGGGGGGGGG,12,1,2,F\n",              # 5 fields: skipped
        b"TTTTTTTTTTTTTTTTTTTTTTTTTTTTTT,13,1,x,F,1\n",            # bad int: skipped
        b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCCA 14 1 2 F 1\r\n",          # blanks + CRLF: accepted
        b"\n",
        b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCCG,15,1,2,F1\n",             # strand and count glued: accepted
        b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCCT,16,1,2,F,1,extra,fields\n",
        b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAC,-5,1,2,F,1",              # no newline at EOF: dropped
    ]
    H.make_bact10_dir(work, db, extra_lines=extra)
    a = H.make_reads(rng, db, 1500, lower_rate=0.01)
    b = H.make_reads(rng, db, 1500, ragged=True, name_prefix="T")
    # reads that hit the hand-written probe lines
    H.write_fastq_gz(os.path.join(reads_dir, "sampA_R1_tr.fastq.gz"), a, members=3)
    H.write_fastq_gz(os.path.join(reads_dir, "sampA_R2_tr.fastq.gz"), b, crlf=True)
    H.write_fastq_gz(os.path.join(reads_dir, "sampB_R1_tr.fastq.gz"), b, final_newline=False)
    H.write_fastq_gz(os.path.join(reads_dir, "sampB_R2_tr.fastq.gz"), a)
    r = H.run_nk10(NK_SMALL, work, reads_dir)
    assert r.returncode == 0, r.stderr
    out_lines = r.stdout.decode().split("\n")
    n_lines, res = oracle_run_dir(work, reads_dir, ["sampA", "sampB"])
    assert out_lines[0] == "tree loaded"
    assert out_lines[1] == f"{n_lines} kmers loaded"
    for name, (rc1, rc2, res_o, reads_o, tct) in res.items():
        assert rc1 == 0 and rc2 == 0
        _compare(reads_dir, name, res_o, reads_o)
        i = out_lines.index(name)
        assert out_lines[i + 2] == f"{tct} reads loaded"
    g, u = H.read_result(os.path.join(reads_dir, "sampA_result.txt"))
    assert g[2:].sum() > 500 and u.sum() > 500, "fixture too easy: hardly any hits"


@needs_ref
def test_fold_order_known_answers(tmp_path):
    """on the shipped b10 tree hits [35,37,35] -> 35 but [35,35,37] -> 5."""
    db = kor.OracleDB(H.B10_NTAXA)
    db.load_tree(os.path.join(H.GOLDEN, "b10", "btree_10.txt"))
    fold = lambda hits: __import__("functools").reduce(lambda f, t: db.msca(t, f), hits)
    assert fold([35, 37, 35]) == 35
    assert fold([35, 35, 37]) == 5
    assert fold([7, 35]) == 5
    assert fold([6, 7]) == 7
    assert fold([7, 6]) == 7
    assert fold([4, 5981]) == 1
    # the same through the real binary: three probes, one read carrying them in each order
    rng = np.random.default_rng(5)
    keys = H.canonical(rng.integers(0, 1 << 60, size=3, dtype=np.uint64))
    sdb = H.SynthDB(keys=keys, taxa=np.array([35, 37, 35], np.uint32), parent=H.b10_parent())
    work = str(tmp_path)
    H.make_bact10_dir(work, sdb)
    fq = os.path.join(work, "fq")
    os.makedirs(fq)

    def read_of(order):
        s = np.concatenate([np.concatenate([H.key_to_bases(int(keys[i])), np.frombuffer(b"NN", np.uint8)])
                            for i in order])
        return s

    # keys[0]:35, keys[1]:37, keys[2]:35
    seqs = [read_of([0, 1, 2]), read_of([0, 2, 1])]
    off = np.concatenate([[0], np.cumsum([s.size for s in seqs])]).astype(np.uint64)
    seq = np.concatenate(seqs)
    batch = H.ReadBatch(seq=seq, qual=np.full(seq.size, ord("I"), np.uint8), off=off,
                        names=[b"@r0", b"@r1"])
    H.write_fastq_gz(os.path.join(fq, "k_R1_tr.fastq.gz"), batch)
    empty = H.ReadBatch(seq=np.zeros(0, np.uint8), qual=np.zeros(0, np.uint8),
                        off=np.zeros(1, np.uint64), names=[])
    H.write_fastq_gz(os.path.join(fq, "k_R2_tr.fastq.gz"), empty)
    r = H.run_nk10(NK_SMALL, work, fq)
    assert r.returncode == 0, r.stderr
    g, u = H.read_result(os.path.join(fq, "k_result.txt"))
    assert g[35] == 1 and g[5] == 1 and g.sum() == 2
    assert u[35] == 2 and u[37] == 1


def test_trim_edge_cases():
    I, B = b"I", b"#"
    assert kor.trim(I * 150, 150) == (0, 149)
    assert kor.trim(B * 150, 150) == (149, 149)
    assert kor.trim(B * 10 + I * 140, 150) == (10, 149)
    assert kor.trim(I * 140 + B * 10, 150) == (0, 139)
    # window rule: '1'(49) passes the single-base cut but four of them sum to 68 -> pass
    assert kor.trim(b"1" * 150, 150) == (0, 149)
    # '0' fails the single-base cut
    assert kor.trim(b"0" * 150, 150)[1] - kor.trim(b"0" * 150, 150)[0] < 30
    # bytes >= 0x80 are negative as signed char
    assert kor.trim(bytes([200]) * 5 + I * 100, 105) == (5, 104)
    assert kor.trim(I, 1) == (0, 0)
