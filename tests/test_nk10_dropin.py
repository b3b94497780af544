"""The C++ host `nk10` (kmer_id_b200/bin/nk10) against the compiled reference on the same files:
stdout lines, <sample>_result.txt and <sample>_reads.txt must be byte-identical.  A byte-identical
_result.txt implies an identical readbatch_10.py CSV (it is a pure function of that file and
refkey10.txt, readbatch_10.py:23-138)."""
import os
import subprocess

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

NK_GPU = os.path.join(H.ROOT, "kmer_id_b200", "bin", "nk10")
NK_REF = H.ref_binary("nk10_small")
SYNTH = os.path.join(H.ROOT, "tools", "kid_synth")


def _run_both(work, fq):
    r_ref = H.run_nk10(NK_REF, work, fq)
    assert r_ref.returncode == 0, r_ref.stderr
    ref_files = {}
    for f in sorted(os.listdir(fq)):
        if f.endswith(("_result.txt", "_reads.txt")):
            p = os.path.join(fq, f)
            ref_files[f] = open(p, "rb").read()
            os.remove(p)
    # default = R1 and R2 classified concurrently on two samples; KID_SERIAL=1 = one after the other;
    # KID_GZ_*: force the multi-threaded inflater (host/pgz.cpp) onto these small files / switch it off
    envs = [{}, {"KID_SERIAL": "1"}, {"KID_GZ_MIN_BYTES": "0", "KID_GZ_PIECE_BYTES": "8192", "KID_NO_CACHE": "1"},
            {"KID_GZ_THREADS": "1"}, {"KID_PARSE_THREADS": "0"}, {"KID_GPUS": "1"},
            # the host reader for every file (by default ordinary gz FASTQ files are inflated and framed on the GPU),
            # and the device reader with small pieces so that these small files are cut into many
            {"KID_GPU_INGEST": "0"}, {"KID_GPU_INGEST": "0", "KID_SERIAL": "1"}, {"KID_GZ_GPU_PIECE": "4096"}]
    import kmer_id_b200 as kid
    if kid.device_count() >= 2:
        # several GPUs in the C++ host (the default {} already uses every visible GPU): each sample's
        # batches dealt to the GPUs with the NCCL / host-sum exchange at sample end, and whole samples
        # dealt to the GPUs
        envs += [{"KID_GPUS": "2"}, {"KID_GPUS": "2", "KID_NO_NCCL": "1"}, {"KID_GPUS": "2", "KID_SERIAL": "1"},
                 {"KID_MULTI_MODE": "samples"}, {"KID_GPUS": "2", "KID_MULTI_MODE": "samples", "KID_SERIAL": "1"},
                 {"KID_MULTI_MODE": "reads"}, {"KID_GPUS": "2", "KID_MULTI_MODE": "reads", "KID_GPU_INGEST": "0"},
                 {"KID_GPUS": "2", "KID_MULTI_MODE": "reads", "KID_NO_NCCL": "1", "KID_GPU_INGEST": "0"}]
    for env in envs:
        r_gpu = subprocess.run([NK_GPU, fq if fq.endswith("/") else fq + "/"], cwd=work, capture_output=True,
                               timeout=600, env=dict(os.environ, **env))
        assert r_gpu.returncode == 0, r_gpu.stderr.decode()
        assert r_gpu.stdout == r_ref.stdout
        assert ref_files
        for f, want in ref_files.items():
            p = os.path.join(fq, f)
            got = open(p, "rb").read()
            os.remove(p)
            assert got == want, f"{f} differs from the reference's ({env})"
    return ref_files


@pytest.mark.skipif(NK_REF is None, reason="oracle/_ref/nk10_small not built")
def test_config0_synthetic_100k_pairs(tmp_path):
    """BASELINE.json configs[0]: 1/100-scale synthetic bact10 DB + 100 k synthetic 150-bp pairs."""
    work = str(tmp_path)
    fq = os.path.join(work, "fq")
    g = os.path.join(H.GOLDEN, "b10")
    subprocess.run([SYNTH, "db", "--golden", g, "--out", work, "--den", "100"], check=True)
    subprocess.run([SYNTH, "reads", "--golden", g, "--out", fq, "--sample", "cfg0", "--pairs", "100000",
                    "--den", "100"], check=True)
    files = _run_both(work, fq)
    res = files["cfg0_result.txt"].decode().split("\n")
    assert len(res) == H.B10_NTAXA + 1
    open(os.path.join(fq, "cfg0_result.txt"), "wb").write(files["cfg0_result.txt"])
    g_, u_ = H.read_result(os.path.join(fq, "cfg0_result.txt"))
    assert g_.sum() == 200000 and g_[0] < 80000 and u_.sum() > 100000


@pytest.mark.skipif(NK_REF is None, reason="oracle/_ref/nk10_small not built")
def test_parser_quirks_two_samples(tmp_path):
    rng = np.random.default_rng(3003)
    db = H.make_db(rng, 5000, n_dup=200, n_zero=40)
    work = str(tmp_path)
    fq = os.path.join(work, "fq")
    os.makedirs(fq)
    extra = [
        b"acgtacgtacgtacgtacgtacgtacgtac,7,0,0,F,1\n",
        b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT,9,1,2,F,1\n",
        b"ACGTACGTACGTNCGTACGTACGTACGTACGTACGTACGTAAAAAAAAAAAAAAAAAAAAAAAAAA,11,1,2,R,1\n",
        b"GGGGGGGGGGGGGGGGGGGGGGGGGGGGGG,12,1,2,F\n",
        b"TTTTTTTTTTTTTTTTTTTTTTTTTTTTTT,13,1,x,F,1\n",
        b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCCA 14 1 2 F 1\r\n",
        b"\n",
        b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCCG,15,1,2,F1\n",
        b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCCT,16,1,2,F,1,extra,fields\n",
        b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCAA,00017,+1,-2,F,1\n",
        b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAC,18,1,2,F,1",
    ]
    H.make_bact10_dir(work, db, extra_lines=extra)
    a = H.make_reads(rng, db, 1200, lower_rate=0.01)
    b = H.make_reads(rng, db, 1200, ragged=True, name_prefix="T")
    # reads made of the hand-written probes so that those lines matter
    special = [b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT", b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCCANNCCCCCCCCCCCCCCCCCCCCCCCCCCCCCG",
               b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCCTACCCCCCCCCCCCCCCCCCCCCCCCCCCCCCAA",
               b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAACGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGG"]
    seq = np.frombuffer(b"".join(special), np.uint8)
    off = np.concatenate([[0], np.cumsum([len(s) for s in special])]).astype(np.uint64)
    c = H.ReadBatch(seq=seq, qual=np.full(seq.size, ord("I"), np.uint8), off=off,
                    names=[b"@sp%d" % i for i in range(len(special))])
    H.write_fastq_gz(os.path.join(fq, "sampA_R1_tr.fastq.gz"), a, members=3)
    H.write_fastq_gz(os.path.join(fq, "sampA_R2_tr.fastq.gz"), b, crlf=True)
    H.write_fastq_gz(os.path.join(fq, "sampB_R1_tr.fastq.gz"), b, final_newline=False)
    H.write_fastq_gz(os.path.join(fq, "sampB_R2_tr.fastq.gz"), c)
    _run_both(work, fq)


@pytest.mark.skipif(NK_REF is None, reason="oracle/_ref/nk10_small not built")
def test_probe_cache_round_trip(tmp_path):
    """Second run uses probes10.txt.gz.kidcache; touching the text file invalidates it."""
    rng = np.random.default_rng(3004)
    db = H.make_db(rng, 3000, n_dup=50, n_zero=10)
    work = str(tmp_path)
    fq = os.path.join(work, "fq")
    os.makedirs(fq)
    H.make_bact10_dir(work, db)
    a = H.make_reads(rng, db, 400)
    H.write_fastq_gz(os.path.join(fq, "c_R1_tr.fastq.gz"), a)
    H.write_fastq_gz(os.path.join(fq, "c_R2_tr.fastq.gz"), a)
    cache = os.path.join(work, "bact10", "probes10.txt.gz.kidcache")
    r1 = H.run_nk10(NK_GPU, work, fq)
    assert r1.returncode == 0 and os.path.exists(cache)
    res1 = open(os.path.join(fq, "c_result.txt"), "rb").read()
    env = dict(os.environ, KID_STATS="1")
    r2 = subprocess.run([NK_GPU, fq + "/"], cwd=work, capture_output=True, env=env)
    assert r2.returncode == 0 and b"cached db" in r2.stderr and r2.stdout == r1.stdout
    assert open(os.path.join(fq, "c_result.txt"), "rb").read() == res1
    # a different probe file of the same name must not be served from the stale cache
    db2 = H.make_db(rng, 3100)
    H.write_probes_gz(os.path.join(work, "bact10", "probes10.txt.gz"), db2)
    r3 = subprocess.run([NK_GPU, fq + "/"], cwd=work, capture_output=True, env=env)
    assert r3.returncode == 0 and b"parse db" in r3.stderr
    assert b"3100 kmers loaded" in r3.stdout
    ref = H.run_nk10(NK_REF, work, fq)
    assert ref.stdout == r3.stdout


@pytest.mark.skipif(NK_REF is None, reason="oracle/_ref/nk10_small not built")
def test_quality_line_shorter_than_its_read_aborts_like_the_reference(tmp_path):
    """qual.at(stop) throws std::out_of_range in the reference (newkmer_10nx.cpp:729) and nothing catches
    it: SIGABRT.  The drop-in dies the same way (from a parse worker or from the reader thread) instead
    of classifying a record the reference never would."""
    rng = np.random.default_rng(3005)
    db = H.make_db(rng, 2000)
    work = str(tmp_path)
    fq = os.path.join(work, "fq")
    os.makedirs(fq)
    H.make_bact10_dir(work, db)
    a = H.make_reads(rng, db, 300)
    H.write_fastq_gz(os.path.join(fq, "q_R2_tr.fastq.gz"), a)
    import gzip
    recs = []
    for r in range(a.n):
        lo, hi = int(a.off[r]), int(a.off[r + 1])
        q = a.qual[lo:hi].tobytes()
        if r == 200:
            q = q[:-7]
        recs.append(a.names[r] + b"\n" + a.seq[lo:hi].tobytes() + b"\n+\n" + q + b"\n")
    with gzip.open(os.path.join(fq, "q_R1_tr.fastq.gz"), "wb") as f:
        f.write(b"".join(recs))
    r_ref = H.run_nk10(NK_REF, work, fq)
    assert r_ref.returncode == -6 and b"out_of_range" in r_ref.stderr
    for env in ({}, {"KID_PARSE_THREADS": "0"}, {"KID_SERIAL": "1"}):
        r_gpu = subprocess.run([NK_GPU, fq + "/"], cwd=work, capture_output=True, timeout=300, env=dict(os.environ, **env))
        assert r_gpu.returncode == -6 and b"out_of_range" in r_gpu.stderr, (env, r_gpu.returncode, r_gpu.stderr[-300:])
