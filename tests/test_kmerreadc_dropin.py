"""The C++ host `kmerreadc` (kmer_read_vf6.cpp drop-in) against the compiled
reference: stdout, <job>_result.txt, <job>_reads.txt and <job>_target_reads.txt byte-identical;
job lists, U/u bases, every input kind, the -target read dump."""
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

import helpers as H
from test_kmerread_dropin import _write_fasta

pytestmark = pytest.mark.gpu

OURS = os.path.join(H.ROOT, "kmer_id_b200", "bin", "kmerreadc")
REF = H.ref_binary("kmerreadc_small")


def _setup(tmp, db, name="fung1"):
    work = str(tmp)
    d = os.path.join(work, name)
    os.makedirs(d)
    # vf6-style database directory: <name>_data.txt / _tree.txt / _probes.txt.gz (config 5's refkey style)
    shutil.copy(os.path.join(H.GOLDEN, "b10", "bData10.txt"), os.path.join(d, name + "_data.txt"))
    shutil.copy(os.path.join(H.GOLDEN, "b10", "btree_10.txt"), os.path.join(d, name + "_tree.txt"))
    H.write_probes_gz(os.path.join(d, name + "_probes.txt.gz"), db)
    return work


def _run(binary, work, args):
    r = subprocess.run([binary] + args, cwd=work, capture_output=True, timeout=900)
    outs = {}
    jd = os.path.join(work, "jobs")
    for f in sorted(os.listdir(jd)):
        if f.endswith(("_result.txt", "_reads.txt")):
            outs[f] = open(os.path.join(jd, f), "rb").read()
            os.remove(os.path.join(jd, f))
    return r, outs


@pytest.mark.skipif(REF is None, reason="oracle/_ref/kmerreadc_small not built")
@pytest.mark.parametrize("target", [0, None])
def test_jobs_match_reference(tmp_path, target):
    rng = np.random.default_rng(51)
    db = H.make_db(rng, 20000, n_dup=300, n_zero=40)
    work = _setup(tmp_path, db)
    jd = os.path.join(work, "jobs")
    os.makedirs(jd)
    a = H.make_reads(rng, db, 1500, lower_rate=0.01)
    b = H.make_reads(rng, db, 700, ragged=True, name_prefix="T")
    # RNA-style reads: U/u instead of T/t must classify exactly like T (kmer_read_vf6.cpp:496-525)
    c = H.make_reads(rng, db, 600, sub_rate=0, name_prefix="U")
    c.seq[c.seq == ord("T")] = ord("U")
    c.seq[::5][c.seq[::5] == ord("U")] = ord("u")
    H.write_fastq_gz(os.path.join(jd, "a.fastq.gz"), a, members=2)
    with gzip.open(os.path.join(jd, "a.fastq.gz"), "rb") as f:
        open(os.path.join(jd, "a.fastq"), "wb").write(f.read())
    H.write_fastq_gz(os.path.join(jd, "c.fastq.gz"), c)
    _write_fasta(os.path.join(jd, "b.fasta"), b, crlf=True)
    _write_fasta(os.path.join(jd, "b.fasta.gz"), b, width=61, gz=True)
    with open(os.path.join(jd, "jobs.txt"), "w") as f:
        f.write("job1 2\r\n./jobs/a.fastq.gz\r\n./jobs/b.fasta\r\n\njob2 3\n./jobs/c.fastq.gz trailing\n"
                "./jobs/b.fasta.gz\n./jobs/a.fastq\n")
    args = ["-name", "fung1", "-jname", "jobs"]
    if target is None:
        # pick a taxon the reference actually assigns so that _target_reads.txt is not empty
        r0, outs0 = _run(REF, work, args)
        g = np.array([int(l.split(b",")[1]) for l in outs0["job1_result.txt"].split(b"\n")[:-1]])
        t = int(np.argmax(g[2:]) + 2)
        args += ["-target", str(t)]
    r_ref, outs_ref = _run(REF, work, args)
    r_gpu, outs_gpu = _run(OURS, work, args)
    assert r_ref.returncode == r_gpu.returncode == 0, (r_ref.stderr, r_gpu.stderr)
    assert r_gpu.stdout == r_ref.stdout
    assert set(outs_ref) == set(outs_gpu) and len(outs_ref) >= 4
    for k in outs_ref:
        assert outs_gpu[k] == outs_ref[k], k
    if target is None:
        assert len(outs_ref["job1_target_reads.txt"]) > 100
        assert outs_ref["job1_reads.txt"] == b""
    else:
        assert len(outs_ref["job2_reads.txt"]) > 1000
    g2 = np.array([int(l.split(b",")[1]) for l in outs_ref["job2_result.txt"].split(b"\n")[:-1]])
    assert g2[2:].sum() > 800, "U/u reads must classify"
