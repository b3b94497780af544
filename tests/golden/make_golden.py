"""Regenerates tests/golden/ref_case1 by running the UNMODIFIED reference (oracle/_ref/nk10, built from
/root/reference/newkmer_10nx.cpp by oracle/Makefile) on a small seeded input.  Run it in the build
container (it needs oracle/_ref/nk10 and ~25 GB of RAM for the reference's table):

    python tests/golden/make_golden.py

Committed outputs: the inputs (bact10/probes10.txt.gz, two FASTQ pairs with parser quirks) and what
the reference wrote for them (stdout.txt, *_result.txt, *_reads.txt).  The taxonomy files are the
shipped ones in tests/golden/b10."""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers as H  # noqa: E402

OUT = os.path.join(HERE, "ref_case1")


def main():
    nk10 = H.ref_binary("nk10")
    assert nk10, "build oracle/_ref first (make -C oracle ref)"
    rng = np.random.default_rng(20260101)
    db = H.make_db(rng, 2500, n_dup=120, n_zero=25)
    extra = [
        b"acgtacgtacgtacgtacgtacgtacgtac,7,0,0,F,1\n",
        b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT,9,1,2,F,1\n",
        b"GGGGGGGGGGGGGGGGGGGGGGGGGGGGGG,12,1,2,F\n",
        b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCCA 14 1 2 F 1\r\n",
        b"\n",
        b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCCG,15,1,2,F1\n",
        b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAC,18,1,2,F,1",
    ]
    work = tempfile.mkdtemp(prefix="kid_golden_")
    H.make_bact10_dir(work, db, extra_lines=extra)
    fq = os.path.join(work, "fq")
    os.makedirs(fq)
    a = H.make_reads(rng, db, 400, lower_rate=0.01)
    b = H.make_reads(rng, db, 300, ragged=True, name_prefix="T")
    H.write_fastq_gz(os.path.join(fq, "g1_R1_tr.fastq.gz"), a, members=2)
    H.write_fastq_gz(os.path.join(fq, "g1_R2_tr.fastq.gz"), b, crlf=True)
    H.write_fastq_gz(os.path.join(fq, "g2_R1_tr.fastq.gz"), b, final_newline=False)
    H.write_fastq_gz(os.path.join(fq, "g2_R2_tr.fastq.gz"), a)
    r = H.run_nk10(nk10, work, fq, timeout=1800)
    assert r.returncode == 0, r.stderr
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(os.path.join(OUT, "fq"))
    shutil.copy(os.path.join(work, "bact10", "probes10.txt.gz"), OUT)
    for f in os.listdir(fq):
        shutil.copy(os.path.join(fq, f), os.path.join(OUT, "fq", f))
    # the reference prints the directory it was given: keep the stdout with that line normalised
    out = r.stdout.decode().replace(fq + "/", "<DIR>/")
    open(os.path.join(OUT, "stdout.txt"), "w").write(out)
    shutil.rmtree(work)
    print("wrote", OUT, sorted(os.listdir(os.path.join(OUT, "fq"))))


if __name__ == "__main__":
    main()
