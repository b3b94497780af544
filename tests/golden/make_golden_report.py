"""Regenerates tests/golden/report by running the reference's UNMODIFIED report scripts
(/root/reference/readbatch_10.py, readbatch_c3.py) on seeded inputs.  The scripts hard-code their
input directory, so they are fed to the interpreter on stdin with only that path literal replaced
(nothing is written outside the repo or a temp directory):

    python tests/golden/make_golden_report.py
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "report")
REF = "/root/reference"


def run_reference_script(name, work, result_dir):
    """-> CSV bytes written by the reference script `name` run with cwd=work on result_dir."""
    src = open(os.path.join(REF, name)).read()
    if name == "readbatch_10.py":
        src = src.replace('dir1 = "/home/mmammel/fastq/"', f'dir1 = "{result_dir}/"')
        csv = "test_b10.csv"
    else:
        src = src.replace('mypath = "W:/Mark_backup/ROAR/Saffron/chloro/"', f'mypath = "{result_dir}/"')
        csv = "saffron_chloro.csv"
    assert result_dir in src
    r = subprocess.run([sys.executable, "-"], input=src.encode(), cwd=work, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    return open(os.path.join(work, csv), "rb").read()


def write_results(path, rng, n_targets, frac_hit, with_ucount=True):
    with open(path, "w") as f:
        for t in range(n_targets):
            hit = t < 2 or rng.random() < frac_hit
            g = int(rng.integers(1, 5000)) if hit else 0
            u = int(rng.integers(1, 4 * g + 2)) if hit else 0
            if t % 97 == 5 and hit:
                u = 1  # fails minuniq
            f.write(f"{t},{g},{u}\n" if with_ucount else f"{t},{g}\n")


def write_c3_refkey(path, rng, n_targets):
    with open(path, "w") as f:
        f.write("target\tname\tprobe count\treads hit\treads tested\ttotal size\tstrains\n")
        f.write("0\tnone\t0\t0\t0\t0\t0\n1\troot\t0\t0\t0\t0\t0\n")
        for t in range(2, n_targets):
            depth = int(rng.integers(3, 9))
            name = "_".join(f"L{int(rng.integers(0, 50))}" for _ in range(depth))
            count = int(rng.integers(0, 400))
            hit = int(rng.integers(0, 300))
            tested = hit + int(rng.integers(0, 1000))
            strains = int(rng.integers(0, 4))
            f.write(f"{t}\t{name}\t{count}\t{hit}\t{tested}\t{int(rng.integers(10000, 200000))}\t{strains}\n")


def main():
    rng = np.random.default_rng(4242)
    shutil.rmtree(OUT, ignore_errors=True)
    b10_in = os.path.join(OUT, "b10_results")
    c3_in = os.path.join(OUT, "c3_results")
    os.makedirs(b10_in)
    os.makedirs(c3_in)
    n_b10 = sum(1 for _ in open(os.path.join(HERE, "b10", "refkey10.txt"))) - 1
    write_results(os.path.join(b10_in, "zeta_result.txt"), rng, n_b10, 0.02)
    write_results(os.path.join(b10_in, "alpha_result.txt"), rng, n_b10, 0.01)
    write_results(os.path.join(b10_in, "old_result.txt"), rng, n_b10, 0.01, with_ucount=False)
    for s in ("g1", "g2"):  # results written by the unmodified nk10 (make_golden.py)
        shutil.copy(os.path.join(HERE, "ref_case1", "fq", s + "_result.txt"), b10_in)
    n_c3 = 400
    write_c3_refkey(os.path.join(OUT, "refKeyc3_mini.txt"), rng, n_c3)
    write_results(os.path.join(c3_in, "s2_result.txt"), rng, n_c3, 0.3)
    write_results(os.path.join(c3_in, "s1_result.txt"), rng, n_c3, 0.2)
    open(os.path.join(c3_in, "empty_result.txt"), "w").write("0,7,0\n1,0,0\n")

    work = tempfile.mkdtemp(prefix="kid_report_")
    os.makedirs(os.path.join(work, "bact10"))
    shutil.copy(os.path.join(HERE, "b10", "refkey10.txt"), os.path.join(work, "bact10"))
    shutil.copy(os.path.join(OUT, "refKeyc3_mini.txt"), os.path.join(work, "refKeyc3.txt"))
    open(os.path.join(OUT, "b10.csv"), "wb").write(run_reference_script("readbatch_10.py", work, b10_in))
    open(os.path.join(OUT, "c3.csv"), "wb").write(run_reference_script("readbatch_c3.py", work, c3_in))
    shutil.rmtree(work)
    for f in ("b10.csv", "c3.csv"):
        print(f, sum(1 for _ in open(os.path.join(OUT, f))), "lines")


if __name__ == "__main__":
    main()
