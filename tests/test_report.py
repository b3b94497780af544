"""Report step: kmer_id_b200/report.py against the reference's own report scripts.

* golden: CSVs written by the UNMODIFIED readbatch_10.py / readbatch_c3.py (tests/golden/report,
  generator tests/golden/make_golden_report.py) - runs anywhere;
* live: where /root/reference exists (the build container) the scripts are run again beside the
  module on fresh seeded inputs."""
import os
import shutil
import sys

import numpy as np
import pytest

import helpers as H

from kmer_id_b200 import report

sys.path.insert(0, H.GOLDEN)
import make_golden_report as gen  # noqa: E402

REP = os.path.join(H.GOLDEN, "report")


def test_b10_matches_golden_csv():
    text, n_c, n_t = report.build_report(os.path.join(REP, "b10_results"),
                                         os.path.join(H.GOLDEN, "b10", "refkey10.txt"), "b10")
    assert text.encode() == open(os.path.join(REP, "b10.csv"), "rb").read()
    assert n_c == 5 and n_t > 0


def test_c3_matches_golden_csv():
    text, n_c, _ = report.build_report(os.path.join(REP, "c3_results"),
                                       os.path.join(REP, "refKeyc3_mini.txt"), "c3")
    assert text.encode() == open(os.path.join(REP, "c3.csv"), "rb").read()
    assert n_c == 3


def test_cli_writes_same_file(tmp_path):
    out = str(tmp_path / "o.csv")
    report.main(["--dir", os.path.join(REP, "b10_results"), "--refkey",
                 os.path.join(H.GOLDEN, "b10", "refkey10.txt"), "--out", out])
    assert open(out, "rb").read() == open(os.path.join(REP, "b10.csv"), "rb").read()


@pytest.mark.skipif(not os.path.exists(os.path.join(gen.REF, "readbatch_10.py")),
                    reason="reference scripts only exist in the build container")
@pytest.mark.parametrize("seed", [1, 2])
def test_live_against_reference_scripts(tmp_path, seed):
    rng = np.random.default_rng(seed)
    work = str(tmp_path)
    os.makedirs(os.path.join(work, "bact10"))
    refkey = os.path.join(work, "bact10", "refkey10.txt")
    shutil.copy(os.path.join(H.GOLDEN, "b10", "refkey10.txt"), refkey)
    res = os.path.join(work, "res")
    os.makedirs(res)
    n = sum(1 for _ in open(refkey)) - 1
    for i in range(3 + seed):
        gen.write_results(os.path.join(res, f"smp{(7 * i) % 5}_{i}_result.txt"), rng, n, 0.01 * (i + 1),
                          with_ucount=(i != 1))
    want = gen.run_reference_script("readbatch_10.py", work, res)
    assert report.build_report(res, refkey, "b10")[0].encode() == want

    c3key = os.path.join(work, "refKeyc3.txt")
    gen.write_c3_refkey(c3key, rng, 250)
    res3 = os.path.join(work, "res3")
    os.makedirs(res3)
    for i in range(2 + seed):
        gen.write_results(os.path.join(res3, f"c{i}_result.txt"), rng, 250, 0.15 * (i + 1))
    want = gen.run_reference_script("readbatch_c3.py", work, res3)
    assert report.build_report(res3, c3key, "c3")[0].encode() == want
