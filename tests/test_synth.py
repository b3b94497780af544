"""The synthetic workload generators: CPU library == CLI files (here), CPU == GPU (on the box)."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from tools import synthlib

G = os.path.join(H.GOLDEN, "b10")


def test_cpu_generator_matches_cli_files(tmp_path):
    parent, prefix = synthlib.load_taxonomy(G, 1, 1000)
    wl = synthlib.Workload(parent, prefix)
    keys, taxa = wl.db_host()
    assert keys.size == wl.n_probes > 50000
    assert np.array_equal(keys, H.canonical(keys)), "probe keys must be canonical like the builder's"
    work = str(tmp_path)
    synth = os.path.join(H.ROOT, "tools", "kid_synth")
    subprocess.run([synth, "db", "--golden", G, "--out", work, "--den", "1000"], check=True, stdout=subprocess.DEVNULL)
    with gzip.open(os.path.join(work, "bact10", "probes10.txt.gz"), "rb") as f:
        lines = f.read().split(b"\n")[:-1]
    assert len(lines) == keys.size
    for i in (0, 1, 777, keys.size - 1):
        t = lines[i].split(b",")
        assert H.bases_to_key(t[0]) == int(keys[i]) and int(t[1]) == int(taxa[i])
    subprocess.run([synth, "reads", "--golden", G, "--out", work, "--sample", "x", "--pairs", "500", "--den", "1000"],
                   check=True, stdout=subprocess.DEVNULL)
    seq, qual = wl.reads_host(0, 1000)
    for mate, fn in ((0, "x_R1_tr.fastq.gz"), (1, "x_R2_tr.fastq.gz")):
        with gzip.open(os.path.join(work, fn), "rb") as f:
            rec = f.read().split(b"\n")
        for p in (0, 17, 499):
            g = 2 * p + mate
            assert rec[4 * p] == b"@S.%d/%d" % (p, mate + 1)
            assert rec[4 * p + 1] == seq[g * 150:(g + 1) * 150].tobytes()
            assert rec[4 * p + 3] == qual[g * 150:(g + 1) * 150].tobytes()
    # the workload has the advertised shape
    s = seq[:150000].reshape(1000, 150)
    assert 0.0005 < (s == ord("N")).mean() < 0.002
    q = qual[:150000].reshape(1000, 150)
    assert 0.1 < (q[:, -1] < ord("1")).mean() < 0.3


@pytest.mark.gpu
def test_gpu_generator_matches_cpu():
    import torch
    parent, prefix = synthlib.load_taxonomy(G, 1, 100)
    wl = synthlib.Workload(parent, prefix)
    keys, taxa = wl.db_host()
    dev = torch.device("cuda:0")
    dk = torch.empty(wl.n_probes, dtype=torch.int64, device=dev)
    dt = torch.empty(wl.n_probes, dtype=torch.int32, device=dev)
    wl.db_device(0, dk, dt)
    torch.cuda.synchronize()
    assert np.array_equal(dk.cpu().numpy().view(np.uint64), keys)
    assert np.array_equal(dt.cpu().numpy().view(np.uint32), taxa)
    n = 20000
    seq, qual = wl.reads_host(12345, n)
    ds = torch.empty(n * 150 + 16, dtype=torch.uint8, device=dev)
    dq = torch.empty(n * 150 + 16, dtype=torch.uint8, device=dev)
    wl.reads_device(0, 12345, n, ds, dq)
    torch.cuda.synchronize()
    assert np.array_equal(ds.cpu().numpy()[:n * 150], seq[:n * 150])
    assert np.array_equal(dq.cpu().numpy()[:n * 150], qual[:n * 150])
