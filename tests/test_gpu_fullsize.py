"""Parity at BASELINE.json's full database size (108 585 519 probes, configs[2]) through
size-independent properties plus an oracle-checked sample:
  * layout M and layout K (two independent table designs) agree read for read on 4 M reads;
  * gcount adds up to the reads kept, ucount is invariant under read order and batch splitting;
  * device-pointer and host-buffer entry points agree;
  * the first 200 k reads match the CPU oracle built over the FULL probe list."""
import os

import numpy as np
import pytest

import helpers as H
from tools import synthlib

pytestmark = pytest.mark.gpu

N_READS = 4_000_000
L = 150


@pytest.fixture(scope="module")
def world():
    import torch
    import kmer_id_b200 as kid
    parent, prefix = synthlib.load_taxonomy(os.path.join(H.GOLDEN, "b10"), 1, 1)
    wl = synthlib.Workload(parent, prefix)
    assert wl.n_probes == 108_585_519
    dev = torch.device("cuda:0")
    dk = torch.empty(wl.n_probes, dtype=torch.int64, device=dev)
    dt = torch.empty(wl.n_probes, dtype=torch.int32, device=dev)
    wl.db_device(0, dk, dt)
    dseq = torch.empty(N_READS * L + 64, dtype=torch.uint8, device=dev)
    dqual = torch.empty(N_READS * L + 64, dtype=torch.uint8, device=dev)
    wl.reads_device(0, 0, N_READS, dseq, dqual)
    doff = torch.arange(N_READS + 1, dtype=torch.int64, device=dev) * L
    torch.cuda.synchronize()
    return dict(kid=kid, torch=torch, wl=wl, parent=parent, dk=dk, dt=dt, dseq=dseq, dqual=dqual, doff=doff, dev=dev)


def _classify_device(w, flags):
    kid, torch = w["kid"], w["torch"]
    db = kid.Database(w["dk"], w["dt"], w["parent"], flags=flags)
    s = kid.Sample(db)
    out = torch.empty(N_READS, dtype=torch.int32, device=w["dev"])
    s.classify_device(w["dseq"], w["dqual"], w["doff"], N_READS, out, None, 0)
    torch.cuda.synchronize()
    g, u = s.counts()
    return db, s, out.cpu().numpy(), g, u, s.counters()


def test_full_db_properties(world):
    w = world
    kid, torch = w["kid"], w["torch"]
    db, s, out_m, g_m, u_m, c_m = _classify_device(w, 0)
    assert db.stats()["n_distinct"] == 108_585_519  # random 60-bit keys: no duplicates expected
    kept = int((out_m >= 0).sum())
    assert int(g_m.sum()) == kept == c_m["reads"]
    assert np.array_equal(np.bincount(out_m[out_m >= 0], minlength=g_m.size), g_m)
    assert 0.6 < (out_m > 1).mean() < 0.8 and u_m.sum() > 1_000_000
    assert (u_m[:2] == 0).all()
    # host entry point, split in odd chunks, reads in a different order of arrival (two halves swapped)
    seq = w["dseq"].cpu().numpy()
    qual = w["dqual"].cpu().numpy()
    half = N_READS // 2
    s.begin()
    s.set_chunk_reads(300_001)
    off_h = (np.arange(half + 1, dtype=np.uint64) * np.uint64(L))
    o2 = s.classify(seq[half * L:], qual[half * L:], off_h)
    o1 = s.classify(np.concatenate([seq[:half * L], np.zeros(16, np.uint8)]),
                    np.concatenate([qual[:half * L], np.zeros(16, np.uint8)]), off_h)
    g2, u2 = s.counts()
    assert np.array_equal(np.concatenate([o1, o2]), out_m)
    assert np.array_equal(g2, g_m) and np.array_equal(u2, u_m)
    del s, db
    # the key-hashed layout is an independent implementation of table + kernel: it must agree
    db_k, s_k, out_k, g_k, u_k, c_k = _classify_device(w, kid.KID_DB_LAYOUT_KEYHASH)
    assert np.array_equal(out_k, out_m)
    assert np.array_equal(g_k, g_m) and np.array_equal(u_k, u_m)
    assert c_k == c_m


def test_sample_against_oracle_with_full_table(world):
    from oracle import kor
    w = world
    kid = w["kid"]
    n = 200_000
    keys = w["dk"].cpu().numpy().view(np.uint64)
    taxa = w["dt"].cpu().numpy().view(np.uint32)
    odb = kor.OracleDB(w["wl"].n_taxa)
    odb.set_parents(w["parent"])
    odb.add_keys(keys, taxa)
    assert odb.n_keys == 108_585_519
    osamp = kor.OracleSample(odb)
    seq = w["dseq"][: n * L + 16].cpu().numpy()
    qual = w["dqual"][: n * L + 16].cpu().numpy()
    off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
    fin_o, span_o = osamp.classify(seq[: n * L], qual[: n * L], off)
    db = kid.Database(w["dk"], w["dt"], w["parent"])
    s = kid.Sample(db)
    fin_g, span_g = s.classify(seq, qual, off, want_span=True)
    g, u = s.counts()
    assert np.array_equal(fin_g, fin_o)
    assert np.array_equal(span_g, span_o.astype(np.uint32))
    assert np.array_equal(g, osamp.gcount) and np.array_equal(u, osamp.ucount)
    c = s.counters()
    assert c["lookups"] == osamp.lookups and c["hits"] == osamp.hits
