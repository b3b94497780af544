"""Parity of the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.
Bit-exact: per-read final taxon, trimmed span, gcount, ucount, lookup and hit counters."""
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

LAYOUT_M, LAYOUT_K = 0, 2  # KID_DB_LAYOUT_KEYHASH = 2


@pytest.fixture(params=[LAYOUT_M, LAYOUT_K], ids=["layoutM", "layoutK"])
def layout(request):
    return request.param


def _oracle(db, flags=0):
    from oracle import kor
    odb = kor.OracleDB(db.n_taxa, flags)
    odb.set_parents(db.parent)
    odb.add_keys(db.keys, db.taxa)
    return odb, kor.OracleSample(odb)


def _gpu(db, flags=0, **kw):
    import kmer_id_b200 as kid
    gdb = kid.Database(db.keys, db.taxa, db.parent, flags=flags, **kw)
    return gdb, kid.Sample(gdb)


def _check_batch(gs, osamp, batch, with_qual=True):
    seq, qual = batch.padded()
    fin_o, span_o = osamp.classify(batch.seq, batch.qual if with_qual else None, batch.off)
    fin_g, span_g = gs.classify(seq, qual if with_qual else None, batch.off, want_span=True)
    bad = np.flatnonzero(fin_o != fin_g)
    assert bad.size == 0, f"{bad.size} reads differ, first {bad[:5]}: oracle {fin_o[bad[:5]]} gpu {fin_g[bad[:5]]}"
    if with_qual:
        assert np.array_equal(span_o.astype(np.uint32), span_g)
    return fin_o


def _check_counts(gs, osamp):
    g, u = gs.counts()
    assert np.array_equal(g, osamp.gcount)
    assert np.array_equal(u, osamp.ucount)
    c = gs.counters()
    assert c["lookups"] == osamp.lookups
    assert c["hits"] == osamp.hits
    assert c["reads"] == osamp.tct


def test_lookup_first_wins_and_zero_taxon(layout):
    rng = np.random.default_rng(7)
    db = H.make_db(rng, 50000, n_dup=5000, n_zero=400)
    odb, _ = _oracle(db)
    gdb, _ = _gpu(db, layout)
    probe = np.concatenate([db.keys, H.canonical(rng.integers(0, 1 << 60, size=20000, dtype=np.uint64))])
    want = np.array([odb.lookup(int(k)) for k in probe.tolist()], dtype=np.uint32)
    got = gdb.lookup(probe)
    assert np.array_equal(want, got)
    assert gdb.stats()["n_distinct"] == odb.n_keys


def test_msca_exhaustive_sample():
    rng = np.random.default_rng(8)
    db = H.make_db(rng, 10)
    odb, _ = _oracle(db)
    gdb, _ = _gpu(db)
    x = rng.integers(1, H.B10_NTAXA, size=20000).astype(np.int32)
    y = rng.integers(1, H.B10_NTAXA, size=20000).astype(np.int32)
    # force plenty of related pairs: y = some ancestor of x, or siblings
    par = db.parent
    y[:5000] = par[x[:5000]]
    y[5000:7000] = par[par[x[5000:7000]]]
    x[7000:8000] = par[y[7000:8000]]
    y[8000:8500] = x[8000:8500]
    y[8500:9000] = 1
    x[9000:9500] = 1
    want = np.array([odb.msca(int(a), int(b)) for a, b in zip(x, y)], dtype=np.int32)
    assert np.array_equal(want, gdb.msca(x, y))
    assert gdb.msca(np.array([37, 35], np.int32), np.array([35, 35], np.int32)).tolist() == [5, 35]


@pytest.mark.parametrize("seed,n,kw", [
    (11, 4000, dict()),
    (12, 3000, dict(ragged=True)),
    (13, 3000, dict(lower_rate=0.05, n_rate=0.01)),
    (14, 2000, dict(length=250, bad_tail=0.6)),
    (15, 1500, dict(length=31)),
    (16, 1500, dict(on_target=1.0, sub_rate=0.0, n_rate=0.0)),
])
def test_classify_matches_oracle(seed, n, kw, layout):
    rng = np.random.default_rng(seed)
    db = H.make_db(rng, 30000, n_dup=500, n_zero=50)
    odb, osamp = _oracle(db)
    gdb, gs = _gpu(db, layout)
    batch = H.make_reads(rng, db, n, **kw)
    fin = _check_batch(gs, osamp, batch)
    _check_counts(gs, osamp)
    if kw.get("length", 150) >= 100:
        assert (fin > 1).sum() > n // 10, "fixture has too few classified reads"
    # second batch accumulates into the same sample; a repeat must not bump ucount
    _check_batch(gs, osamp, batch)
    _check_counts(gs, osamp)
    # new sample
    gs.begin()
    osamp.reset()
    b2 = H.make_reads(rng, db, 500, **kw)
    _check_batch(gs, osamp, b2)
    _check_counts(gs, osamp)


def test_classify_without_quality_fasta_rule(layout):
    rng = np.random.default_rng(21)
    db = H.make_db(rng, 20000)
    odb, osamp = _oracle(db)
    gdb, gs = _gpu(db, layout)
    batch = H.make_reads(rng, db, 2500, ragged=True)
    _check_batch(gs, osamp, batch, with_qual=False)
    _check_counts(gs, osamp)


def test_accept_u_flag(layout):
    import kmer_id_b200 as kid
    from oracle import kor
    rng = np.random.default_rng(22)
    db = H.make_db(rng, 5000)
    batch = H.make_reads(rng, db, 1500, sub_rate=0, n_rate=0)
    batch.seq[batch.seq == ord("T")] = ord("U")
    batch.seq[::7][batch.seq[::7] == ord("U")] = ord("u")
    for flags_o, flags_g in ((0, 0), (kor.FLAG_ACCEPT_U, kid.KID_DB_ACCEPT_U)):
        odb, osamp = _oracle(db, flags_o)
        gdb, gs = _gpu(db, flags_g | layout)
        fin = _check_batch(gs, osamp, batch)
        _check_counts(gs, osamp)
        assert ((fin > 1).sum() > 100) == bool(flags_o)


def test_displaced_entries_small_table_high_load(layout):
    """Force bucket overflow: 2^22 buckets hold 16.7 M slots; 6 M keys -> many displaced keys."""
    rng = np.random.default_rng(23)
    db = H.make_db(rng, 6_000_000)
    gdb, gs = _gpu(db, layout, log2_sectors=22)
    st = gdb.stats()
    assert st["n_displaced"] > 10000
    odb, osamp = _oracle(db)
    probe = np.concatenate([db.keys[:200000], H.canonical(rng.integers(0, 1 << 60, size=200000, dtype=np.uint64))])
    from oracle import kor
    want = np.zeros(probe.size, np.uint32)
    for i, k in enumerate(probe.tolist()):
        want[i] = odb.lookup(k)
    assert np.array_equal(want, gdb.lookup(probe))
    batch = H.make_reads(rng, db, 3000)
    _check_batch(gs, osamp, batch)
    _check_counts(gs, osamp)


def test_device_resident_entry_point_and_chunking(layout):
    import torch
    rng = np.random.default_rng(24)
    db = H.make_db(rng, 20000)
    odb, osamp = _oracle(db)
    gdb, gs = _gpu(db, layout)
    batch = H.make_reads(rng, db, 5000, ragged=True)
    seq, qual = batch.padded()
    fin_o, _ = osamp.classify(batch.seq, batch.qual, batch.off)
    dev = torch.device("cuda:0")
    dseq = torch.from_numpy(seq).to(dev)
    dqual = torch.from_numpy(qual).to(dev)
    doff = torch.from_numpy(batch.off.view(np.int64)).to(dev)
    dout = torch.full((batch.n,), -7, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    gs.begin(stream)
    gs.classify_device(dseq, dqual, doff, batch.n, dout, None, stream)
    torch.cuda.synchronize()
    assert np.array_equal(dout.cpu().numpy(), fin_o)
    _check_counts(gs, osamp)
    # host path split into many small chunks must give the same sample totals
    gs.begin()
    gs.set_chunk_reads(333)
    out = gs.classify(seq, qual, batch.off)
    assert np.array_equal(out, fin_o)
    _check_counts(gs, osamp)


def test_rejects_bad_inputs():
    import kmer_id_b200 as kid
    rng = np.random.default_rng(25)
    db = H.make_db(rng, 100)
    bad_taxa = db.taxa.copy()
    bad_taxa[5] = H.B10_NTAXA  # gcount[] out of bounds in the reference
    with pytest.raises(kid.KidError) as e:
        kid.Database(db.keys, bad_taxa, db.parent)
    assert e.value.code == -4
    cyc = db.parent.copy()
    cyc[100], cyc[101] = 101, 100  # Tree1::msca would never terminate
    with pytest.raises(kid.KidError) as e:
        kid.Database(db.keys, db.taxa, cyc)
    assert e.value.code == -5
    # empty database and empty batch are fine
    gdb = kid.Database(np.zeros(0, np.uint64), np.zeros(0, np.uint32), db.parent)
    gs = kid.Sample(gdb)
    out = gs.classify(np.zeros(16, np.uint8), np.zeros(16, np.uint8), np.zeros(1, np.uint64))
    assert out.size == 0
    b = H.make_reads(rng, db, 50)
    seq, qual = b.padded()
    assert (gs.classify(seq, qual, b.off) <= 0).all()
    # context warm-up entry: fine on a real device, KID_EINVAL on one that does not exist
    assert kid.lib.kid_device_init(0) == 0
    assert kid.lib.kid_device_init(kid.device_count()) == -1
    assert b"device" in kid.lib.kid_last_error()


def test_build_is_deterministic(layout):
    """Two builds of the same probe list must give bit-identical tables: in a multi-GPU run the seen
    bitmaps of the replicas are OR-ed by slot index (kmer_id_b200/multi_gpu.py)."""
    import torch
    from kmer_id_b200 import multi_gpu
    rng = np.random.default_rng(31)
    db = H.make_db(rng, 3_000_000, n_dup=100000, n_zero=1000)
    tabs = []
    for _ in range(2):
        gdb, _ = _gpu(db, layout, log2_sectors=22)
        ptr, n_sectors = gdb.table_device()
        nbytes = gdb.stats()["table_bytes"]
        t = multi_gpu.wrap_device(ptr, nbytes // 4, "<i4", torch.device("cuda:0")).clone()
        tabs.append((t, gdb.stats()))
        del gdb
    assert tabs[0][1] == tabs[1][1]
    assert torch.equal(tabs[0][0], tabs[1][0])
    assert tabs[0][1]["n_displaced"] > 1000


def test_msca_deep_tree():
    """kid_msca on a deep random tree (the shipped trees are at most 6 deep) against the oracle."""
    import kmer_id_b200 as kid
    from oracle import kor
    rng = np.random.default_rng(61)
    n = 400
    parent = np.ones(n, np.int32)
    for v in range(2, n):  # random recursive tree with a long spine: depth well beyond 7
        parent[v] = v - 1 if v < 40 else int(rng.integers(1, v))
    odb = kor.OracleDB(n)
    odb.set_parents(parent)
    gdb = kid.Database(np.zeros(0, np.uint64), np.zeros(0, np.uint32), parent)
    x = rng.integers(1, n, size=20000).astype(np.int32)
    y = rng.integers(1, n, size=20000).astype(np.int32)
    want = np.array([odb.msca(int(a), int(b)) for a, b in zip(x, y)], dtype=np.int32)
    assert np.array_equal(want, gdb.msca(x, y))
    # and a read-level check: hits of a deep lineage folded in order
    keys = H.canonical(rng.integers(0, 1 << 60, size=3000, dtype=np.uint64))
    db = H.SynthDB(keys=keys, taxa=rng.integers(2, n, size=3000).astype(np.uint32), parent=parent)
    o2, osamp = _oracle(db)
    g2, gs = _gpu(db)
    batch = H.make_reads(rng, db, 1500)
    _check_batch(gs, osamp, batch)
    _check_counts(gs, osamp)


def test_sample_end_kernels_with_two_shards_on_one_gpu(layout):
    """The multi-GPU sample-end kernels, exercised on ONE device: two samples play two ranks (same
    replica, different shards, 300 reads in common).  Both transports - OR into a slice then
    histogram (kid_seen_or_device + kid_ucount_range_device) and the fused peer kernel
    (kid_ucount_or_range_device) - must give the ucount of the union, which is NOT the sum."""
    import torch
    import kmer_id_b200 as kid
    rng = np.random.default_rng(71)
    db = H.make_db(rng, 20000)
    odb, osamp = _oracle(db)
    gdb, s0 = _gpu(db, layout)
    s1 = kid.Sample(gdb)
    common = H.make_reads(rng, db, 300, name_prefix="C")
    shards = [H.make_reads(rng, db, 1500, name_prefix="A"), H.make_reads(rng, db, 1500, name_prefix="B")]
    per_shard_u = []
    for s, sh in zip((s0, s1), shards):
        for b in (sh, common):
            seq, qual = b.padded()
            s.classify(seq, qual, b.off)
            osamp.classify(b.seq, b.qual, b.off)
        per_shard_u.append(s.counts()[1])
    want_u = osamp.ucount
    assert not np.array_equal(per_shard_u[0] + per_shard_u[1], want_u), "fixture must make the naive sum wrong"
    p0, n_words = s0.seen_device()
    p1, _ = s1.seen_device()
    dev = torch.device("cuda:0")
    half = n_words // 2
    assert half % 4 == 0
    # fused: each "rank" histograms its half of the OR of both bitmaps, partials add up
    part = [torch.zeros(gdb.n_taxa, dtype=torch.int32, device=dev) for _ in range(2)]
    s0.ucount_or_range([p0, p1], 0, half, part[0])
    s1.ucount_or_range([p0, p1], half, n_words - half, part[1])
    torch.cuda.synchronize()
    assert np.array_equal((part[0] + part[1]).cpu().numpy(), want_u)
    # staged: OR both bitmaps' slice into a scratch buffer placed at the slice's own position
    scratch = torch.zeros(n_words, dtype=torch.int32, device=dev)
    part2 = torch.zeros(gdb.n_taxa, dtype=torch.int32, device=dev)
    for w0, n in ((0, half), (half, n_words - half)):
        s0.seen_or(scratch.data_ptr() + 4 * w0, [p0, p1], w0, n)
        s0.ucount_range(scratch.data_ptr(), w0, n, part2)
    torch.cuda.synchronize()
    assert np.array_equal(part2.cpu().numpy(), want_u)


def test_trim_adversarial_qualities(layout):
    """process_qual (:714-760) corner cases: every length 1..70 and longer ones, quality strings made
    of runs around the two thresholds (single base >= '1', 4-window sum(q-32) >= 68), good/bad
    islands near both ends, bytes >= 0x80 (negative as signed char).  Spans and drop decisions must
    match the oracle bit for bit - this exercises the fast path, the preloaded-register path and the
    32-positions-per-step fallback scans of the CUDA trim."""
    rng = np.random.default_rng(81)
    db = H.make_db(rng, 2000)
    odb, osamp = _oracle(db)
    gdb, gs = _gpu(db, layout)
    alphabet = np.array([33, 35, 47, 48, 49, 50, 51, 52, 53, 60, 73, 127, 128, 200, 255], dtype=np.uint8)
    seqs, quals = [], []
    lengths = list(range(1, 71)) * 40 + list(rng.integers(71, 400, size=3000)) + [511, 512, 513, 1000] * 10
    for L in lengths:
        L = int(L)
        mode = rng.integers(0, 5)
        if mode == 0:
            q = rng.choice(alphabet, size=L)
        elif mode == 1:  # runs
            q = np.repeat(rng.choice(alphabet, size=L), rng.integers(1, 9, size=L))[:L]
        elif mode == 2:  # good middle, ragged ends
            q = np.full(L, 73, np.uint8)
            a, b = int(rng.integers(0, min(L, 45))), int(rng.integers(0, min(L, 45)))
            q[:a] = rng.choice(alphabet[:9], size=a)
            q[L - b:] = rng.choice(alphabet[:9], size=b)
        elif mode == 3:  # borderline windows: values 48..52 only
            q = rng.integers(48, 53, size=L).astype(np.uint8)
        else:  # mostly bad with a few good islands
            q = np.full(L, 35, np.uint8)
            for _ in range(int(rng.integers(0, 4))):
                a = int(rng.integers(0, L))
                q[a:a + int(rng.integers(1, 40))] = 73
        seqs.append(np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=L)])
        quals.append(q[:L])
    off = np.concatenate([[0], np.cumsum([s.size for s in seqs])]).astype(np.uint64)
    batch = H.ReadBatch(seq=np.concatenate(seqs), qual=np.concatenate(quals), off=off, names=[])
    fin = _check_batch(gs, osamp, batch)
    _check_counts(gs, osamp)
    assert (fin == -1).sum() > 1000 and (fin >= 0).sum() > 1000


def test_many_taxa_global_histogram_path(layout):
    """More than 24 576 taxa: the per-block shared-memory gcount histogram no longer fits and the
    kernels fall back to global atomics; also exercises taxon ids close to the 21-bit field limit of
    the packed sectors."""
    rng = np.random.default_rng(101)
    n = 2_000_000  # taxa ids up to 2e6 (< 2^21)
    parent = np.ones(n, np.int32)
    inner = rng.integers(2, 1000, size=n)
    parent[1000:] = inner[1000:]          # leaves hang off 998 inner nodes ...
    parent[2:1000] = rng.integers(1, 2, size=998)  # ... which hang off the root
    keys = H.canonical(rng.integers(0, 1 << 60, size=20000, dtype=np.uint64))
    taxa = np.concatenate([rng.integers(1000, n, size=15000), rng.integers(2, 1000, size=5000)]).astype(np.uint32)
    db = H.SynthDB(keys=keys, taxa=taxa, parent=parent)
    odb, osamp = _oracle(db)
    gdb, gs = _gpu(db, layout)
    batch = H.make_reads(rng, db, 2000)
    fin = _check_batch(gs, osamp, batch)
    _check_counts(gs, osamp)
    assert fin.max() > 1_000_000


def test_hot_gcount_bin():
    """All reads of a large batch end in ONE gcount bin (7 M reads with no hit -> bin 0; 7 M copies of a
    read that hits one probe -> one leaf bin): the per-block shared-memory histograms and their
    flush must not lose or wrap counts (> 40 000 increments of one counter per block)."""
    import torch
    import kmer_id_b200 as kid
    rng = np.random.default_rng(77)
    db = H.make_db(rng, 500)
    gdb = kid.Database(db.keys, db.taxa, db.parent)
    n, L = 7_000_000, 31
    dev = torch.device("cuda:0")
    off = torch.arange(n + 1, dtype=torch.int64, device=dev) * L
    out = torch.empty(n, dtype=torch.int32, device=dev)
    # (a) poly-A reads: one k-mer each, not in the table
    s = kid.Sample(gdb)
    seq = torch.full((n * L + 64,), ord("A"), dtype=torch.uint8, device=dev)
    s.classify_device(seq, None, off, n, out, None, 0)
    g, u = s.counts()
    assert g[0] == n and g.sum() == n and u.sum() == 0
    # (b) every read is probe 7 followed by one base: two k-mers, the first one hits
    t = int(db.taxa[7])
    assert t > 1
    one = np.concatenate([H.key_to_bases(int(db.keys[7])), np.frombuffer(b"A", np.uint8)]).astype(np.uint8)
    seq[: n * L] = torch.from_numpy(np.tile(one, n)).to(dev)
    s.begin()
    s.classify_device(seq, None, off, n, out, None, 0)
    g, u = s.counts()
    v = int(out[0])  # every read is the same read: t, unless its second k-mer happens to hit as well
    assert v > 1 and g[v] == n and g.sum() == n and bool((out == v).all())
    assert u[t] == 1


@pytest.mark.parametrize("sub_bits", [3, 4])
def test_wide_minimizer_groups(sub_bits, monkeypatch):
    """Layout M for very large databases: a minimizer addresses 2 or 4 lines instead of one
    (kid_table2.cuh sub_bits; chosen automatically beyond 4e8 keys, forced here).  Same results,
    also on a table loaded so high that keys get displaced."""
    monkeypatch.setenv("KID_DB_SUB_BITS", str(sub_bits))
    rng = np.random.default_rng(300 + sub_bits)
    db = H.make_db(rng, 6000, n_dup=200, n_zero=50)
    odb, osamp = _oracle(db)
    for kw in ({}, {"log2_sectors": 12}):  # default size, and 4096 sectors for ~5800 keys
        gdb, gs = _gpu(db, LAYOUT_M, **kw)
        want = np.array([odb.lookup(int(k)) for k in db.keys[:2000]], dtype=np.uint32)
        assert np.array_equal(gdb.lookup(db.keys[:2000]), want)
        osamp.reset()
        _check_batch(gs, osamp, H.make_reads(rng, db, 3000, ragged=True))
        _check_counts(gs, osamp)


@pytest.mark.parametrize("kw", [dict(), dict(lower_rate=0.05, n_rate=0.02), dict(length=250, bad_tail=0.6), dict(ragged=True)])
def test_20mer_minimizers(kw, monkeypatch):
    """Layout M with 20-mer minimizers (11 windows) instead of 16-mers (KID_DB_MM=20, an option for
    very large databases): the table is placed differently, the answers are the same - also on a table
    loaded so high that keys get displaced."""
    monkeypatch.setenv("KID_DB_MM", "20")
    rng = np.random.default_rng(400 + len(kw))
    db = H.make_db(rng, 30000, n_dup=500, n_zero=50)
    odb, osamp = _oracle(db)
    for tkw in ({}, {"log2_sectors": 14}):  # default size, and 16384 sectors for ~30000 keys
        gdb, gs = _gpu(db, LAYOUT_M, **tkw)
        want = np.array([odb.lookup(int(k)) for k in db.keys[:3000]], dtype=np.uint32)
        assert np.array_equal(gdb.lookup(db.keys[:3000]), want)
        osamp.reset()
        _check_batch(gs, osamp, H.make_reads(rng, db, 3000, **kw))
        _check_counts(gs, osamp)
