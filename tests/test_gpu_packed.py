"""Packed read batches on the GPU (through the C-ABI): the device packer against the host packer, and
the k-mer scan over host-packed batches against the CPU oracle - bit-exact per-read taxa, gcount,
ucount, lookup and hit counters.  (Text batches take the same scan after kid_pack_kernel, so
tests/test_gpu_parity.py covers device packer + scan end to end.)"""
import numpy as np
import pytest

import helpers as H
from test_pack_host_cpu import unpack

pytestmark = pytest.mark.gpu


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


def _check_counts(gs, osamp):
    g, u = gs.counts()
    assert np.array_equal(g, osamp.gcount)
    assert np.array_equal(u, osamp.ucount)
    c = gs.counters()
    assert (c["lookups"], c["hits"], c["reads"]) == (osamp.lookups, osamp.hits, osamp.tct)


CASES = [
    (21, 4000, dict()),
    (22, 3000, dict(ragged=True)),
    (23, 3000, dict(lower_rate=0.05, n_rate=0.02)),          # most reads carry validity words
    (24, 2000, dict(length=250, bad_tail=0.6)),
    (25, 1500, dict(length=31)),
    (26, 1500, dict(on_target=1.0, sub_rate=0.0, n_rate=0.0)),
    (27, 300, dict(length=3000, n_rate=0.002)),               # longer than the strip: window path
    (28, 40, dict(length=70000, n_rate=0.0005, bad_tail=0.0)),
]


@pytest.mark.parametrize("seed,n,kw", CASES)
def test_device_packer_equals_host_packer(seed, n, kw):
    import torch
    import kmer_id_b200 as kid
    rng = np.random.default_rng(seed)
    db = H.make_db(rng, 3000)
    gdb, _ = _gpu(db)
    batch = H.make_reads(rng, db, min(n, 1500), **kw)
    seq, qual = batch.padded()
    hw, hm, hspan = kid.pack_reads(batch.seq, batch.qual, batch.off, want_span=True)
    total = int(batch.off[-1])
    cap = int(kid.lib.kid_pack_bound(batch.n, total))
    dseq, dqual = torch.from_numpy(seq).cuda(), torch.from_numpy(qual).cuda()
    doff = torch.from_numpy(batch.off.astype(np.int64)).cuda()
    dwords = torch.full((cap,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
    dmeta = torch.zeros(2 * (batch.n + 1), dtype=torch.int32, device="cuda")
    dspan = torch.zeros(2 * batch.n, dtype=torch.int32, device="cuda")
    gdb.pack_device(dseq, dqual, doff, total, batch.n, dwords, cap, dmeta, dspan)
    torch.cuda.synchronize()
    gw = dwords.cpu().numpy().view(np.uint32)
    gm = dmeta.cpu().numpy().view(np.uint32)
    assert np.array_equal(dspan.cpu().numpy().view(np.uint32).reshape(-1, 2), hspan)
    for r in range(batch.n):
        ht, hf, hc, hv, hp, _ = unpack(hw, hm, r)
        gt, gf, gc, gv, gp, _ = unpack(gw, gm, r)
        assert gt == ht and gp, r
        # the device packer also flags every read that takes its one-at-a-time path; never fewer
        assert gf == hf, r
        assert np.array_equal(gc, hc) and np.array_equal(gv, hv), r
        w0 = int(gm[2 * r]) & 0x7FFFFFFF
        rel = int(batch.off[r])
        assert w0 == rel // 16 + rel // 32 + 2 * r
    assert (int(gm[2 * batch.n]) & 0x7FFFFFFF) <= cap


@pytest.mark.parametrize("seed,n,kw", CASES)
def test_packed_classify_matches_oracle(seed, n, kw):
    import kmer_id_b200 as kid
    rng = np.random.default_rng(seed)
    db = H.make_db(rng, 30000, n_dup=500, n_zero=50)
    odb, osamp = _oracle(db)
    gdb, gs = _gpu(db)
    batch = H.make_reads(rng, db, n, **kw)
    fin_o, _ = osamp.classify(batch.seq, batch.qual, batch.off)
    words, meta = kid.pack_reads(batch.seq, batch.qual, batch.off)
    out = np.full(batch.n, -2, np.int32)
    gs.classify_packed_host(words, meta, batch.n, out)
    bad = np.flatnonzero(out != fin_o)
    assert bad.size == 0, f"{bad.size} reads differ, first {bad[:5]}: oracle {fin_o[bad[:5]]} gpu {out[bad[:5]]}"
    _check_counts(gs, osamp)
    # chunked (3 reads per chunk makes every group ragged), word0 != 0, repeat must not bump ucount
    gs.set_chunk_reads(1000 if batch.n > 1000 else 7)
    words2, meta2 = kid.pack_reads(batch.seq, batch.qual, batch.off, word0=1000)
    out2 = np.full(batch.n, -2, np.int32)
    gs.classify_packed_host(words2, meta2, batch.n, out2, word0=1000)
    osamp.classify(batch.seq, batch.qual, batch.off)
    assert np.array_equal(out2, fin_o)
    _check_counts(gs, osamp)


def test_packed_device_entry_and_text_entry_agree():
    import torch
    import kmer_id_b200 as kid
    rng = np.random.default_rng(31)
    db = H.make_db(rng, 20000)
    odb, osamp = _oracle(db)
    gdb, gs = _gpu(db)
    batch = H.make_reads(rng, db, 5000, n_rate=0.003)
    fin_o, span_o = osamp.classify(batch.seq, batch.qual, batch.off)
    words, meta = kid.pack_reads(batch.seq, batch.qual, batch.off)
    dw = torch.from_numpy(words.view(np.int32).copy()).cuda()
    dm = torch.from_numpy(meta.view(np.int32).copy()).cuda()
    dout = torch.zeros(batch.n, dtype=torch.int32, device="cuda")
    gs.classify_packed_device(dw, dm, batch.n, dout)
    torch.cuda.synchronize()
    assert np.array_equal(dout.cpu().numpy(), fin_o)
    _check_counts(gs, osamp)
    # same reads as text through the device entry point: pack kernel + scan
    gs.begin()
    seq, qual = batch.padded()
    dseq, dqual = torch.from_numpy(seq).cuda(), torch.from_numpy(qual).cuda()
    doff = torch.from_numpy(batch.off.astype(np.int64)).cuda()
    dspan = torch.zeros(2 * batch.n, dtype=torch.int32, device="cuda")
    dout.zero_()
    gs.classify_device(dseq, dqual, doff, batch.n, dout, dspan)
    torch.cuda.synchronize()
    assert np.array_equal(dout.cpu().numpy(), fin_o)
    assert np.array_equal(dspan.cpu().numpy().view(np.uint32).reshape(-1, 2), span_o.astype(np.uint32))
    _check_counts(gs, osamp)
    with pytest.raises(kid.KidError):  # unaligned text is refused, not read out of bounds
        gs.classify_device(dseq[1:], dqual[1:], doff, batch.n, dout, dspan)


def test_async_slots_pipeline():
    """One host thread, KID_MAX_SLOTS batches in flight: submit, keep packing, wait, reuse."""
    import kmer_id_b200 as kid
    rng = np.random.default_rng(32)
    db = H.make_db(rng, 20000)
    odb, osamp = _oracle(db)
    gdb, gs = _gpu(db)
    batches = [H.make_reads(rng, db, 700 + 100 * i, n_rate=0.002, ragged=(i % 3 == 0)) for i in range(9)]
    want = [osamp.classify(b.seq, b.qual, b.off)[0] for b in batches]
    outs = [np.full(b.n, -2, np.int32) for b in batches]
    held = {}
    for i, b in enumerate(batches):
        slot = i % kid.KID_MAX_SLOTS
        if slot in held:
            gs.wait(slot)
        if i % 2 == 0:  # packed submission
            words, meta = kid.pack_reads(b.seq, b.qual, b.off, word0=17 * i)
            held[slot] = (words, meta)
            gs.classify_packed_async(slot, words, meta, b.n, outs[i], word0=17 * i)
        else:           # text submission with spans
            seq, qual = b.padded()
            span = np.zeros((b.n, 2), np.uint32)
            held[slot] = (seq, qual, span)
            gs.classify_async(slot, seq, qual, b.off, b.n, outs[i], span)
    for slot in range(kid.KID_MAX_SLOTS):
        gs.wait(slot)
    for i in range(len(batches)):
        assert np.array_equal(outs[i], want[i]), i
    _check_counts(gs, osamp)
    with pytest.raises(kid.KidError):
        gs.wait(kid.KID_MAX_SLOTS)


def test_packed_rejected_for_keyhash_layout():
    import kmer_id_b200 as kid
    rng = np.random.default_rng(33)
    db = H.make_db(rng, 2000)
    gdb, gs = _gpu(db, flags=kid.KID_DB_LAYOUT_KEYHASH)
    b = H.make_reads(rng, db, 10)
    words, meta = kid.pack_reads(b.seq, b.qual, b.off)
    with pytest.raises(kid.KidError):
        gs.classify_packed_host(words, meta, b.n, None)


def test_more_taxa_than_the_shared_histogram_holds():
    """gcount lives in a shared-memory histogram up to 24 576 taxa (96 KB); beyond that the kernel adds to
    global memory directly.  40 000 taxa in a random recursive tree: same answers."""
    import kmer_id_b200 as kid
    rng = np.random.default_rng(34)
    n = 40_000
    parent = np.ones(n, np.int32)
    for v in range(2, n):
        parent[v] = int(rng.integers(1, v)) if v > 50 else max(1, v - 1)
    keys = H.canonical(rng.integers(0, 1 << 60, size=20000, dtype=np.uint64))
    db = H.SynthDB(keys=keys, taxa=rng.integers(2, n, size=keys.size).astype(np.uint32), parent=parent)
    odb, osamp = _oracle(db)
    gdb, gs = _gpu(db)
    batch = H.make_reads(rng, db, 3000, n_rate=0.002)
    fin_o, _ = osamp.classify(batch.seq, batch.qual, batch.off)
    words, meta = kid.pack_reads(batch.seq, batch.qual, batch.off)
    out = np.full(batch.n, -2, np.int32)
    gs.classify_packed_host(words, meta, batch.n, out)
    assert np.array_equal(out, fin_o)
    _check_counts(gs, osamp)
    assert (fin_o > 24_576).sum() > 100


@pytest.mark.parametrize("seed,n,kw", CASES)
def test_dense_classify_matches_oracle(seed, n, kw):
    """Dense batches (what the hosts ship: no padding, offsets only, non-ACGT positions as a list) through
    kid_expand_kernel + the scan: bit-exact against the oracle, whole and in chunks of 32 / 1024 reads."""
    import kmer_id_b200 as kid
    rng = np.random.default_rng(seed + 100)
    db = H.make_db(rng, 30000, n_dup=500, n_zero=50)
    odb, osamp = _oracle(db)
    gdb, gs = _gpu(db)
    batch = H.make_reads(rng, db, n, **kw)
    fin_o, _ = osamp.classify(batch.seq, batch.qual, batch.off)
    dense = kid.DenseBatch(batch.n, int(batch.off[-1]), max_inv=int(batch.off[-1]))
    half = batch.n // 3
    for a, b in ((0, half), (half, batch.n)):  # appended in two calls
        lo, hi = int(batch.off[a]), int(batch.off[b])
        dense.append(batch.seq[lo:hi], batch.qual[lo:hi], batch.off[a:b + 1] - batch.off[a])
    out = np.full(batch.n, -2, np.int32)
    gs.classify_dense_host(dense, out)
    bad = np.flatnonzero(out != fin_o)
    assert bad.size == 0, f"{bad.size} reads differ, first {bad[:5]}: oracle {fin_o[bad[:5]]} gpu {out[bad[:5]]}"
    _check_counts(gs, osamp)
    for chunk in (32, 1024):
        gs.set_chunk_reads(chunk)
        out2 = np.full(batch.n, -2, np.int32)
        gs.classify_dense_host(dense, out2)
        osamp.classify(batch.seq, batch.qual, batch.off)
        assert np.array_equal(out2, fin_o)
        _check_counts(gs, osamp)


def test_dense_async_slots():
    import kmer_id_b200 as kid
    rng = np.random.default_rng(35)
    db = H.make_db(rng, 20000)
    odb, osamp = _oracle(db)
    gdb, gs = _gpu(db)
    batches = [H.make_reads(rng, db, 500 + 77 * i, n_rate=0.003, ragged=(i % 2 == 0)) for i in range(7)]
    want = [osamp.classify(b.seq, b.qual, b.off)[0] for b in batches]
    outs = [np.full(b.n, -2, np.int32) for b in batches]
    held = {}
    for i, b in enumerate(batches):
        slot = i % kid.KID_MAX_SLOTS
        if slot in held:
            gs.wait(slot)
        d = kid.DenseBatch(b.n, int(b.off[-1]), max_inv=int(b.off[-1]))
        d.append(b.seq, b.qual, b.off)
        held[slot] = d
        gs.classify_dense_async(slot, d, outs[i])
    for slot in range(kid.KID_MAX_SLOTS):
        gs.wait(slot)
    for i in range(len(batches)):
        assert np.array_equal(outs[i], want[i]), i
    _check_counts(gs, osamp)


def test_block_strip_and_window_boundaries():
    """Read lengths on every internal boundary of the scan: a block is 128 k-mers (lengths 157-161,
    285-289), the strip holds 2048 bases (2040-2080, flagged reads carry validity words so their limit is
    lower: 1360-1372), a long read is walked in windows of 1920 k-mers (1948-1951, 3868-3871) - as text,
    as packed and as dense batches, with and without non-ACGT bases."""
    import kmer_id_b200 as kid
    rng = np.random.default_rng(36)
    db = H.make_db(rng, 20000)
    odb, osamp = _oracle(db)
    gdb, gs = _gpu(db)
    lengths = (list(range(155, 163)) + list(range(283, 291)) + list(range(1358, 1374)) + list(range(2040, 2082)) +
               [1947, 1948, 1949, 1950, 1951, 1952, 3867, 3868, 3869, 3870, 3871, 3872, 5789, 5790])
    seqs, quals = [], []
    keys_b = [H.key_to_bases(int(k)) for k in db.keys[:400]]
    for L in lengths:
        for with_n in (False, True):
            parts, tot = [], 0
            while tot < L:
                p = keys_b[int(rng.integers(0, len(keys_b)))] if rng.random() < 0.5 else H._BASES[rng.integers(0, 4, size=17)]
                parts.append(p)
                tot += p.size
            s = np.concatenate(parts)[:L].copy()
            if with_n:
                s[rng.integers(0, L, size=max(1, L // 300))] = ord("N")
                s[L - 1] = ord("n") if rng.random() < 0.5 else s[L - 1]
            seqs.append(s)
            quals.append(np.full(L, ord("I"), np.uint8))
    off = np.concatenate([[0], np.cumsum([s.size for s in seqs])]).astype(np.uint64)
    seq, qual = np.concatenate(seqs), np.concatenate(quals)
    n = len(seqs)
    fin_o, _ = osamp.classify(seq, qual, off)
    assert (fin_o > 0).sum() > n // 2  # most reads hit something (mixed lineages fold to the root, taxon 1)
    pad = np.zeros(16, np.uint8)
    out_t = gs.classify(np.concatenate([seq, pad]), np.concatenate([qual, pad]), off)
    assert np.array_equal(out_t, fin_o)
    _check_counts(gs, osamp)
    words, meta = kid.pack_reads(seq, qual, off)
    out_p = np.full(n, -2, np.int32)
    gs.classify_packed_host(words, meta, n, out_p)
    osamp.classify(seq, qual, off)
    assert np.array_equal(out_p, fin_o)
    _check_counts(gs, osamp)
    dense = kid.DenseBatch(n, int(off[-1]), max_inv=int(off[-1]))
    dense.append(seq, qual, off)
    for chunk in (1 << 18, 32):
        gs.set_chunk_reads(chunk)
        out_d = np.full(n, -2, np.int32)
        gs.classify_dense_host(dense, out_d)
        osamp.classify(seq, qual, off)
        assert np.array_equal(out_d, fin_o)
        _check_counts(gs, osamp)
