"""kid_pack_reads (the host-side writer of packed read batches, include/kmer_id.h) against the oracle's
trim and a byte-by-byte statement of the format.  Host memory only - runs without a GPU."""
import numpy as np
import pytest

import helpers as H


def _kid():
    import kmer_id_b200 as kid
    return kid


def unpack(words, meta, r):
    """(tlen, flagged, codes[tlen], valid[tlen]) of read r of a packed batch"""
    w0 = int(meta[2 * r]) & 0x7FFFFFFF
    flagged = bool(int(meta[2 * r]) >> 31)
    tlen = int(meta[2 * r + 1])
    cw, vw = (tlen + 15) // 16, (tlen + 31) // 32
    codes = np.zeros(tlen, np.uint8)
    valid = np.ones(tlen, np.uint8)
    for i in range(tlen):
        codes[i] = (int(words[w0 + i // 16]) >> (30 - 2 * (i % 16))) & 3
        if flagged:
            valid[i] = (int(words[w0 + cw + i // 32]) >> (31 - (i % 32))) & 1
    pad_ok = True
    if tlen % 16:
        pad_ok &= (int(words[w0 + cw - 1]) & ((1 << (2 * (16 - tlen % 16))) - 1)) == 0
    if flagged and tlen % 32:
        pad_ok &= (int(words[w0 + cw + vw - 1]) & ((1 << (32 - tlen % 32)) - 1)) == 0
    return tlen, flagged, codes, valid, pad_ok, cw + (vw if flagged else 0)


def expected(seq, start, stop, accept_u):
    lut_c = np.zeros(256, np.uint8)
    lut_v = np.zeros(256, np.uint8)
    for ch, c in zip(b"ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
        lut_c[ch], lut_v[ch] = c, 1
    if accept_u:
        for ch in b"Uu":
            lut_c[ch], lut_v[ch] = 3, 1
    s = seq[start:stop + 1]
    return lut_c[s] * lut_v[s], lut_v[s]


def check_batch(batch, flags=0, impls=(0, 0x100, 0x200), word0=0):
    from oracle import kor
    kid = _kid()
    outs = []
    for impl in impls:
        words, meta, span = kid.pack_reads(batch.seq, batch.qual, batch.off, flags | impl, word0=word0, want_span=True)
        outs.append((words.copy(), meta.copy(), span.copy()))
    for o in outs[1:]:
        assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1]) and np.array_equal(o[2], outs[0][2])
    words, meta, span = outs[0]
    words = np.concatenate([np.zeros(word0, np.uint32), words])  # index as meta does
    nxt = word0
    for r in range(batch.n):
        a, b = int(batch.off[r]), int(batch.off[r + 1])
        st, sp = kor.trim(batch.qual[a:b].tobytes(), b - a) if batch.qual is not None else (0, b - a - 1)
        assert (int(span[r, 0]), int(np.int32(span[r, 1]))) == (st, sp), r
        tlen, flagged, codes, valid, pad_ok, used = unpack(words, meta, r)
        assert (int(meta[2 * r]) & 0x7FFFFFFF) == nxt, "words must be dense"
        if sp - st < 30:
            assert tlen == 0 and not flagged
            continue
        assert tlen == sp - st + 1
        ec, ev = expected(batch.seq[a:b], st, sp, bool(flags & 1))
        assert flagged == bool((ev == 0).any())
        assert np.array_equal(codes, ec) and np.array_equal(valid, ev) and pad_ok, r
        nxt += used
    assert int(meta[2 * batch.n]) == nxt and int(meta[2 * batch.n + 1]) == 0
    return words, meta


@pytest.mark.parametrize("seed,n,kw", [
    (1, 600, dict()),
    (2, 400, dict(ragged=True)),
    (3, 400, dict(lower_rate=0.1, n_rate=0.02)),
    (4, 300, dict(length=250, bad_tail=0.8)),
    (5, 300, dict(length=31, bad_tail=0.5)),
])
def test_pack_matches_format_and_oracle_trim(seed, n, kw):
    rng = np.random.default_rng(seed)
    db = H.make_db(rng, 2000)
    check_batch(H.make_reads(rng, db, n, **kw))


def test_pack_every_byte_value_and_u_flag():
    rng = np.random.default_rng(6)
    L = 200
    seqs = [rng.integers(0, 256, size=L, dtype=np.uint8) for _ in range(60)]
    seqs.append(np.frombuffer(b"ACGTUacgtuNn.-*RYKM" * 11, dtype=np.uint8)[:L].copy())
    seq = np.concatenate(seqs)
    qual = np.full(seq.size, ord("I"), np.uint8)
    off = (np.arange(len(seqs) + 1) * L).astype(np.uint64)
    batch = H.ReadBatch(seq=seq, qual=qual, off=off, names=[b"@x"] * len(seqs))
    check_batch(batch, flags=0)
    check_batch(batch, flags=1)  # KID_DB_ACCEPT_U


def test_pack_adversarial_qualities_every_length():
    rng = np.random.default_rng(7)
    seqs, quals = [], []
    for L in list(range(1, 71)) + [127, 128, 129, 511, 512, 513, 1000, 2047, 2048, 2049, 5000]:
        for _ in range(3):
            seqs.append(H._BASES[rng.integers(0, 4, size=L)])
            quals.append(rng.choice(np.array([33, 40, 47, 48, 49, 50, 52, 53, 54, 73, 128, 255], np.uint8), size=L))
    off = np.concatenate([[0], np.cumsum([s.size for s in seqs])]).astype(np.uint64)
    batch = H.ReadBatch(seq=np.concatenate(seqs), qual=np.concatenate(quals), off=off, names=[b"@x"] * len(seqs))
    check_batch(batch, word0=12345)


def test_pack_without_quality_keeps_whole_reads():
    rng = np.random.default_rng(8)
    db = H.make_db(rng, 500)
    b = H.make_reads(rng, db, 100, ragged=True, n_rate=0.01)
    check_batch(H.ReadBatch(seq=b.seq, qual=None, off=b.off, names=b.names))


def test_pack_errors_and_bound():
    kid = _kid()
    import ctypes as C
    seq = np.frombuffer(b"ACGT" * 40, dtype=np.uint8)
    qual = np.full(160, ord("I"), np.uint8)
    off = np.array([0, 160], np.uint64)
    assert kid.lib.kid_pack_bound(1, 160) >= 10 + 5
    words = np.zeros(4, np.uint32)
    meta = np.zeros(4, np.uint32)
    nw = C.c_size_t(0)
    rc = kid.lib.kid_pack_reads(seq.ctypes.data, qual.ctypes.data, off.ctypes.data, 1, 0, 0, words.ctypes.data, 4,
                                meta.ctypes.data, None, C.byref(nw))
    assert rc == -3  # KID_ENOMEM: words_cap too small
    words = np.zeros(32, np.uint32)
    rc = kid.lib.kid_pack_reads(seq.ctypes.data, qual.ctypes.data, off.ctypes.data, 1, 0, 0x7FFFFFFF, words.ctypes.data, 32,
                                meta.ctypes.data, None, C.byref(nw))
    assert rc == -4  # KID_ERANGE: word indices must stay below 2^31
    # n_reads = 0 still writes the end entry
    rc = kid.lib.kid_pack_reads(None, None, None, 0, 0, 7, None, 0, meta.ctypes.data, None, C.byref(nw))
    assert rc == 0 and nw.value == 0 and meta[0] == 7 and meta[1] == 0


# ------------------------------------------------------------------------------------ dense batches
def unpack_dense(b, r):
    """(tlen, flagged, codes[tlen], valid[tlen]) of read r of a kid.DenseBatch"""
    a, e = int(b.boff[r]), int(b.boff[r + 1])
    tlen = e - a
    flagged = bool((int(b.flagbits[r >> 5]) >> (r & 31)) & 1)
    pos = np.arange(a, e)
    codes = ((b.codes[pos >> 4] >> (30 - 2 * (pos & 15)).astype(np.uint32)) & 3).astype(np.uint8)
    valid = np.ones(tlen, np.uint8)
    inv = b.inv[:b.n_inv]
    mine = inv[(inv >= a) & (inv < e)]
    valid[mine - a] = 0
    return tlen, flagged, codes, valid


@pytest.mark.parametrize("seed,n,kw,pieces", [
    (41, 600, dict(), 1),
    (42, 500, dict(ragged=True), 7),
    (43, 400, dict(lower_rate=0.1, n_rate=0.02), 3),
    (44, 300, dict(length=250, bad_tail=0.8), 5),
    (45, 300, dict(length=31, bad_tail=0.5), 2),
])
def test_dense_packer_equals_word_packer(seed, n, kw, pieces):
    """kid_pack_reads_dense (no padding, offsets only, non-ACGT bases as a position list) carries exactly
    what kid_pack_reads does, read for read - also when a batch is appended to in several calls (base and
    read positions that are not multiples of 16 / 32), and with every packer implementation."""
    kid = _kid()
    rng = np.random.default_rng(seed)
    db = H.make_db(rng, 2000)
    batch = H.make_reads(rng, db, n, **kw)
    words, meta, span = kid.pack_reads(batch.seq, batch.qual, batch.off, want_span=True)
    for impl in (0, kid.KID_PACK_IMPL_BYTES, kid.KID_PACK_IMPL_SWAR):
        dense = kid.DenseBatch(batch.n, int(batch.off[-1]), flags=impl, max_inv=int(batch.off[-1]))
        cuts = sorted(set([0, batch.n] + [int(x) for x in rng.integers(0, batch.n, size=pieces - 1)]))
        spans = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            o = batch.off[a:b + 1] - batch.off[a]
            lo, hi = int(batch.off[a]), int(batch.off[b])
            spans.append(dense.append(batch.seq[lo:hi], batch.qual[lo:hi], o, want_span=True))
        assert dense.n == batch.n and np.array_equal(np.concatenate(spans), span)
        inv = dense.inv[:dense.n_inv]
        assert np.all(np.diff(inv.astype(np.int64)) > 0)
        for r in range(batch.n):
            ht, hf, hc, hv, _, _ = unpack(words, meta, r)
            dt, df, dc, dv = unpack_dense(dense, r)
            assert (dt, df) == (ht, hf), r
            assert np.array_equal(dc, hc) and np.array_equal(dv, hv), r
        assert dense.n_bases == int(dense.boff[batch.n]) == sum(int(meta[2 * r + 1]) for r in range(batch.n))
        # nothing but zeros behind the last base
        tail = dense.codes[(dense.n_bases + 15) >> 4:((dense.n_bases + 15) >> 4) + 2]
        assert not tail.any()
        if dense.n_bases & 15:
            assert (int(dense.codes[dense.n_bases >> 4]) & ((1 << (2 * (16 - (dense.n_bases & 15)))) - 1)) == 0


def test_dense_packer_errors():
    kid = _kid()
    seq = np.frombuffer(b"ACGN" * 40, dtype=np.uint8)
    qual = np.full(160, ord("I"), np.uint8)
    off = np.array([0, 160], np.uint64)
    small = kid.DenseBatch(4, 160, max_inv=3)
    with pytest.raises(kid.KidError):  # 40 positions do not fit 3
        small.append(seq, qual, off)
    tiny = kid.DenseBatch(4, 160, codes=np.zeros(4, np.uint32))
    with pytest.raises(kid.KidError):
        tiny.append(seq, qual, off)
