"""gz FASTQ files read on the device (kid_fastq_*, kmer_id_b200/csrc/kid_ingest.cu) against a Python
restatement of process_fqgz's framing (newkmer_10nx.cpp:762-816) + the CPU oracle: per-read taxa, gcount,
ucount, and the header / trimmed bases that go to _reads.txt (:608-611).  Files the device path must leave
to the host reader have to be refused before anything is counted."""
import gzip
import os
import zlib

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def frame_fastq(data: bytes):
    """(header, bases, qualities) per record the way process_fqgz hands them to process_qual"""
    recs, mod4, acc, seq = [], 0, None, None
    for line in data.split(b"\n")[:-1]:  # bytes after the last '\n' are never looked at (:812-813)
        if line.endswith(b"\r"):         # :786-787
            line = line[:-1]
        if len(line) > 0:                # :788 - empty lines do not advance mod4
            if mod4 == 0:
                acc = line
            elif mod4 == 1:
                seq = line
            elif mod4 == 3:
                recs.append((acc, seq, line))
            mod4 = (mod4 + 1) % 4
    return recs


def fastq_text(batch: H.ReadBatch, rng, crlf=False, blank_rate=0.0, plus_text=False):
    eol = b"\r\n" if crlf else b"\n"
    out = []
    for r in range(batch.n):
        a, b = int(batch.off[r]), int(batch.off[r + 1])
        lines = [batch.names[r], batch.seq[a:b].tobytes(), b"+" + (batch.names[r][1:] if plus_text else b""),
                 batch.qual[a:b].tobytes()]
        for ln in lines:
            if blank_rate and rng.random() < blank_rate:
                out.append(rng.choice([b"\n", b"\r\n", b"\n\n"]))
            out.append(ln + eol)
    return b"".join(out)


@pytest.fixture(scope="module")
def world():
    import kmer_id_b200 as kid
    from oracle import kor
    rng = np.random.default_rng(77)
    db = H.make_db(rng, 4000, n_dup=50, n_zero=20)
    gdb = kid.Database(db.keys, db.taxa, db.parent, device=0)
    odb = kor.OracleDB(db.n_taxa)
    odb.set_parents(db.parent)
    odb.add_keys(db.keys, db.taxa)
    return kid, kor, db, gdb, odb


def check_file(world, path, data):
    kid, kor, db, gdb, odb = world
    recs = [r for r in frame_fastq(data)]
    n = len(recs)
    seq = np.frombuffer(b"".join(r[1] for r in recs), dtype=np.uint8)
    qual = np.frombuffer(b"".join(r[2][: len(r[1])] for r in recs), dtype=np.uint8)
    off = np.concatenate([[0], np.cumsum([len(r[1]) for r in recs])]).astype(np.uint64)
    osamp = kor.OracleSample(odb)
    pad = np.zeros(16, np.uint8)
    fin_o, span_o = osamp.classify(np.concatenate([seq, pad]), np.concatenate([qual, pad]), off) if n else (np.zeros(0, np.int32), np.zeros((0, 2)))
    s = kid.Sample(gdb)
    fq = kid.GzFastq(gdb)
    assert fq.load(path), fq.why
    assert fq.n_reads == n
    fin = fq.classify(s)
    assert np.array_equal(fin, fin_o)
    g, u = s.counts()
    assert np.array_equal(g, osamp.gcount) and np.array_equal(u, osamp.ucount)
    kept = np.flatnonzero(fin >= 0)
    if kept.size:
        pick = kept[:: max(1, kept.size // 500)]
        got = fq.fetch(pick)
        span = np.asarray(span_o).reshape(-1, 2)
        for (h, b), r in zip(got, pick.tolist()):
            assert h == recs[r][0]
            assert b == recs[r][1][int(span[r][0]): int(span[r][1]) + 1]
    st = fq.stats()
    assert st["text_bytes"] == len(data)
    return fq, s, st


@pytest.mark.parametrize("case", ["plain", "crlf", "blank_lines", "no_final_newline", "cut_record", "ragged", "members", "level9"])
def test_device_reader_matches_reference_framing(world, tmp_path, case):
    kid, kor, db, gdb, odb = world
    rng = np.random.default_rng(abs(hash(case)) % 1000)
    batch = H.make_reads(rng, db, 6000, ragged=(case == "ragged"), lower_rate=0.01)
    data = fastq_text(batch, rng, crlf=(case == "crlf"), blank_rate=0.05 if case == "blank_lines" else 0.0,
                      plus_text=(case == "level9"))
    if case == "no_final_newline":
        data = data[:-1]  # the last quality line has no '\n': that record is never classified
    if case == "cut_record":
        data = data[: len(data) - 200]
    p = str(tmp_path / "a_R1_tr.fastq.gz")
    if case == "members":
        blob = b"".join(gzip.compress(data[i:i + 300000], 1 + (i // 300000) % 9) for i in range(0, len(data), 300000))
    else:
        blob = gzip.compress(data, 9 if case == "level9" else 1)
    open(p, "wb").write(blob)
    fq, s, st = check_file(world, p, data)
    assert st["pieces"] >= 10


def test_same_object_many_files(world, tmp_path):
    """buffers are reused from file to file (shrinking and growing), with read-ahead of the next file"""
    kid, kor, db, gdb, odb = world
    rng = np.random.default_rng(5)
    files = []
    for i, n in enumerate([3000, 200, 0, 8000, 1]):
        batch = H.make_reads(rng, db, n)
        data = fastq_text(batch, rng)
        p = str(tmp_path / ("f%d.gz" % i))
        open(p, "wb").write(gzip.compress(data, 1))
        files.append((p, data))
    fq = kid.GzFastq(gdb)
    for i, (p, data) in enumerate(files):
        assert fq.load(p), fq.why
        if i + 1 < len(files):
            fq.prefetch(files[i + 1][0])
        recs = frame_fastq(data)
        assert fq.n_reads == len(recs)
        s = kid.Sample(gdb)
        fin = fq.classify(s)
        g, _ = s.counts()
        assert int(g.sum()) == int((fin >= 0).sum())
    # a prefetched file that is then not asked for
    fq.prefetch(files[0][0])
    assert fq.load(files[3][0]) and fq.n_reads == len(frame_fastq(files[3][1]))


def test_files_for_the_host_reader(world, tmp_path):
    kid, kor, db, gdb, odb = world
    rng = np.random.default_rng(9)
    batch = H.make_reads(rng, db, 4000)
    data = fastq_text(batch, rng)
    good = gzip.compress(data, 6)
    bad_crc = bytearray(good)
    bad_crc[-6] ^= 1
    long_line = data[:5000] + b"@x\n" + b"A" * 20000 + b"\n+\n" + b"I" * 20000 + b"\n" + data[5000:]
    recs = frame_fastq(data)
    short_q = data.replace(recs[100][2] + b"\n", recs[100][2][:-3] + b"\n", 1)
    co = zlib.compressobj(6, zlib.DEFLATED, 31, 8, zlib.Z_FIXED)
    cases = {
        "truncated": good[: len(good) // 2],
        "bad_crc": bytes(bad_crc),
        "trailing": good + b"garbage",
        "not_gzip": data,
        "fixed_only": co.compress(data) + co.flush(),
        "long_line": gzip.compress(long_line, 6),
        "short_quality": gzip.compress(short_q, 6),
    }
    s = kid.Sample(gdb)
    fq = kid.GzFastq(gdb)
    for name, blob in cases.items():
        p = str(tmp_path / (name + ".gz"))
        open(p, "wb").write(blob)
        assert not fq.load(p), name
        assert "host reader" in fq.why
    assert not fq.load(str(tmp_path / "missing.gz"))
    g, u = s.counts()
    assert int(g.sum()) == 0 and int(u.sum()) == 0
    # and the object still works afterwards
    p = str(tmp_path / "good.gz")
    open(p, "wb").write(good)
    check_file(world, p, data)
