"""kmer_id_b200/host/pgz.cpp (multi-threaded inflate) against zlib.

Through tests/hosttest/host_dump (no GPU): the raw inflater on every stream shape zlib can write,
its refusal (-> zlib fallback) on damaged files, and GzLineBlocks - the only consumer - with the
fallback in place.  The reference semantics being protected: the bytes gzread() would deliver
(newkmer_10nx.cpp:675-707, :770-810)."""
import gzip
import os
import subprocess
import zlib

import numpy as np
import pytest

import helpers as H

HD = os.path.join(H.ROOT, "tests", "hosttest", "host_dump")


@pytest.fixture(scope="module", autouse=True)
def _build():
    subprocess.run(["make", "-C", os.path.dirname(HD)], check=True, capture_output=True)


def _probe_text(rng, n_lines):
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)
    return b"".join(bases[rng.integers(0, 4, 30 + int(rng.integers(0, 40)))].tobytes() +
                    b",%d,%d,%d,F,1\n" % (rng.integers(2, 5000), i % 7, i) for i in range(n_lines))


def _fastq_text(rng, n):
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)
    return b"".join(b"@r%d/1\n%s\n+\n%s\n" % (i, bases[rng.integers(0, 4, 150)].tobytes(),
                                               rng.integers(33, 74, 150).astype(np.uint8).tobytes())
                    for i in range(n))


def _gz(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, 31, 9, strategy)
    return c.compress(data) + c.flush()


def _gunzip(tmp_path, comp, threads=4, piece=32768):
    src = str(tmp_path / "in.gz")
    dst = str(tmp_path / "out.bin")
    with open(src, "wb") as f:
        f.write(comp)
    r = subprocess.run([HD, "gunzip", src, str(threads), str(piece), dst], capture_output=True)
    with open(dst, "rb") as f:
        return r.returncode, f.read(), r.stderr.decode()


@pytest.fixture(scope="module")
def texts():
    rng = np.random.default_rng(99)
    return _probe_text(rng, 60000), _fastq_text(rng, 15000)


@pytest.mark.parametrize("level", [1, 6, 9])
def test_plain_gzip_levels(tmp_path, texts, level):
    for raw in texts:
        rc, got, log = _gunzip(tmp_path, gzip.compress(raw, level))
        assert rc == 0 and got == raw, log
        # ordinary zlib output must really go through the speculative path
        found, again = int(log.split("pieces")[1].split()[0]), int(log.split("again")[1].split()[0])
        assert found > 10 and again <= 2, log


def test_stream_shapes(tmp_path, texts):
    probes, fastq = texts
    rnd = np.random.default_rng(5).integers(0, 256, 1_000_000, dtype=np.uint8).tobytes()
    c = zlib.compressobj(6, zlib.DEFLATED, 31)
    flushed = b"".join(c.compress(fastq[i:i + 70000]) + c.flush(zlib.Z_SYNC_FLUSH if (i // 70000) % 2 else zlib.Z_FULL_FLUSH)
                       for i in range(0, len(fastq), 70000)) + c.flush()
    body = gzip.compress(probes[:800000], 6)[10:]
    fancy_header = bytes([0x1f, 0x8b, 8, 4 | 8 | 16, 0, 0, 0, 0, 0, 3, 5, 0]) + b"EXTRA" + b"name.txt\0" + b"comment\0"
    cases = {
        "members": (b"".join(gzip.compress(probes[i:i + 250000], 6) for i in range(0, len(probes), 250000)), probes),
        "tiny members": (b"".join(gzip.compress(fastq[i:i + 40000], 5) for i in range(0, len(fastq), 40000)), fastq),
        "stored": (gzip.compress(probes[:900000], 0), probes[:900000]),
        "incompressible": (gzip.compress(rnd, 6), rnd),
        "fixed codes": (_gz(probes[:900000], 6, zlib.Z_FIXED), probes[:900000]),
        "huffman only": (_gz(probes[:900000], 6, zlib.Z_HUFFMAN_ONLY), probes[:900000]),
        "rle": (_gz(probes[:900000], 6, zlib.Z_RLE), probes[:900000]),
        "flushes": (flushed, fastq),
        "long matches": (gzip.compress(b"A" * 3_000_000, 6), b"A" * 3_000_000),
        "header fields": (fancy_header + body, probes[:800000]),
        "empty members": (gzip.compress(probes[:300000]) + gzip.compress(b"") + gzip.compress(probes[300000:500000]) +
                          gzip.compress(b""), probes[:500000]),
    }
    for name, (comp, raw) in cases.items():
        rc, got, log = _gunzip(tmp_path, comp)
        assert rc == 0 and got == raw, (name, log)


@pytest.mark.parametrize("threads,piece", [(8, 2048), (3, 20000), (2, 1 << 20)])
def test_piece_sizes_and_threads(tmp_path, texts, threads, piece):
    rc, got, log = _gunzip(tmp_path, gzip.compress(texts[1], 6), threads, piece)
    assert rc == 0 and got == texts[1], log


def test_damaged_files_are_left_to_zlib(tmp_path, texts):
    fastq = texts[1]
    g = gzip.compress(fastq, 6)
    crc = bytearray(g); crc[-6] ^= 0xff
    flip = bytearray(g); flip[len(g) // 2] ^= 0x55
    zeros = bytes(30_000_000)
    for name, comp in {"truncated": g[:len(g) // 2], "bad crc": bytes(crc), "trailing bytes": g + b"garbage!",
                       "flipped byte": bytes(flip), "absurd expansion": gzip.compress(zeros, 9)}.items():
        rc, got, log = _gunzip(tmp_path, comp)
        assert rc == 4 and "GIVEUP" in log, (name, log)
        zo, zout = zlib.decompressobj(31), []
        try:
            for i in range(0, len(comp), 1024):
                zout.append(zo.decompress(comp[i:i + 1024]))
        except zlib.error:
            pass
        zout = b"".join(zout)
        m = min(len(zout), len(got))
        assert got[:m] == zout[:m], name  # what was delivered is what zlib delivers too


def test_not_gzip_is_not_applicable(tmp_path):
    p = str(tmp_path / "plain.txt")
    with open(p, "wb") as f:
        f.write(b"ACGT\n" * 100000)
    r = subprocess.run([HD, "gunzip", p, "4", "65536"], capture_output=True)
    assert r.returncode == 3


def _lines(path, threads, env=None):
    out = path + f".lines{threads}"
    r = subprocess.run([HD, "lines", path, str(threads), out], capture_output=True, env={**os.environ, **(env or {})})
    with open(out, "rb") as f:
        return r.returncode, f.read(), r.stderr.decode()


def test_line_blocks_parallel_equals_zlib(tmp_path, texts):
    small = {"KID_GZ_MIN_BYTES": "0", "KID_GZ_PIECE_BYTES": "30000"}
    probes = texts[0]
    for name, raw in {"terminated": probes, "unterminated tail": probes + b"ACGTACGT,no newline",
                      "crlf": probes.replace(b"\n", b"\r\n")}.items():
        p = str(tmp_path / "f.gz")
        with open(p, "wb") as f:
            f.write(gzip.compress(raw, 6))
        rc1, one, _ = _lines(p, 1, small)
        rc8, many, _ = _lines(p, 8, small)
        want = raw[:raw.rfind(b"\n") + 1]  # the tail after the last newline is never delivered
        assert rc1 == 0 and rc8 == 0 and one == want and many == want, name


def test_line_blocks_fall_back_like_zlib(tmp_path, texts):
    small = {"KID_GZ_MIN_BYTES": "0", "KID_GZ_PIECE_BYTES": "30000"}
    g = gzip.compress(texts[0], 6)
    bad = bytearray(g); bad[-6] ^= 0xff
    cut = g[:len(g) * 2 // 3]
    long_line = gzip.compress(texts[0][:200000] + b"A" * 20000 + b"\n" + texts[0][200000:], 6)
    for name, comp in {"bad crc": bytes(bad), "truncated": cut, "line over 16 KiB": long_line}.items():
        p = str(tmp_path / "f.gz")
        with open(p, "wb") as f:
            f.write(comp)
        rc1, _, err1 = _lines(p, 1, small)
        rc8, _, err8 = _lines(p, 8, small)
        assert rc1 == 255 and rc8 == 255 and err1 == err8 and err1.strip(), (name, err1, err8)


def test_lines_longer_than_a_piece(tmp_path):
    """Lines of up to 15 KB (under the reference's 16 KiB limit) with 1-KiB... 4-KiB inflated pieces: one
    line is assembled from several pieces, and blank lines / CRLF survive untouched."""
    rng = np.random.default_rng(123)
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)
    parts = []
    for i in range(400):
        parts.append(b">contig%d\n" % i)
        parts.append(bases[rng.integers(0, 4, int(rng.integers(1, 15000)))].tobytes() + (b"\r\n" if i % 3 == 0 else b"\n"))
        if i % 7 == 0:
            parts.append(b"\n")
    raw = b"".join(parts)
    p = str(tmp_path / "long.gz")
    with open(p, "wb") as f:
        f.write(gzip.compress(raw, 6))
    env = {"KID_GZ_MIN_BYTES": "0", "KID_GZ_PIECE_BYTES": "1024"}
    rc1, one, _ = _lines(p, 1, env)
    rc8, many, _ = _lines(p, 8, env)
    assert rc1 == 0 and rc8 == 0 and one == raw and many == raw
