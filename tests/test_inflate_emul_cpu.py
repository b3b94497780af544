"""The device-side inflate (kmer_id_b200/csrc/kid_inflate.cuh) run on the CPU by tests/hosttest/inflate_emul.cpp
the way kid_ingest.cu runs it on the GPU - block finder, per-piece decode to symbols + copy codes, the copy
pass, the chain walk, grouped window maps, marker resolve, CRC-32 in chunks - and a lane-by-lane restatement of
the warp-per-piece decoder the device uses inside blocks, against zlib on gzip files of
every kind a FASTQ reader meets (process_fqgz reads them through gzread, newkmer_10nx.cpp:770-780).
Exit codes of the harness: 0 = text identical to zlib's, 3 = cleanly refused (the product then uses the host
reader), anything else = a wrong answer."""
import gzip
import os
import random
import subprocess
import zlib

import pytest

import helpers as H

HT = os.path.join(H.ROOT, "tests", "hosttest")
EXE = os.path.join(HT, "inflate_emul")


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", EXE, os.path.join(HT, "inflate_emul.cpp"), "-lz"],
                   check=True)
    return EXE


def _fastq(n, seed, length=150):
    rnd = random.Random(seed)
    out = []
    for i in range(n):
        s = "".join(rnd.choice("ACGT") for _ in range(length))
        q = "".join(rnd.choice("FFFFFFFF:,#") for _ in range(length))
        out.append("@M0:%d:000-X:1:%d:%d 1:N:0:1\n%s\n+\n%s\n" % (i, rnd.randint(1000, 30000), rnd.randint(1000, 30000), s, q))
    return "".join(out).encode()


def _run(exe, path, *args, env=None):
    r = subprocess.run([exe, path, *map(str, args)], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, **(env or {})))
    return r.returncode, r.stdout


def _deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, flush_every=0):
    co = zlib.compressobj(level, zlib.DEFLATED, 31, 8, strategy)
    if not flush_every:
        return co.compress(data) + co.flush()
    out = b""
    for i in range(0, len(data), flush_every):
        out += co.compress(data[i:i + flush_every]) + co.flush(zlib.Z_SYNC_FLUSH)
    return out + co.flush()


@pytest.fixture(scope="module")
def text():
    return _fastq(12000, 5)


@pytest.mark.parametrize("piece", [4096, 32768])
@pytest.mark.parametrize("kind", ["l1", "l6", "l9", "members", "bgzf_like", "sync_flush", "huffman_only", "rle"])
def test_identical_to_zlib(emul, text, tmp_path, kind, piece):
    p = str(tmp_path / (kind + ".gz"))
    if kind in ("l1", "l6", "l9"):
        blob = gzip.compress(text, int(kind[1]))
    elif kind == "members":  # concatenated members; zlib ends each with a short fixed-code block
        blob = b"".join(gzip.compress(text[i:i + 700000], 6) for i in range(0, len(text), 700000))
    elif kind == "bgzf_like":  # many small members
        blob = b"".join(gzip.compress(text[i:i + 65280], 6) for i in range(0, len(text), 65280))
    elif kind == "sync_flush":  # what pigz writes between its blocks: empty stored blocks inside the stream
        blob = _deflate(text, flush_every=100000)
    elif kind == "huffman_only":
        blob = _deflate(text, strategy=zlib.Z_HUFFMAN_ONLY)
    else:
        blob = _deflate(text, strategy=zlib.Z_RLE)
    open(p, "wb").write(blob)
    # the block finder's text-only filter changes speculation, not results; KIDZ_WARP: the warp-per-piece decoder
    # of kid_ingest.cu (every lane decodes the token that would start at its bit, the real ones are chained from
    # lane 0), its lanes played one after the other on the CPU
    for env in ({}, {"KIDZ_TEXT_ONLY": "1"}, {"KIDZ_WARP": "1"}):
        rc, out = _run(emul, p, piece, env=env)
        assert rc == 0, out
        assert "crc ok" in out and "identical to zlib's" in out


def test_small_files(emul, tmp_path):
    for name, data in (("empty", b""), ("tiny", b"@a\nACGT\n+\nFFFF\n"), ("one_line", b"x" * 70000 + b"\n")):
        p = str(tmp_path / (name + ".gz"))
        open(p, "wb").write(gzip.compress(data))
        rc, out = _run(emul, p)
        assert rc == 0, out


def test_refused_not_wrong(emul, text, tmp_path):
    """files this decoder leaves to zlib: every one must be REFUSED (rc 3), never inflated to different bytes"""
    good = gzip.compress(text, 6)
    bad_crc = bytearray(good)
    bad_crc[-6] ^= 1
    flipped = bytearray(good)
    flipped[len(good) // 2] ^= 0x10
    cases = {
        "truncated": good[: len(good) // 2],
        "bad_crc": bytes(bad_crc),
        "bit_flip": bytes(flipped),
        "trailing_bytes": good + b"\0\0\0\0hello",
        "stored_only": gzip.compress(text[:3000000], 0),
        "fixed_only": _deflate(text[:3000000], strategy=zlib.Z_FIXED),
        "zeros": gzip.compress(b"\0" * 5000000, 9),  # expands beyond a piece's room
        "not_gzip": text[:100000],
    }
    for name, blob in cases.items():
        p = str(tmp_path / (name + ".gz"))
        open(p, "wb").write(blob)
        for env in ({}, {"KIDZ_WARP": "1"}):
            rc, out = _run(emul, p, env=env)
            assert rc == 3, (name, env, rc, out)


def test_random_piece_sizes(emul, text, tmp_path):
    p = str(tmp_path / "r.gz")
    open(p, "wb").write(b"".join(gzip.compress(text[i:i + 1500000], 1 + (i // 1500000) % 9) for i in range(0, len(text), 1500000)))
    rnd = random.Random(11)
    for _ in range(4):
        piece = rnd.choice([1024, 3000 & ~3, 8192, 20000, 65536])
        rc, out = _run(emul, p, piece, 12, rnd.choice([1, 3, 64]), env=rnd.choice([{}, {"KIDZ_WARP": "1"}]))
        assert rc == 0, (piece, out)
