"""CPU-only checks of the C++ host logic (kmer_id_b200/host) through tests/hosttest/host_dump:
probe-line parsing against the oracle's istream emulation, the reference's 16-probe lookup cap
against a literal simulation of its table, and the four read-file parsers against the reference's
line rules restated in Python."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from oracle import kor

DUMP = os.path.join(H.ROOT, "tests", "hosttest", "host_dump")


@pytest.fixture(scope="module", autouse=True)
def _build():
    subprocess.run(["make", "-C", os.path.dirname(DUMP)], check=True, stdout=subprocess.DEVNULL)


def _dump(*args):
    r = subprocess.run([DUMP, *map(str, args)], capture_output=True, timeout=120)
    return r


WEIRD = [
    b"acgtacgtacgtacgtacgtacgtacgtac,7,0,0,F,1\n",
    b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT,9,1,2,F,1\n",
    b"ACGTACGTACGTNCGTACGTACGTACGTACGTACGTACGTAAAAAAAAAAAAAAAAAAAAAAAAAA,11,1,2,R,1\n",
    b"GGGGGGGGGGGGGGGGGGGGGGGGGGGGGG,12,1,2,F\n",
    b"TTTTTTTTTTTTTTTTTTTTTTTTTTTTTT,13,1,x,F,1\n",
    b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCCA 14 1 2 F 1\r\n",
    b"\n", b"   \n", b"\r\n",
    b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCCG,15,1,2,F1\n",
    b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCCT,16,1,2,F,1,extra,fields\n",
    b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCAA,00017,+1,-2,F,1\n",
    b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCAC,18,99999999999,2,F,1\n",   # int overflow: skipped
    b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCAG,4294967295,1,2,F,1\n",     # fits unsigned, not int
    b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCAT,-1,1,2,F,1\n",             # unsigned wrap / negative int
    b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCCC,19,1,2,,1\n",              # empty strand field: next token eaten
    b"  CCCCCCCCCCCCCCCCCCCCCCCCCCCCGA\t20\t1\t2\tF\t1\n",
    b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCGC,21,1,2,F,1x\n",
    b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCGG,0,1,2,F,1\n",              # taxon 0: kept in the list, invisible later
    b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAC,22,1,2,F,1",               # no newline: dropped
]


def test_probe_lines_match_oracle_parser(tmp_path):
    rng = np.random.default_rng(91)
    db = H.make_db(rng, 3000, n_dup=100, n_zero=20)
    path = os.path.join(str(tmp_path), "p.txt.gz")
    H.write_probes_gz(path, db, extra_lines=WEIRD)
    r = _dump("probes", path, 0, 0)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.decode().split("\n")
    n_lines = int(lines[0].split()[1])
    entries = [tuple(map(int, l.split())) for l in lines[1:] if l]
    odb = kor.OracleDB(1 << 22)
    assert odb.load_probes_gz(path) == n_lines
    first = {}
    for k, t in entries:
        if t != 0 and k not in first:
            first[k] = t
    assert len(first) == odb.n_keys
    for k, t in list(first.items())[:: max(1, len(first) // 2000)]:
        assert odb.lookup(k) == t
    # the signed-target variant (kmer_read_m3.cpp) differs exactly on the two odd target columns
    r2 = _dump("probes", path, 1, 0)
    n2 = int(r2.stdout.decode().split("\n")[0].split()[1])
    assert n2 == n_lines - 1  # 4294967295 does not fit an int -> that line is skipped


def _reference_depths(keys, taxa, log2_cells):
    """literal replay of Hashtable::add_kmer (kmer_read_m3.cpp:237-266) on a small table"""
    mask = (1 << log2_cells) - 1
    occ = set()
    out = []
    for k, t in zip(keys, taxa):
        if t == 0:
            out.append(None)
            continue
        h = int(k)
        h ^= h >> 33; h = (h * 0xff51afd7ed558ccd) & (2**64 - 1)
        h ^= h >> 33; h = (h * 0xc4ceb9fe1a85ec53) & (2**64 - 1)
        h ^= h >> 33
        reprobe, i = 0, 0
        while True:
            idx = (h + reprobe) & mask
            i += 1
            reprobe += i
            if idx not in occ:
                occ.add(idx)
                break
        out.append(i)
    return out


def test_probe_cap_replay_matches_literal_simulation(tmp_path):
    rng = np.random.default_rng(92)
    db = H.make_db(rng, 3400, n_dup=150, n_zero=30)
    path = os.path.join(str(tmp_path), "p.txt.gz")
    H.write_probes_gz(path, db)
    r = _dump("probes", path, 1, 12)
    lines = r.stdout.decode().split("\n")
    hidden = int(lines[0].split()[3])
    entries = [tuple(map(int, l.split())) for l in lines[1:] if l]
    depths = _reference_depths(db.keys.tolist(), db.taxa.tolist(), 12)
    first_depth = {}
    for k, t, d in zip(db.keys.tolist(), db.taxa.tolist(), depths):
        if t != 0 and k not in first_depth:
            first_depth[k] = d
    invisible = {k for k, d in first_depth.items() if d > 16}
    assert hidden == len(invisible) > 5, "fixture should push some keys beyond 16 probes"
    for (k, t), (k0, t0) in zip(entries, zip(db.keys.tolist(), db.taxa.tolist())):
        assert k == k0
        assert t == (0 if k in invisible else t0)


def _expect_fastq(data: bytes, plain: bool):
    """process_fqgz (newkmer_10nx.cpp:762-816) / process_fq (kmer_read_m3.cpp:895-931) line rules.
    In the plain readers `linestream >> lseq` leaves lseq UNTOUCHED when a line has no token (the
    sentry fails before the string is cleared), so a blank line repeats the previous token."""
    recs, mod4, acc, seq, tok = [], 0, b"", b"", b""
    lines = data.split(b"\n")
    if not plain:
        lines = lines[:-1]  # bytes after the last newline are never seen (:812-813)
    elif lines and lines[-1] == b"":
        lines = lines[:-1]
    for line in lines:
        if line.endswith(b"\r"):
            line = line[:-1]
        if plain:
            toks = line.split()
            tok = toks[0] if toks else tok
            line = tok
        if not line:
            continue
        if mod4 == 0:
            acc = line
        elif mod4 == 1:
            seq = line
        elif mod4 == 3:
            recs.append((acc, seq, line[: len(seq)]))
        mod4 = (mod4 + 1) % 4
    return recs


def _expect_fasta(data: bytes, plain: bool):
    recs, acc, seq, tok = [], b"", b"", b""
    lines = data.split(b"\n")
    lines = lines[:-1] if (not plain or (lines and lines[-1] == b"")) else lines
    for line in lines:
        if line.endswith(b"\r"):
            line = line[:-1]
        if plain:
            toks = line.split()
            tok = toks[0] if toks else tok  # a blank line repeats the previous token (see above)
            if tok[:1] == b">":
                if len(seq) > 30:
                    recs.append((acc, seq, b""))
                seq, acc = b"", tok[1:]
            else:
                seq += tok
        else:
            if not line:
                continue
            if line[:1] == b">":
                if len(seq) > 30:
                    recs.append((acc, seq, b""))
                seq, acc = b"", line[1:]
            else:
                seq += line
    if len(seq) > 30:
        recs.append((acc, seq, b""))
    return recs


def _records(kind, path):
    r = _dump("reads", kind, path)
    assert r.returncode == 0, r.stderr
    out = []
    for l in r.stdout.split(b"\n")[:-1]:
        a, s, q = l.split(b"\t")
        out.append((a, s, q))
    return out


def test_read_parsers_follow_the_reference_line_rules(tmp_path):
    rng = np.random.default_rng(93)
    db = H.make_db(rng, 50)
    batch = H.make_reads(rng, db, 60, ragged=True)
    d = str(tmp_path)
    recs = []
    for r in range(batch.n):
        a, b = int(batch.off[r]), int(batch.off[r + 1])
        recs.append((b"@r%d extra words" % r, batch.seq[a:b].tobytes(), batch.qual[a:b].tobytes().replace(b"\n", b"I")))
    # FASTQ with CRLF, blank lines in odd places, a longer-than-sequence quality line, no final newline
    fq = b""
    for i, (a, s, q) in enumerate(recs):
        q = bytes(c if c not in (9, 10, 11, 12, 13, 32) else 73 for c in q)
        eol = b"\r\n" if i % 3 == 0 else b"\n"
        fq += a + eol + (b"\n" if i % 5 == 0 else b"") + s + eol + b"+" + eol + q + (b"IIII" if i % 7 == 0 else b"") + eol
    fq += b"@last\nACGT\n+\nIIII"
    with gzip.open(os.path.join(d, "a.fastq.gz"), "wb") as f:
        f.write(fq)
    assert _records("gzfastq", os.path.join(d, "a.fastq.gz")) == _expect_fastq(fq, plain=False)
    # plain FASTQ: a blank line would repeat the previous token and derail the 4-line state machine
    # (then the reference dies in qual.at()), so the plain file carries none
    fqp = fq.replace(b"\n\n", b"\n").replace(b"\r\n\n", b"\r\n")
    open(os.path.join(d, "a.fastq"), "wb").write(fqp)
    assert _records("fastq", os.path.join(d, "a.fastq")) == _expect_fastq(fqp, plain=True)
    assert len(_expect_fastq(fqp, plain=True)) == len(recs) + 1  # incl. the unterminated last record
    # FASTA: wrapped lines, CRLF, empty lines, a short record (dropped), no final newline
    fa = b""
    for i, (a, s, q) in enumerate(recs):
        eol = b"\r\n" if i % 2 else b"\n"
        fa += b">" + a[1:] + eol + b"".join(s[j:j + 37] + eol for j in range(0, len(s), 37)) + (b"\n" if i % 4 == 0 else b"")
    fa += b">tail no newline\n" + b"ACGT" * 20
    with gzip.open(os.path.join(d, "a.fasta.gz"), "wb") as f:
        f.write(fa)
    open(os.path.join(d, "a.fasta"), "wb").write(fa)
    assert _records("gzfasta", os.path.join(d, "a.fasta.gz")) == _expect_fasta(fa, plain=False)
    assert _records("fasta", os.path.join(d, "a.fasta")) == _expect_fasta(fa, plain=True)
    r = _dump("reads", "fasta", os.path.join(d, "missing.fasta"))
    assert b"OPEN_FAILED" in r.stdout


def test_tree_loader_rules(tmp_path):
    p = os.path.join(str(tmp_path), "t.txt")
    open(p, "wb").write(b"2 3\r\n3 4\n\n5 6 junk\nx y\n")
    r = _dump("tree", p, 10)
    parent = [int(x) for x in r.stdout.split()]
    # blank line / unparsable line: i becomes 0 and j keeps its last value (istream rules, :979)
    assert parent[3] == 2 and parent[4] == 3 and parent[6] == 0
    open(p, "wb").write(b"2 30\n")
    assert _dump("tree", p, 10).stdout.startswith(b"ERR")
    assert [int(x) for x in _dump("tree", os.path.join(str(tmp_path), "none.txt"), 5).stdout.split()] == [1] * 5


def test_reader_packed_batches_match_text_batches_and_oracle_trim(tmp_path):
    """The hosts' batches (BatchMode::Packed: trimmed, 2-bit packed by kid_pack_reads per record) carry
    exactly what the text batches + process_qual would give, across batch boundaries (7 reads / 4 KiB
    per batch) and for a FASTA record that is longer than a whole batch."""
    rng = np.random.default_rng(94)
    db = H.make_db(rng, 50)
    batch = H.make_reads(rng, db, 80, ragged=True, n_rate=0.01, lower_rate=0.05)
    d = str(tmp_path)
    fq = b""
    for r in range(batch.n):
        a, b = int(batch.off[r]), int(batch.off[r + 1])
        q = bytes(c if c not in (9, 10, 11, 12, 13, 32) else 73 for c in batch.qual[a:b].tobytes())
        # CRLF, blank and "\r"-only lines in odd places (they do not advance the 4-line state), quality
        # lines longer than their read
        eol = b"\r\n" if r % 3 == 0 else b"\n"
        fq += b"@r%d" % r + eol + (b"\n" if r % 5 == 0 else b"") + batch.seq[a:b].tobytes() + eol + (b"\r\n" if r % 4 == 1 else b"")
        fq += b"+" + eol + q + (b"IIII" if r % 7 == 0 else b"") + eol + (b"\n\n" if r % 11 == 0 else b"")
    fq += b"@dropped no final newline\nACGTACGTACGTACGTACGTACGTACGTACGTACGT\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII"
    with gzip.open(os.path.join(d, "p.fastq.gz"), "wb") as f:
        f.write(fq)
    text = _records("gzfastq", os.path.join(d, "p.fastq.gz"))
    r = _dump("packed", "gzfastq", os.path.join(d, "p.fastq.gz"), 0)
    assert r.returncode == 0, r.stderr
    rows = [l.split(b"\t") for l in r.stdout.split(b"\n") if l]
    assert len(rows) == len(text) == batch.n
    lut = np.full(256, ord("N"), np.uint8)
    for ch in b"ACGT":
        lut[ch] = ch
        lut[ch | 0x20] = ch
    for (acc, seq, qual), row in zip(text, rows):
        st, sp = kor.trim(qual, len(seq))
        assert row[0] == acc and (int(row[1]), int(row[2])) == (st, sp)
        if sp - st < 30:
            assert int(row[3]) == 0 and (len(row) < 6 or row[5] == b"")
            continue
        want = lut[np.frombuffer(seq[st:sp + 1], np.uint8)].tobytes()
        assert int(row[3]) == sp - st + 1 and row[5] == want
        assert int(row[4]) == int(b"N" in want)
    # FASTA through the same path: whole records, U accepted with flag 1
    big = H._BASES[rng.integers(0, 4, size=30000)].tobytes()  # > 4096 + 16384 bytes: the batch buffer grows for it
    fa = b">x\n" + b"ACGU" * 20 + b"\n>y\nACGTNNACGT" + b"ACGT" * 10 + b"\n>big\n"
    fa += b"".join(big[j:j + 60] + b"\n" for j in range(0, len(big), 60)) + b">z\n" + b"ACGT" * 10 + b"\n"
    open(os.path.join(d, "p.fasta"), "wb").write(fa)
    r0 = [l.split(b"\t") for l in _dump("packed", "fasta", os.path.join(d, "p.fasta"), 0).stdout.split(b"\n") if l]
    r1 = [l.split(b"\t") for l in _dump("packed", "fasta", os.path.join(d, "p.fasta"), 1).stdout.split(b"\n") if l]
    assert r0[0][5] == b"ACGN" * 20 and r1[0][5] == b"ACGT" * 20 and r0[1][5] == b"ACGTNNACGT" + b"ACGT" * 10
    assert (int(r0[0][4]), int(r1[0][4]), int(r0[1][4])) == (1, 0, 1)
    assert r0[2][0] == b"big" and r0[2][5] == big and int(r0[2][3]) == 30000 and r0[3][5] == b"ACGT" * 10


def test_reader_fuzz_parallel_framing_against_serial_text_reader(tmp_path):
    """Random FASTQ files full of the reference's parser quirks (CRLF, blank and "\\r"-only lines anywhere,
    '+' lines with text, long quality lines, unterminated last record, several gzip members, 0 to 600
    records of 1 to 1000 bases) through the hosts' reader with 0, 1 and 3 framing threads and 4 KiB
    super-blocks: record for record what the serial text reader + the oracle's trim give."""
    import random
    rnd = random.Random(20261018)
    lut = np.full(256, ord("N"), np.uint8)
    for ch in b"ACGT":
        lut[ch] = ch
        lut[ch | 0x20] = ch
    p = os.path.join(str(tmp_path), "fz.fastq.gz")
    for it in range(36):
        n = rnd.choice([0, 1, 2, 3, 5, 17, 60, 200, 600])
        fq = b""
        for r in range(n):
            L = rnd.choice([1, 2, 29, 30, 31, 32, 50, 100, 150, 151, 300, 1000])
            s = bytes(rnd.choice(b"ACGTACGTACGTNacgtn") for _ in range(L))
            q = bytes(rnd.choice(b"#+05?I!~") for _ in range(L + rnd.choice([0, 0, 0, 3])))
            eol = rnd.choice([b"\n", b"\r\n"])
            blank = lambda: rnd.choice([b"", b"", b"", b"\n", b"\r\n", b"\n\n"])
            fq += b"@r%d x" % r + eol + blank() + s + eol + blank() + rnd.choice([b"+", b"+r%d" % r]) + eol + blank() + q + eol + blank()
        if rnd.random() < 0.3:
            fq += b"@tail\nACGTACGTACGTACGTACGTACGTACGTACGTAC"  # unterminated: dropped
        members = rnd.choice([1, 1, 3])
        with open(p, "wb") as f:
            step = max(1, (len(fq) + members - 1) // members)
            for i in range(0, max(len(fq), 1), step):
                f.write(gzip.compress(fq[i:i + step], 1))
        text = _records("gzfastq", p)
        for th in ("0", "1", "3"):
            env = dict(os.environ, KID_PARSE_THREADS=th, KID_GZ_MIN_BYTES="0",
                       KID_GZ_PIECE_BYTES=str(rnd.choice([2048, 8192, 1 << 20])))
            r = subprocess.run([DUMP, "packed", "gzfastq", p, "0"], capture_output=True, timeout=120, env=env)
            assert r.returncode == 0, r.stderr
            rows = [l.split(b"\t") for l in r.stdout.split(b"\n") if l]
            assert len(rows) == len(text), (it, th)
            for (acc, seq, qual), row in zip(text, rows):
                st, sp = kor.trim(qual, len(seq))
                assert row[0] == acc and (int(row[1]), int(row[2])) == (st, sp), (it, th, acc)
                if sp - st < 30:
                    assert int(row[3]) == 0
                    continue
                want = lut[np.frombuffer(seq[st:sp + 1], np.uint8)].tobytes()
                assert int(row[3]) == sp - st + 1 and row[5] == want and int(row[4]) == int(b"N" in want), (it, th, acc)
