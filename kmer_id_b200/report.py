"""Report step: `<sample>_result.txt` files -> abundance CSV.

Host-side mirror of the reference's two report scripts; the reference's are
Python, so this is Python too.  Not part of the GPU hot path: it consumes the files the drop-in
binaries (kmer_id_b200/bin/nk10, kmerread, kmerreadc) write.

* style "b10": readbatch_10.py - 4-column refkey (target, name, probe count, use), a fixed
  exclusion list, percentage = count / (probes + 10), normalised per sample (:91-108).
* style "c3":  readbatch_c3.py - 7-column refkey, factor = tested / hit / (size / strains) (:30-46),
  numpy arithmetic (so the floating-point results are bit-identical to the script's).

Both keep the reference's conventions: a row of a result file is matched to the i-th *in-use* target
by position, not by id (readbatch_10.py:82-86); counts below the thresholds are zeroed but still
count towards `total` (:78-81); targets whose best percentage is 0 are not printed (:133).
tests/test_report.py runs the unmodified scripts beside this module and compares the CSV bytes.
"""
from __future__ import annotations

import argparse
import os
from dataclasses import dataclass, field

import numpy as np

RESULT_EXT = "_result.txt"

# readbatch_10.py:15-21
_B10_EXCLUDE = frozenset([4178, 1744, 2539, 5624, 1575, 5647, 323, 2728, 268, 5317, 297, 3867, 314, 1344,
                          2947, 2935, 4213, 4976, 2767, 2763, 118, 3390, 1757]) | frozenset(range(1928, 2339))


@dataclass
class Style:
    mincount: float
    minuniq: float
    maxrat: float
    exclude: frozenset = field(default_factory=frozenset)


STYLES = {"b10": Style(2.0, 3.0, 80000.0, _B10_EXCLUDE),  # readbatch_10.py:11-13
          "c3": Style(2.0, 2.0, 80.0)}                    # readbatch_c3.py:12-14


def _lines(path):
    with open(path, "r") as f:
        return [ln for ln in f.read().split("\n") if len(ln) > 1]


def _load_refkey_b10(path, style):
    """-> (in_use per refkey row, names of in-use targets, divisor per in-use target) (:28-45)"""
    in_use, names, weight = [], [], []
    with open(path, "r") as f:
        raw = f.read().split("\n")[1:]  # header dropped unconditionally
    for ln in raw:
        if len(ln) <= 1:
            continue
        target, name, count, use = ln.split("\t")
        if int(target) in style.exclude:
            use = "0"
        in_use.append(int(use))
        if use == "1":
            weight.append(float(count) + 10.0)
            names.append(name)
    return in_use, names, weight


def _load_refkey_c3(path, style):
    in_use, names, factor = [], [], []
    with open(path, "r") as f:
        raw = f.read().split("\n")[1:]
    for ln in raw:
        if len(ln) <= 1:
            continue
        target, name, count, hit, tested, gsize, nstrains = ln.split("\t")
        gensize = float(gsize) / float(nstrains) if nstrains != "0" else 1.0
        ok = not (int(target) in style.exclude or float(count) < 10.0 or float(hit) < 10.0
                  or len(name.split("_")) < 6)
        in_use.append(1 if ok else 0)
        if ok:
            names.append(name)
            factor.append(float(tested) / float(hit) / gensize)
    return in_use, names, factor


def _read_result(path, style, in_use, n_targets, need_uniq):
    """One result file -> (column of filtered counts, total reads, list of no-id counts)."""
    col = [0] * n_targets
    total = 0.0
    noid = []
    index = 0
    for ln in _lines(path):
        row = ln.split(",")
        target = int(row[0])
        count = float(row[1])
        uniq = float(row[2]) if (need_uniq or len(row) > 2) else count
        kept = count
        if kept < style.mincount or uniq < style.minuniq or kept / uniq > style.maxrat:
            kept = 0.0
        total += count
        if target > 0:
            if in_use[target] == 1:
                col[index] = kept
                index += 1
        else:
            noid.append(int(count))
    return col, total, noid


def build_report(result_dir: str, refkey: str, style_name: str = "b10"):
    """Returns (CSV text the reference script would write for result_dir, n samples, n in-use targets)."""
    style = STYLES[style_name]
    b10 = style_name == "b10"
    in_use, names, weight = (_load_refkey_b10 if b10 else _load_refkey_c3)(refkey, style)
    n_t = len(names)

    files = [f for f in os.listdir(result_dir)
             if os.path.isfile(os.path.join(result_dir, f)) and f.endswith(RESULT_EXT)]
    samples, cols, totals, noids = [], [], [], []
    for f in files:
        # readbatch_10.py cuts at the FIRST "_result.txt" in the name (:60-62); c3 cuts the last 11 chars
        samples.append(f[:f.find(RESULT_EXT)] if b10 else f[:-len(RESULT_EXT)])
        col, total, noid = _read_result(os.path.join(result_dir, f), style, in_use, n_t, need_uniq=not b10)
        cols.append(col)
        totals.append(total)
        noids.extend(noid)
    n_c = len(files)

    if b10:
        # sequential sums in row order, exactly as :96-108, so the doubles agree to the last bit
        pct = [[0] * n_c for _ in range(n_t)]
        rowmax = [0] * n_t
        for c in range(n_c):
            s = 0
            for r in range(n_t):
                pct[r][c] = cols[c][r] / weight[r]
                s += pct[r][c]
            if s < 0.00000009:
                s = 0.0000001
            for r in range(n_t):
                pct[r][c] = pct[r][c] * 100.0 / s
                rowmax[r] = max(rowmax[r], pct[r][c])
        cell = lambda r, c: (cols[c][r], pct[r][c])
    else:
        m = np.zeros((n_t, n_c))
        for c in range(n_c):
            m[:, c] = cols[c]
        b = m * np.array(weight)[:, None]
        sums = np.sum(b, axis=0)
        for c in range(n_c):
            if sums[c] < 0.00000009:
                sums[c] = 0.0000001
        b = b / sums[None, :]
        b = b * 100.0
        rowmax = b.max(axis=1) if n_c else np.zeros(n_t)
        cell = lambda r, c: (m[r, c], b[r, c])

    order = sorted(range(n_c), key=lambda k: samples[k])
    out = ["name," + "".join(samples[k] + ",," for k in order),
           "total," + "".join(str(totals[k]) + ",," for k in order),
           "no_id," + "".join(str(noids[k]) + ",," for k in order)]
    for r in range(n_t):
        if rowmax[r] > 0.000:
            line = names[r]
            for k in order:
                cnt, p = cell(r, k)
                line += "," + str(cnt) + "," + str(p)
            out.append(line)
    return "\n".join(out) + "\n", n_c, n_t


def main(argv=None):
    ap = argparse.ArgumentParser(description="abundance CSV from *_result.txt files")
    ap.add_argument("--dir", required=True, help="directory holding <sample>_result.txt files")
    ap.add_argument("--refkey", required=True, help="refkey10.txt (b10) or refKeyc3.txt (c3)")
    ap.add_argument("--style", choices=sorted(STYLES), default="b10")
    ap.add_argument("--out", required=True)
    a = ap.parse_args(argv)
    text, n_cols, n_targets = build_report(a.dir, a.refkey, a.style)
    with open(a.out, "w") as f:
        f.write(text)
    print(n_cols, " x ", n_targets)


if __name__ == "__main__":
    main()
