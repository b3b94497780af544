"""Executed-instruction mix by SASS opcode from an .ncu-rep.  python tools/ncu_sass_mix.py rep"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and "Address" in r and "Source" in r)
hdr = rows[hi]
si, ei = hdr.index("Source"), hdr.index("Instructions Executed")
mix = collections.Counter(); tot = 0
for r in rows[hi + 1:]:
    if len(r) <= ei: continue
    try: ex = int(r[ei])
    except ValueError: continue
    toks = r[si].split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.split(".")[0]
    mix[op] += ex; tot += ex
print("total", tot)
for op, n in mix.most_common(30):
    print(f"{op:12s} {n:14d} {100*n/tot:5.1f}%")
