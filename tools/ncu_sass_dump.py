"""SASS listing with executed counts (per read) from an .ncu-rep, in address order.
   python tools/ncu_sass_dump.py rep n_reads [min_per_read]"""
import csv, subprocess, sys
rep = sys.argv[1]; n_reads = float(sys.argv[2]); thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and "Address" in r and "Source" in r)
hdr = rows[hi]
si, ei, ai = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Address")
sm = hdr.index("# Samples") if "# Samples" in hdr else None
for r in rows[hi + 1:]:
    if len(r) <= ei: continue
    try: ex = int(r[ei])
    except ValueError: continue
    per = ex / n_reads
    if per >= thr:
        smp = r[sm] if sm is not None else ""
        print(f"{r[ai][-5:]} {per:6.2f} {smp:>6s}  {r[si]}")
