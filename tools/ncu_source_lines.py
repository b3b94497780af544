"""Per-source-line executed instructions and stall samples from an .ncu-rep (needs -lineinfo).
   python tools/ncu_source_lines.py rep.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
import os
cur_file = ""
ci = {n: hdr.index(n) for n in ("Line No", "Source", "# Samples", "Instructions Executed")}

stall = {n: hdr.index(n) for n in hdr if n.startswith("stall_") and "Not Issued" not in n}
def num(x):
    try:
        return int(x)
    except ValueError:
        return 0
data = []
for r in rows:
    if r and r[0] == "File Path":
        cur_file = os.path.basename(r[1])
        continue
    if len(r) < len(hdr) or not r[0].isdigit():
        continue
    ex = num(r[ci["Instructions Executed"]])
    smp = num(r[ci["# Samples"]])
    st = {k[6:]: num(r[v]) for k, v in stall.items()}
    data.append((f"{cur_file}:{r[0]}", r[ci["Source"]].strip()[:90], ex, smp, st))
tot_ex = sum(d[2] for d in data); tot_s = sum(d[3] for d in data)
print(f"total executed warp-instructions {tot_ex}, samples {tot_s}")
print("--- by executed instructions")
for d in sorted(data, key=lambda d: -d[2])[:top]:
    print(f"{d[0]:>24s} {100*d[2]/tot_ex:5.1f}% ex {100*d[3]/max(1,tot_s):5.1f}% smp  {d[1]}")
print("--- by stall samples")
for d in sorted(data, key=lambda d: -d[3])[:top // 2]:
    top3 = sorted(d[4].items(), key=lambda kv: -kv[1])[:3]
    print(f"{d[0]:>24s} {100*d[3]/max(1,tot_s):5.1f}% smp  {top3}  {d[1]}")
