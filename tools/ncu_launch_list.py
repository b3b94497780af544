"""Launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X) -> text summary for profiles/.
   python tools/ncu_launch_list.py gpurun_out/launches.csv "<command that was profiled>" """
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = []
for r in rows:
    if r is hdr or r[0] == "ID":
        continue
    v = float(r[vi].replace(",", ""))
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
    launches.append((r[ki][:70], v * scale))  # microseconds
tot = sum(v for _, v in launches)
agg = collections.OrderedDict()
for k, v in launches:
    a = agg.setdefault(k, [0.0, 0])
    a[0] += v
    a[1] += 1
print("# ncu --metrics gpu__time_duration.sum --clock-control none  " + (sys.argv[2] if len(sys.argv) > 2 else ""))
print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
print("# total %.3f ms over %d launches\n" % (tot / 1e3, len(launches)))
for k, (v, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%10.3f ms  %5.1f%%  x%-4d %s" % (v / 1e3, 100 * v / tot, n, k))
print("\n# launch sequence (us)")
for k, v in launches:
    print("%12.1f  %s" % (v, k))
