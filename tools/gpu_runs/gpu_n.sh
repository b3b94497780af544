#!/bin/bash
# round-2 GPU check N: dense batches (tests, bench e2e, hosts)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/gputests_n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_n.log
tail -n 4 gpurun_out/gputests_n.log
timeout 900 python bench.py > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench_n.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_n.json").read().strip().splitlines()[-1])
r=d["roofline"]; print("value %.1fM kernel_ms %.3f frac %.3f" % (d["value"]/1e6, r["kernel_ms"], r["frac"]))
for k in ("e2e","e2e_packed","e2e_text","host_pack","files_e2e"): print(k, d[k])
P
