#!/bin/bash
# what the driver runs at round end: the GPU suite, smoke(), the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -m gpu -q --tb=short -x > gpurun_out/gputests_p2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_p2.log
tail -n 6 gpurun_out/gputests_p2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_p2.json 2> gpurun_out/bench_p2.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open("gpurun_out/bench_p2.json").read().strip().splitlines()[-1])
r=d["roofline"]; print("value %.1fM e2e %.1fM ms_per_step %.3f kernel_ms %.3f pack_ms %.3f frac %.3f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], r["kernel_ms"], r["pack_kernel_ms"], r["frac"]))
print(json.dumps(d["files_e2e"], indent=1))
P
