#!/bin/bash
# round-2 GPU check P: what the driver runs at round end - GPU suite, smoke(), default bench line
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputests_p.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_p.log
tail -n 8 gpurun_out/gputests_p.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -n 5
( time timeout 900 python bench.py ) > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; echo "bench rc=$?"; tail -n 4 gpurun_out/bench_p.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_p.json").read().strip().splitlines()[-1])
r=d["roofline"]; print("value %.1fM e2e %.1fM kernel_ms %.3f frac %.3f launches %d" % (d["value"]/1e6, d["e2e"]["value"]/1e6, r["kernel_ms"], r["frac"], d["gpu_launches"]))
P
