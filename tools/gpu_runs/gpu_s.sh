#!/bin/bash
# round-2 GPU check S: final state - GPU suite, smoke(), default bench line, reference arm
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputests_s.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_s.log
tail -n 8 gpurun_out/gputests_s.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 900 python bench.py > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err; echo "bench rc=$?"; tail -n 2 gpurun_out/bench_s.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_s.json").read().strip().splitlines()[-1])
r=d["roofline"]; print("value %.1fM e2e %.1fM e2e_packed %.1fM e2e_text %.1fM kernel_ms %.3f frac %.3f launches %d files %.2fM" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e_packed"]["value"]/1e6, d["e2e_text"]["value"]/1e6, r["kernel_ms"], r["frac"], d["gpu_launches"], d["files_e2e"]["value"]/1e6))
P
