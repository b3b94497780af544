#!/bin/bash
# round-2 GPU check F: config-scale parity (mito, x10) and bench lines for --config mito / x10
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests/test_gpu_configs_fullsize.py -m gpu -q --tb=short -x ) > gpurun_out/gputests_f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_f.log
tail -n 12 gpurun_out/gputests_f.log
timeout 600 python bench.py --config mito > gpurun_out/bench_f_mito.json 2> gpurun_out/bench_f_mito.err; echo "mito rc=$?"
( time timeout 1500 python bench.py --config x10 --cpu-baseline-reads 100000 ) > gpurun_out/bench_f_x10.json 2> gpurun_out/bench_f_x10.err; echo "x10 rc=$?"
python - <<'P'
import json
for c in ("mito","x10"):
    try:
        d=json.loads(open("gpurun_out/bench_f_%s.json"%c).read().strip().splitlines()[-1])
        r=d["roofline"]; print(c, "value %.1fM e2e %.1fM kernel_ms %.3f pack_ms %.3f G lookups/s %.1f frac %.3f displaced %d" % (d["value"]/1e6, d["e2e"]["value"]/1e6, r["kernel_ms"], r["pack_kernel_ms"], r["lookups_per_s"]/1e9, r["frac"], d["table"]["displaced"]), d.get("cpu_baseline",{}).get("value"))
    except Exception as e: print(c, "failed", e)
P
tail -n 4 gpurun_out/bench_f_x10.err
