#!/bin/bash
# round-2 GPU check G: split-sector table + 4 chunks in flight
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_packed.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_configs_fullsize.py -m gpu -q --tb=short -x > gpurun_out/gputests_g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_g.log
tail -n 6 gpurun_out/gputests_g.log
show() { python - "$1" "$2" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]; e=d.get("e2e") or {}
    print("%s value %.1fM e2e %.1fM kernel_ms %.3f pack_ms %.3f G lookups/s %.1f frac %.3f" % (sys.argv[2], d["value"]/1e6, e.get("value",0)/1e6, r["kernel_ms"], r["pack_kernel_ms"], r["lookups_per_s"]/1e9, r["frac"]))
except Exception as ex: print(sys.argv[2], "failed", ex)
P
}
for t in 0 1 4 5; do
  KID_TUNE=$t timeout 300 python bench.py --no-cpu-baseline --no-files-e2e --no-e2e > gpurun_out/bench_g_t$t.json 2> gpurun_out/bench_g_t$t.err; show gpurun_out/bench_g_t$t.json "tune $t"
done
timeout 300 python bench.py --config mito --no-cpu-baseline --no-e2e > gpurun_out/bench_g_mito.json 2> gpurun_out/bench_g_mito.err; show gpurun_out/bench_g_mito.json mito
timeout 600 python bench.py --config x10 --no-cpu-baseline --no-e2e > gpurun_out/bench_g_x10.json 2> gpurun_out/bench_g_x10.err; show gpurun_out/bench_g_x10.json x10
KID_TUNE=1 timeout 600 python bench.py --config x10 --no-cpu-baseline --no-e2e > gpurun_out/bench_g_x10_t1.json 2> gpurun_out/bench_g_x10_t1.err; show gpurun_out/bench_g_x10_t1.json "x10 tune1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kid_classify3 --launch-skip 3 -c 1 -f -o gpurun_out/prof_r2_v3 python bench.py --pairs 1000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-files-e2e > gpurun_out/ncu_g.log 2>&1
