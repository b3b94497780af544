#!/bin/bash
# round-2 GPU check C: whole GPU suite with the async packed hosts, group-size variants, e2e chunk/slot sweep, ncu
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/gputests_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_c.log
tail -n 3 gpurun_out/gputests_c.log
show() { python - "$1" "$2" <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d["roofline"]; e=d.get("e2e") or {}
print("%s value %.1fM e2e %.1fM (%.2f ms) kernel_ms %.3f pack_ms %.3f frac %.3f" % (sys.argv[2], d["value"]/1e6, e.get("value",0)/1e6, e.get("ms_per_step",0), r["kernel_ms"], r["pack_kernel_ms"], r["frac"]))
P
}
for t in 0 4 5 6; do
  KID_TUNE=$t timeout 300 python bench.py --no-cpu-baseline --no-files-e2e > gpurun_out/bench_c_t$t.json 2> gpurun_out/bench_c_t$t.err; show gpurun_out/bench_c_t$t.json "tune $t"
done
for cfg in "2 262144" "4 262144" "3 131072" "3 524288" "4 1048576"; do
  set -- $cfg
  KID_HOST_SLOTS=$1 timeout 300 python bench.py --no-cpu-baseline --no-files-e2e --chunk-reads $2 > gpurun_out/bench_c_s$1_$2.json 2> gpurun_out/bench_c_s$1_$2.err; show gpurun_out/bench_c_s$1_$2.json "slots $1 chunk $2"
done
timeout 900 python bench.py > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; show gpurun_out/bench_c.json "default full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kid_classify3 --launch-skip 3 -c 1 -f -o gpurun_out/prof_r2_v2 python bench.py --pairs 1000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-files-e2e > gpurun_out/ncu_c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kid_pack_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_r2_pack python bench.py --pairs 1000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-files-e2e > gpurun_out/ncu_c2.log 2>&1
