#!/bin/bash
# round-2 GPU check B: lookup variants of kid_classify3_kernel (TUNE build), new pack kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_packed.py tests/test_gpu_parity.py -q --tb=short -x > gpurun_out/gputests_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_b.log
for t in 2 3; do KID_TUNE=$t timeout 600 python -m pytest tests/test_gpu_packed.py -q --tb=short -x > gpurun_out/gputests_b_t$t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_b_t$t.log; done
for t in 0 1 2 3; do
  KID_TUNE=$t timeout 300 python bench.py --no-cpu-baseline --no-files-e2e > gpurun_out/bench_b_t$t.json 2> gpurun_out/bench_b_t$t.err
  python - <<P
import json
d=json.loads(open("gpurun_out/bench_b_t$t.json").read().strip().splitlines()[-1])
r=d["roofline"]; print("tune $t value %.1fM e2e %.1fM kernel_ms %.3f pack_ms %.3f frac %.3f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, r["kernel_ms"], r["pack_kernel_ms"], r["frac"]))
P
done
tail -3 gpurun_out/gputests_b*.log
