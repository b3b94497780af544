#!/bin/bash
# round-2 GPU check U: ucount histogram in shared memory vs global atomics
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_packed.py -m gpu -x -q > gpurun_out/gputests_u.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_u.log
tail -n 3 gpurun_out/gputests_u.log
for g in 0 1 0 1; do
  if [ $g = 1 ]; then export KID_UCOUNT_GLOBAL=1; else unset KID_UCOUNT_GLOBAL; fi
  timeout 300 python bench.py --no-cpu-baseline --no-files-e2e --no-e2e > gpurun_out/bench_u_$g.json 2> gpurun_out/bench_u_$g.err
  python - <<P
import json
d=json.loads(open("gpurun_out/bench_u_$g.json").read().strip().splitlines()[-1])
r=d["roofline"]; print("global=$g value %.1fM ms_per_step %.3f kernel_ms %.3f pack_ms %.3f -> rest %.3f ms" % (d["value"]/1e6, d["ms_per_step"], r["kernel_ms"], r["pack_kernel_ms"], d["ms_per_step"]-r["kernel_ms"]-r["pack_kernel_ms"]))
P
done
unset KID_UCOUNT_GLOBAL
timeout 300 python bench.py --config mito --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('mito value %.1fM ms_per_step %.3f kernel %.3f pack %.3f' % (d['value']/1e6, d['ms_per_step'], r['kernel_ms'], r['pack_kernel_ms']))"
