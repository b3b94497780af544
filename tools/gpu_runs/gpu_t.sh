#!/bin/bash
# round-2 GPU check T: 20-mer minimizers with (order, identity) pairs on the 10x database
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_packed.py -m gpu -x -q > gpurun_out/gputests_t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_t.log
tail -n 3 gpurun_out/gputests_t.log
KID_DB_MM=20 timeout 900 python -m pytest tests/test_gpu_packed.py -m gpu -x -q 2>&1 | tail -n 1
show() { python - "$1" "$2" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print("%s value %.1fM kernel_ms %.3f G lookups/s %.1f frac %.3f displaced %d" % (sys.argv[2], d["value"]/1e6, r["kernel_ms"], r["lookups_per_s"]/1e9, r["frac"], d["table"]["displaced"]))
except Exception as ex: print(sys.argv[2], "failed", ex)
P
}
for mm in 20 16; do
  KID_DB_MM=$mm timeout 600 python bench.py --config x10 --no-cpu-baseline --no-e2e > gpurun_out/bench_t_x10_mm$mm.json 2> gpurun_out/bench_t_x10_mm$mm.err; show gpurun_out/bench_t_x10_mm$mm.json "x10 mm=$mm"
done
KID_DB_MM=20 timeout 300 python bench.py --no-cpu-baseline --no-files-e2e --no-e2e > gpurun_out/bench_t_mm20.json 2> gpurun_out/bench_t_mm20.err; show gpurun_out/bench_t_mm20.json "bact10 mm=20"
