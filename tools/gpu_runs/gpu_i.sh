#!/bin/bash
# round-2 GPU check I: vote / range-mask variants; 20-mer minimizers on the 10x database
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_packed.py tests/test_gpu_parity.py -m gpu -q --tb=short -x > gpurun_out/gputests_i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_i.log
tail -n 3 gpurun_out/gputests_i.log
show() { python - "$1" "$2" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]; e=d.get("e2e") or {}
    print("%s value %.1fM kernel_ms %.3f G lookups/s %.1f frac %.3f displaced %d" % (sys.argv[2], d["value"]/1e6, r["kernel_ms"], r["lookups_per_s"]/1e9, r["frac"], d["table"]["displaced"]))
except Exception as ex: print(sys.argv[2], "failed", ex)
P
}
for t in 0 2 3 4; do
  KID_TUNE=$t timeout 300 python bench.py --no-cpu-baseline --no-files-e2e --no-e2e > gpurun_out/bench_i_t$t.json 2> gpurun_out/bench_i_t$t.err; show gpurun_out/bench_i_t$t.json "tune $t"
done
for mm in 16 20; do
  KID_DB_MM=$mm timeout 600 python bench.py --config x10 --no-cpu-baseline --no-e2e > gpurun_out/bench_i_x10_mm$mm.json 2> gpurun_out/bench_i_x10_mm$mm.err; show gpurun_out/bench_i_x10_mm$mm.json "x10 mm=$mm"
done
KID_DB_MM=20 KID_TUNE=1 timeout 600 python bench.py --config x10 --no-cpu-baseline --no-e2e > gpurun_out/bench_i_x10_mm20_t1.json 2> gpurun_out/bench_i_x10_mm20_t1.err; show gpurun_out/bench_i_x10_mm20_t1.json "x10 mm=20 2+2"
KID_DB_MM=20 timeout 300 python bench.py --no-cpu-baseline --no-files-e2e --no-e2e > gpurun_out/bench_i_mm20.json 2> gpurun_out/bench_i_mm20.err; show gpurun_out/bench_i_mm20.json "bact10 mm=20"
