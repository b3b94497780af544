#!/bin/bash
# round-2 GPU check K: final default bench line + reference arm, ncu launch list, ncu captures (bact10 and 10x), file path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_packed.py tests/test_gpu_parity.py tests/test_nk10_dropin.py -m gpu -q --tb=short -x > gpurun_out/gputests_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_k.log
tail -n 3 gpurun_out/gputests_k.log
( time timeout 900 python bench.py ) > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_k.json").read().strip().splitlines()[-1])
r=d["roofline"]; print("value %.1fM e2e %.1fM e2e_text %.1fM kernel_ms %.3f pack_ms %.3f G lookups/s %.1f frac %.3f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e_text"]["value"]/1e6, r["kernel_ms"], r["pack_kernel_ms"], r["lookups_per_s"]/1e9, r["frac"]))
print(d["files_e2e"]); print(d["cpu_baseline"])
P
( time timeout 900 python bench.py --impl reference ) > gpurun_out/bench_k_ref.json 2> gpurun_out/bench_k_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_k_ref.json | cut -c1-1200; tail -n 4 gpurun_out/bench_k_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-files-e2e > gpurun_out/ncu_k_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kid_classify3 --launch-skip 3 -c 1 -f -o gpurun_out/prof_r2_v4 python bench.py --pairs 1000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-files-e2e > gpurun_out/ncu_k.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kid_classify3 --launch-skip 3 -c 1 -f -o gpurun_out/prof_r2_x10 python bench.py --config x10 --pairs 1000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_k_x10.log 2>&1
ls -la gpurun_out/prof_r2_v4.ncu-rep gpurun_out/prof_r2_x10.ncu-rep gpurun_out/r2_launches.csv
