#!/bin/bash
# round-2 GPU check A: host facts, GPU tests, bench (new packed path vs round-1 fused kernel), ncu capture
mkdir -p gpurun_out
{ lscpu | head -25; nvidia-smi topo -m; free -g; df -h /tmp; nvidia-smi --query-gpu=name,memory.total --format=csv; } > gpurun_out/host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/gputests_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_a.log
timeout 900 python bench.py > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "rc=$?" >> gpurun_out/bench_a.err
KID_FUSED_TEXT_KERNEL=1 timeout 400 python bench.py --no-e2e --no-cpu-baseline --no-files-e2e > gpurun_out/bench_a_fused.json 2> gpurun_out/bench_a_fused.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kid_classify3 --launch-skip 3 -c 1 -f -o gpurun_out/prof_r2_v1 python bench.py --pairs 1000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-files-e2e > gpurun_out/ncu_a.log 2>&1
tail -5 gpurun_out/gputests_a.log; head -c 1500 gpurun_out/bench_a.json; tail -3 gpurun_out/bench_a.err
