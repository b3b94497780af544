#!/bin/bash
# round-2 GPU check L: A/B of the smaller-footprint kernel build against the committed one; file path after the buffer sizing fix
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_packed.py tests/test_gpu_parity.py -m gpu -q --tb=short -x > gpurun_out/gputests_l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_l.log
tail -n 3 gpurun_out/gputests_l.log
show() { python - "$1" "$2" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print("%s value %.1fM kernel_ms %.3f G lookups/s %.1f frac %.3f" % (sys.argv[2], d["value"]/1e6, r["kernel_ms"], r["lookups_per_s"]/1e9, r["frac"]), d.get("files_e2e") and d["files_e2e"]["sample_s"])
except Exception as ex: print(sys.argv[2], "failed", ex)
P
}
B=$GRAFT_REPO_ROOT/tools/micro/libkmerid_b200_B.so
timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_l_A.json 2> gpurun_out/bench_l_A.err; show gpurun_out/bench_l_A.json "A bact10"
KID_LIB_PATH=$B timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-files-e2e > gpurun_out/bench_l_B.json 2> gpurun_out/bench_l_B.err; show gpurun_out/bench_l_B.json "B bact10"
timeout 600 python bench.py --config x10 --no-cpu-baseline --no-e2e > gpurun_out/bench_l_A_x10.json 2> gpurun_out/bench_l_A_x10.err; show gpurun_out/bench_l_A_x10.json "A x10"
KID_LIB_PATH=$B timeout 600 python bench.py --config x10 --no-cpu-baseline --no-e2e > gpurun_out/bench_l_B_x10.json 2> gpurun_out/bench_l_B_x10.err; show gpurun_out/bench_l_B_x10.json "B x10"
timeout 300 python bench.py --config mito --no-cpu-baseline --no-e2e > gpurun_out/bench_l_A_mito.json 2> gpurun_out/bench_l_A_mito.err; show gpurun_out/bench_l_A_mito.json "A mito"
KID_LIB_PATH=$B timeout 300 python bench.py --config mito --no-cpu-baseline --no-e2e > gpurun_out/bench_l_B_mito.json 2> gpurun_out/bench_l_B_mito.err; show gpurun_out/bench_l_B_mito.json "B mito"
