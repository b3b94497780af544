#!/bin/bash
# round-2 GPU check M: one-multiply minimizer hash (build B), first halves bypassing L1 (KID_TUNE=4)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_packed.py -m gpu -q --tb=short -x > gpurun_out/gputests_m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_m.log
tail -n 3 gpurun_out/gputests_m.log
show() { python - "$1" "$2" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print("%s value %.1fM kernel_ms %.3f G lookups/s %.1f frac %.3f displaced %d" % (sys.argv[2], d["value"]/1e6, r["kernel_ms"], r["lookups_per_s"]/1e9, r["frac"], d["table"]["displaced"]))
except Exception as ex: print(sys.argv[2], "failed", ex)
P
}
B=$GRAFT_REPO_ROOT/tools/micro/libkmerid_b200_B.so
for rep in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-files-e2e > gpurun_out/bench_m_A.json 2> gpurun_out/bench_m_A.err; show gpurun_out/bench_m_A.json "A bact10"
KID_LIB_PATH=$B timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-files-e2e > gpurun_out/bench_m_B.json 2> gpurun_out/bench_m_B.err; show gpurun_out/bench_m_B.json "B(short hash) bact10"
KID_TUNE=4 timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-files-e2e > gpurun_out/bench_m_A4.json 2> gpurun_out/bench_m_A4.err; show gpurun_out/bench_m_A4.json "A noL1 bact10"
done
KID_LIB_PATH=$B timeout 300 python -m pytest tests/test_gpu_packed.py -m gpu -q --tb=short -x 2>&1 | tail -n 1
KID_LIB_PATH=$B timeout 600 python bench.py --config x10 --no-cpu-baseline --no-e2e > gpurun_out/bench_m_B_x10.json 2> gpurun_out/bench_m_B_x10.err; show gpurun_out/bench_m_B_x10.json "B x10"
timeout 600 python bench.py --config x10 --no-cpu-baseline --no-e2e > gpurun_out/bench_m_A_x10.json 2> gpurun_out/bench_m_A_x10.err; show gpurun_out/bench_m_A_x10.json "A x10"
KID_LIB_PATH=$B timeout 300 python bench.py --config mito --no-cpu-baseline --no-e2e > gpurun_out/bench_m_B_mito.json 2> gpurun_out/bench_m_B_mito.err; show gpurun_out/bench_m_B_mito.json "B mito"
timeout 300 python bench.py --config mito --no-cpu-baseline --no-e2e > gpurun_out/bench_m_A_mito.json 2> gpurun_out/bench_m_A_mito.err; show gpurun_out/bench_m_A_mito.json "A mito"
