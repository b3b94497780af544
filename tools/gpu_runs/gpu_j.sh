#!/bin/bash
# round-2 GPU check J (8 GPUs): nk10 on 8 GPUs vs the reference, torchrun bench at N=8 (bact10 and the 10x database)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1; nproc >> gpurun_out/topo8.txt; free -g >> gpurun_out/topo8.txt
timeout 900 python -m pytest tests/test_nk10_dropin.py::test_parser_quirks_two_samples tests/test_multi_gpu.py -m gpu -q --tb=short -x > gpurun_out/gputests_j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_j.log
tail -n 4 gpurun_out/gputests_j.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 > gpurun_out/bench_j_n8.json 2> gpurun_out/bench_j_n8.err; echo "bench n8 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --config x10 > gpurun_out/bench_j_n8_x10.json 2> gpurun_out/bench_j_n8_x10.err; echo "bench n8 x10 rc=$?"
python - <<'P'
import json
for f in ("bench_j_n8","bench_j_n8_x10"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.1fM e2e %.1fM (h2d %.2f GB/step/rank) e2e_text %.1fM lookups/s %.1fG" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["h2d_bytes_per_step"]/1e9, d["e2e_text"]["value"]/1e6, d["lookups_per_s_whole_step"]/1e9), d.get("parity_checked_reads"), d.get("files_e2e",{}) and {k:d["files_e2e"].get(k) for k in ("value","n_gpus","mode","sample_s","host_cores")})
    except Exception as e: print(f, "failed", e)
P
tail -n 3 gpurun_out/bench_j_n8.err gpurun_out/bench_j_n8_x10.err
