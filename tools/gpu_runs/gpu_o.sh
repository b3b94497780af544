#!/bin/bash
# round-2 GPU check O (8 GPUs): the default bench line with dense batches on the wire
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 > gpurun_out/bench_o_n8.json 2> gpurun_out/bench_o_n8.err; echo "bench n8 rc=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_o_n8.json").read().strip().splitlines()[-1])
print("value %.1fM e2e %.1fM (h2d %.2f GB/step/rank, %.1f ms) e2e_packed %.1fM e2e_text %.1fM" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["h2d_bytes_per_step"]/1e9, d["e2e"]["ms_per_step"], d["e2e_packed"]["value"]/1e6, d["e2e_text"]["value"]/1e6), d.get("parity_checked_reads"), d["files_e2e"]["sample_s"], d["files_e2e"]["value"])
P
tail -n 2 gpurun_out/bench_o_n8.err
