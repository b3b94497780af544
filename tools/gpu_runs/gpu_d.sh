#!/bin/bash
# round-2 GPU check D (2 GPUs): multi-GPU nk10 drop-in tests, torchrun bench at N=2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo2.txt 2>&1
timeout 1200 python -m pytest tests/test_nk10_dropin.py tests/test_golden_ref_case.py tests/test_multi_gpu.py -m gpu -q --tb=short -x > gpurun_out/gputests_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_d.log
tail -n 15 gpurun_out/gputests_d.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_d_n2.json 2> gpurun_out/bench_d_n2.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_d_n2.json").read().strip().splitlines()[-1])
print("N=2 value %.1fM e2e %.1fM e2e_text %.1fM" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e_text"]["value"]/1e6), d.get("parity_checked_reads"), d.get("files_e2e"))
P
tail -n 5 gpurun_out/bench_d_n2.err
