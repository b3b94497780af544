#!/bin/bash
# round-2 GPU check R: hosts after the warm-up change (drop-in tests, file-level timing)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nk10_dropin.py tests/test_kmerread_dropin.py tests/test_kmerreadc_dropin.py tests/test_golden_ref_case.py -m gpu -x -q > gpurun_out/gputests_r.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_r.log
tail -n 3 gpurun_out/gputests_r.log
timeout 600 python bench.py --no-e2e --no-cpu-baseline --steps 3 > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_r.json").read().strip().splitlines()[-1])
print(d["files_e2e"])
P
