#!/bin/bash
# round-2 GPU check T2: the device-reader tests, the drop-in tests in every host mode, smoke()
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ingest.py tests/test_nk10_dropin.py tests/test_golden_ref_case.py -m gpu -q --tb=short -x > gpurun_out/gputests_t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_t2.log
tail -n 12 gpurun_out/gputests_t2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
