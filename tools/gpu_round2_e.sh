#!/bin/bash
# Round-2 GPU batch E: K4 with the full-warp transpose column sums for states of 32 rows or more (cfg4)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_value_tc.py tests/test_gpu_parity.py -m gpu -q -x ) > $O/e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/e_pytest.log
for wl in cfg4 cfg3; do
  WORKLOAD=$wl timeout 300 python tools/k4_only.py tc_fp16x2 5 > $O/e_k4_$wl.txt 2>&1
  WORKLOAD=$wl timeout 300 python tools/trace_tc.py tc_fp16x2 > $O/e_trace_$wl.txt 2>&1
done
