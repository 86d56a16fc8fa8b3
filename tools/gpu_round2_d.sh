#!/bin/bash
# Round-2 GPU batch D: K4 after the in-warp softmax / pooling tail for states that fit a warp (cfg3): value tests, full-size
# parity of the other configs, K4 timings, phase traces of the three shapes.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_value_tc.py tests/test_gpu_parity.py -m gpu -q -x ) > $O/d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/d_pytest.log
for wl in cfg2 cfg3 cfg4; do
  WORKLOAD=$wl timeout 300 python tools/k4_only.py tc_fp16x2 5 > $O/d_k4_$wl.txt 2>&1
  WORKLOAD=$wl timeout 300 python tools/trace_tc.py tc_fp16x2 > $O/d_trace_$wl.txt 2>&1
done
