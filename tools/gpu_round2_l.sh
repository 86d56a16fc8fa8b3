#!/bin/bash
# Round-2 GPU batch L: K4 with the uniform-datapath MMA warp loop
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_value_tc.py tests/test_gpu_parity.py -m gpu -q -x ) > $O/l_pytest.log 2>&1; echo "pytest rc=$?" >> $O/l_pytest.log
for wl in cfg2 cfg3 cfg4; do
  WORKLOAD=$wl timeout 300 python tools/k4_only.py tc_fp16x2 5 > $O/l_k4_$wl.txt 2>&1
done
timeout 300 python tools/k4_only.py tc_bf16 5 > $O/l_k4_cfg2_bf16.txt 2>&1
timeout 300 python tools/k4_only.py tc_fp32 5 > $O/l_k4_cfg2_bf16x3.txt 2>&1
WORKLOAD=cfg2 timeout 300 python tools/trace_tc.py tc_fp16x2 > $O/l_trace_cfg2.txt 2>&1
