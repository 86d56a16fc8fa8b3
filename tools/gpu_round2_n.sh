#!/bin/bash
# Round-2 GPU batch N: attention-weight output of K4 (parity tests) and K4 alone on the three shapes (the output is off: no cost)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_attention_weights.py tests/test_gpu_value_tc.py -m gpu -q -x ) > $O/n_pytest.log 2>&1; echo "pytest rc=$?" >> $O/n_pytest.log
for wl in cfg2 cfg3 cfg4; do
  WORKLOAD=$wl timeout 300 python tools/k4_only.py tc_fp16x2 5 > $O/n_k4_$wl.txt 2>&1
done
tail -n 6 $O/n_pytest.log; cat $O/n_k4_*.txt
