#!/bin/bash
# Round-2 closing evidence on the final tree: ncu launch list of the bench command + ncu --set full of the K4 entity kernel (cfg2).
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustained-seconds 0 > $O/f_plain.log 2>&1 &&
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/f_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustained-seconds 0 > $O/f_ncu_launches.log 2>&1
WORKLOAD=cfg2 timeout 200 python tools/k4_only.py tc_fp16x2 3 > $O/f_k4_cfg2.txt 2>&1 &&
  WORKLOAD=cfg2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_entity -s 1 -c 1 \
      -o $O/f_ncu_k4_cfg2 -f python tools/k4_only.py tc_fp16x2 3 > $O/f_ncu_k4_cfg2.log 2>&1
cat $O/f_k4_cfg2.txt | tail -3
ls -la $O | tail -8
