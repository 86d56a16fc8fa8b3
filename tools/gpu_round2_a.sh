#!/bin/bash
# Round-2 GPU batch A: parity suite, bench lines (cfg2 / cfg3 / cfg4 / cfg5 at N = 1), phase traces and ncu captures of the
# K4 entity kernel at the cfg3 / cfg4 shapes.  Everything lands in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/a_gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q --durations=15 ) > $O/a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/a_bench_cfg2.json 2> $O/a_bench_cfg2.err
timeout 600 python bench.py --workload cfg3 --steps 20 --warmup 3 --no-cpu-baseline > $O/a_bench_cfg3.json 2> $O/a_bench_cfg3.err
timeout 600 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline > $O/a_bench_cfg4.json 2> $O/a_bench_cfg4.err
timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 1 > $O/a_bench_cfg5.json 2> $O/a_bench_cfg5.err
for wl in cfg3 cfg4; do
  WORKLOAD=$wl timeout 300 python tools/trace_tc.py tc_fp16x2 > $O/a_trace_$wl.txt 2>&1
  WORKLOAD=$wl timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_entity -s 1 -c 1 \
      -o $O/a_ncu_$wl -f python tools/k4_only.py tc_fp16x2 3 > $O/a_ncu_$wl.log 2>&1
done
ls -la $O > $O/a_ls.txt
