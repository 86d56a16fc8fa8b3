#!/bin/bash
# Round-2 GPU batch G: the evidence set of the round -- whole GPU suite, smoke, bench line + reference arm, ncu launch list of
# the bench command, ncu --set full of the K4 entity kernel at the three shapes and of the fused ORCA+step kernel.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q --durations=10 ) > $O/g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/g_smoke.log 2>&1; echo "smoke rc=$?" >> $O/g_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/g_bench_ref.json 2> $O/g_bench_ref.err
timeout 600 python bench.py --steps 20 --warmup 3 > $O/g_bench.json 2> $O/g_bench.err
timeout 600 python bench.py --workload cfg3 --steps 20 --warmup 3 --no-cpu-baseline > $O/g_bench_cfg3.json 2> $O/g_bench_cfg3.err
timeout 600 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline > $O/g_bench_cfg4.json 2> $O/g_bench_cfg4.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustained-seconds 0 > $O/g_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/g_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustained-seconds 0 > $O/g_ncu_launches.log 2>&1
for wl in cfg2 cfg3 cfg4; do
  WORKLOAD=$wl timeout 300 python tools/k4_only.py tc_fp16x2 3 > $O/g_k4_$wl.txt 2>&1 &&
  WORKLOAD=$wl timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_entity -s 1 -c 1 \
      -o $O/g_ncu_k4_$wl -f python tools/k4_only.py tc_fp16x2 3 > $O/g_ncu_k4_$wl.log 2>&1
done
timeout 300 python tools/sim_kernels.py > $O/g_sim_kernels.txt 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lookahead_group" -s 2 -c 1 \
      -o $O/g_ncu_sim -f python tools/sim_kernels.py > $O/g_ncu_sim.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"orca_kernel|step_kernel" -s 4 -c 2 \
      -o $O/g_ncu_k1k2 -f python tools/sim_kernels.py > $O/g_ncu_k1k2.log 2>&1
timeout 600 python tools/sim_only.py > $O/g_sim_only.txt 2>&1
for wl in cfg2 cfg3 cfg4; do
  WORKLOAD=$wl timeout 300 python tools/trace_tc.py tc_fp16x2 > $O/g_trace_$wl.txt 2>&1
  WORKLOAD=$wl timeout 300 python tools/trace_mma.py tc_fp16x2 > $O/g_trace_mma_$wl.txt 2>&1
done
ls -la $O > $O/g_ls.txt
