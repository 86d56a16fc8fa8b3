#!/bin/bash
# Round-2 GPU batch B: new tests only, cfg5 profile, sim-only before/after (half-warp K1), cfg3/cfg4 K4 after the reductions change
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q --durations=8 --deselect tests/test_gpu_round2.py::test_parity_report_10k_decisions ) > $O/b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/b_pytest.log
timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 1 --profile > $O/b_bench_cfg5.json 2> $O/b_bench_cfg5.err
timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 1 --episodes-per-iter 64 > $O/b_bench_cfg5_k64.json 2> $O/b_bench_cfg5_k64.err
for wl in cfg3 cfg4; do
  WORKLOAD=$wl timeout 300 python tools/k4_only.py tc_fp16x2 5 > $O/b_k4_$wl.txt 2>&1
done
timeout 300 python tools/k4_only.py tc_fp16x2 5 > $O/b_k4_cfg2.txt 2>&1
for g in 16 32; do
  EBC_ORCA_GROUP=$g timeout 600 python tools/sim_only.py > $O/b_sim_only_g$g.txt 2>&1
done
ls -la $O > $O/b_ls.txt
