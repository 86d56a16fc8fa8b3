#!/bin/bash
# Copies the evidence set of gpurun_out/g_* (tools/gpu_round2_g.sh) into profiles/ under the r2 names.
set -u
cd "$(dirname "$0")/.."
G=gpurun_out; P=profiles
for f in g_bench:r2_bench g_bench_cfg3:r2_bench_side_cfg3 g_bench_cfg4:r2_bench_side_cfg4 g_bench_ref:r2_bench_reference_arm; do
  grep '^{' $G/${f%%:*}.json > $P/${f##*:}.json
done
cp $G/g_launches.csv $P/r2_launches.csv
for w in cfg2 cfg3 cfg4; do
  cp $G/g_trace_$w.txt $P/r2_tc_entity_phase_trace_$w.txt
  cp $G/g_trace_mma_$w.txt $P/r2_tc_entity_mma_issue_trace_$w.txt
  python tools/ncu_extract.py $G/g_ncu_k4_$w.ncu-rep $P/r2_tc_entity_${w}_ncu_full.csv
done
python tools/ncu_extract.py $G/g_ncu_sim.ncu-rep $P/r2_sim_kernels_ncu.csv
[ -f $G/g_ncu_k1k2.ncu-rep ] && python tools/ncu_extract.py $G/g_ncu_k1k2.ncu-rep $P/r2_k1_k2_ncu_full.csv
cp $G/g_sim_only.txt $P/r2_sim_only.txt
cp $G/g_sim_kernels.txt $P/r2_sim_kernels.txt
tail -4 $G/g_pytest.log > $P/r2_gpu_pytest_tail.txt
