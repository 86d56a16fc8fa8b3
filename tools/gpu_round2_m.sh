#!/bin/bash
# Round-2 GPU batch M: grid local map (use_grid_map = true) -- parity tests, timing, ncu capture of the
# 65,536-episode launch (launch 23 of the kernel in tools/sim_kernels.py), then the whole -m gpu suite
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time timeout 600 python -m pytest tests/test_local_map_grid.py -m gpu -q -x ) > $O/m_pytest_grid.log 2>&1; echo "pytest rc=$?" >> $O/m_pytest_grid.log
timeout 600 python tools/sim_kernels.py > $O/m_sim_kernels.txt 2>&1; echo "rc=$?" >> $O/m_sim_kernels.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grid_map -s 22 -c 1 -o $O/m_grid_map python tools/sim_kernels.py > $O/m_ncu.log 2>&1
if [ "${1:-}" = "all" ]; then
  ( time timeout 1500 python -m pytest tests -m gpu -q -x ) > $O/m_pytest_all.log 2>&1; echo "pytest rc=$?" >> $O/m_pytest_all.log
  tail -n 6 $O/m_pytest_all.log
fi
tail -n 3 $O/m_pytest_grid.log; cat $O/m_sim_kernels.txt
