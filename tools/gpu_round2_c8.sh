#!/bin/bash
# Round-2 GPU batch C (ONE box with 8 B200s): BASELINE configs[4] (cfg5, the RL training loop) at N = 1 / 2 / 4 / 8 with the
# gradient all-reduce timed, configs[3] (cfg4) as 65,536 episodes over the 8 GPUs, the bench line (cfg2) at N = 8, and the
# rl.train entry point on 2 ranks with per-rank imitation memories that straddle the batch size.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.max.sm,power.limit --format=csv > $O/c8_gpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29610
timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 1 > $O/c8_cfg5_n1.json 2> $O/c8_cfg5_n1.err
for n in 2 4 8; do
  port=$((port + 1))
  timeout 900 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n --workload cfg5 --steps 3 --warmup 1 \
      > $O/c8_cfg5_n$n.json 2> $O/c8_cfg5_n$n.err
done
port=$((port + 1))
timeout 900 $TR --nproc-per-node 8 --master-port $port bench.py --gpus 8 --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline \
    > $O/c8_cfg4_n8.json 2> $O/c8_cfg4_n8.err
port=$((port + 1))
timeout 900 $TR --nproc-per-node 8 --master-port $port bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline \
    > $O/c8_cfg2_n8.json 2> $O/c8_cfg2_n8.err
port=$((port + 1))
CFG=tests/golden/configs
( cd eb-cadrl_b200 && timeout 900 $TR --nproc-per-node 2 --master-port $port -m rl.train --env_config ../$CFG/env_fast_train.config \
    --policy sarl --policy_config ../$CFG/policy.config --train_config ../$CFG/test_train.config \
    --output_dir /tmp/train2 --episodes_per_iter 8 --gpu ) > $O/c8_train_2gpu.log 2>&1
echo "train rc=$?" >> $O/c8_train_2gpu.log
ls -la $O > $O/c8_ls.txt
