"""K1/K2 (fused ORCA + step) at 65,536 episodes and K3 (lookahead) at 4,096 episodes of the bench workload:
a short command for ncu captures and for the large-batch sim-only rate.
    python tools/sim_kernels.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "eb-cadrl_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from ebc import synth
from ebc.actions import build_action_space
from ebc.engine import BatchedSim
shape, cfg = bench.workload(); w, _ = bench.value_net_weights()
def make(N, with_vin):
    sim = BatchedSim(cfg, N, shape.H, shape.Smax, shape.Rmax, 81, device="cuda:0")
    sim.set_actions(build_action_space(shape.robot_v_pref)); sim.set_weights(w)
    synth.load(sim, synth.generate(shape, np.arange(N)))
    return sim
def timed(fn, iters):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
sim = make(4096, True)
sim.orca()
ms = timed(sim.lookahead, 5)
n = shape.H + shape.Smax
print("K3 lookahead, 4096 episodes x 81 actions: %.3f ms, %.1f GB/s of value-network input written (%.0f MB)" % (
    ms, 4096 * 81 * n * cfg.D * 4 / ms / 1e6, 4096 * 81 * n * cfg.D * 4 / 1e6))
del sim
N = 65536
big = make(N, False)
zero = torch.zeros(N, dtype=torch.int32, device="cuda:0")
ms = timed(lambda: big.step(action_idx=zero, fused_orca=True), 10)
print("K1+K2 fused ORCA + step, %d episodes: %.3f ms = %.3e agent-steps/s, %.1f GB/s algorithmic (48 H + 90 B per episode-step)" % (
    N, ms, N * (shape.H + 1) / ms * 1e3, N * (48 * shape.H + 90) / ms / 1e6))
