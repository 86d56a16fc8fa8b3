"""K1 + K2 (fused ORCA + committed step) and K1 alone at several batch sizes / shapes: the sim-only rate.
    [EBC_ORCA_GROUP=32] python tools/sim_only.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "eb-cadrl_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from ebc import synth
from ebc.actions import build_action_space
from ebc.engine import BatchedSim


def timed(fn, iters):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


print("EBC_ORCA_GROUP =", os.environ.get("EBC_ORCA_GROUP", "(default: 16 where it fits)"))
for wl, N in (("cfg2", 4096), ("cfg2", 65536), ("cfg4", 65536), ("cfg3", 65536)):
    shape, cfg = bench.workload(wl)
    sim = BatchedSim(cfg, N, shape.H, shape.Smax, shape.Rmax, 81, device="cuda:0")
    sim.set_actions(build_action_space(shape.robot_v_pref, cfg.robot_kinematics))
    synth.load(sim, synth.generate(shape, np.arange(min(N, 8192))))
    if N > 8192:      # tile the first 8192 scenes over the batch
        for k in ("hum_pv", "hum_gr", "hum_type", "hum_count", "stat", "stat_count", "rect", "rect_count", "rob_pv", "rob_gr", "rob_theta"):
            t = getattr(sim, k)
            t[8192:] = t[:8192].repeat((N // 8192 - 1,) + (1,) * (t.dim() - 1))
    zero = torch.zeros(N, dtype=torch.int32, device="cuda:0")
    for _ in range(10):      # a few steps into the crossing: the LPs have real work
        sim.step(action_idx=zero, fused_orca=True)
    ms_f = timed(lambda: sim.step(action_idx=zero, fused_orca=True), 20)
    ms_o = timed(sim.orca, 20)
    H = shape.H
    print("%s N=%6d H=%2d: fused K1+K2 %.4f ms = %.3e agent-steps/s (%.1f GB/s algorithmic); K1 alone %.4f ms = %.3e human-steps/s" % (
        wl, N, H, ms_f, N * (H + 1) / ms_f * 1e3, N * (48 * H + 90) / ms_f / 1e6, ms_o, N * H / ms_o * 1e3))
    del sim
