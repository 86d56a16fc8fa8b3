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
    scenes = synth.generate(shape, np.arange(min(N, 8192)))
    zero = torch.zeros(N, dtype=torch.int32, device="cuda:0")

    def prep():
        """The same state before every measurement: fresh scenes, ten steps into the crossing (the LPs have real work)."""
        synth.load(sim, scenes)
        if N > 8192:      # tile the first 8192 scenes over the batch
            for k in ("hum_pv", "hum_gr", "hum_type", "hum_count", "stat", "stat_count", "rect", "rect_count", "rob_pv", "rob_gr", "rob_theta", "time"):
                t = getattr(sim, k)
                t[8192:] = t[:8192].repeat((N // 8192 - 1,) + (1,) * (t.dim() - 1))
        for _ in range(10):
            sim.step(action_idx=zero, fused_orca=True)
        torch.cuda.synchronize()

    H = shape.H
    prep(); ms_f = timed(lambda: sim.step(action_idx=zero, fused_orca=True), 10)
    prep(); ms_u = timed(lambda: (sim.orca(), sim.step(action_idx=zero)), 10)        # K1 then K2, two launches
    prep(); ms_o = timed(sim.orca, 10)
    prep(); sim.orca(); ms_s = timed(lambda: sim.step(action_idx=zero), 10)          # K2 alone (hum_nv from one K1)
    print("%s N=%6d H=%2d: fused K1+K2 %.4f ms = %.3e agent-steps/s (%.1f GB/s algorithmic); two launches %.4f ms = %.3e agent-steps/s; "
          "K1 alone %.4f ms = %.3e human-steps/s; K2 alone %.4f ms" % (
              wl, N, H, ms_f, N * (H + 1) / ms_f * 1e3, N * (48 * H + 90) / ms_f / 1e6, ms_u, N * (H + 1) / ms_u * 1e3,
              ms_o, N * H / ms_o * 1e3, ms_s))
    del sim
