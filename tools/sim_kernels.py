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
# angular local map: 3 wall quadrilaterals per episode, 48 sectors
del big
for N in (4096, 65536):
    sim = make(N, False)
    rng = np.random.default_rng(1)
    c = rng.uniform(-3.5, 3.5, (N, 3, 2)); hw = np.stack([np.full((N, 3), 1.5), np.full((N, 3), 0.5)], -1)
    xy = np.zeros((N, 3, 4, 2))
    for k, (sx, sy) in enumerate(((1, 1), (-1, 1), (-1, -1), (1, -1))):
        xy[:, :, k, 0] = c[..., 0] + sx * hw[..., 0]; xy[:, :, k, 1] = c[..., 1] + sy * hw[..., 1]
    xy_d = torch.tensor(xy, device="cuda:0"); cnt = torch.full((N,), 3, dtype=torch.int32, device="cuda:0")
    out = torch.empty(N, 48, dtype=torch.float64, device="cuda:0")
    ms = timed(lambda: sim.local_map_angular(xy_d, cnt, 3.0, -np.pi, np.pi, 48, out=out), 10)
    print("angular local map, %d episodes x 3 obstacles x 48 sectors: %.4f ms = %.3e maps/s (%.1f us per 1000 episodes)" % (
        N, ms, N / ms * 1e3, ms * 1e3 / N * 1000))
    # binary grid sub-map (use_grid_map = true): 5 m window of the 0.1 m grid = 50 x 50 cells per episode,
    # the scene's 3 wall rectangles, random headings; algorithmic bytes = the size x size bytes written + the rectangles read
    sim.rob_theta.copy_(torch.tensor(rng.uniform(-np.pi, np.pi, N), dtype=torch.float32))
    sim.rob_pv[:, :2] = torch.tensor(rng.uniform(-4.0, 4.0, (N, 2)), dtype=torch.float32)
    for sub in (5.0, 6.0):
        S = int(round(sub / 0.1))
        g = torch.empty(N, S, S, dtype=torch.uint8, device="cuda:0")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
        sim.local_map_grid(sub, out=g); torch.cuda.synchronize()
        evs = []
        for _ in range(10):     # L2 flushed before every timed launch (a 25-236 MB output would otherwise partly stay in L2)
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); sim.local_map_grid(sub, out=g); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
        byt = N * (S * S + 8 * shape.Rmax + 20)
        print("grid local map, %d episodes x %d x %d cells: %.4f ms = %.3e maps/s, %.1f GB/s algorithmic (%.0f MB), zero cells %.1f %%" % (
            N, S, S, ms, N / ms * 1e3, byt / ms / 1e6, byt / 1e6, 100.0 * float((g == 0).float().mean())))
        del g, flush
    del sim
