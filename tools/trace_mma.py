"""Per-k-step issue times of the K4 entity kernel's MMA warp (EBC_TC_TRACE=3): CTA 0's elected lane stamps clock64()
after every k-step it issues, tagged with the schedule entry.  Prints, for the tiles in the middle of the trace, the
cycles between consecutive issues per schedule entry (mean / max) and one tile in full, next to the tensor pipe's time
for that k-step (3 MMAs x N / 2 cycles in the fp16x2 mode).
    WORKLOAD=cfg2 python tools/trace_mma.py [mode]"""
import os, sys, ctypes
os.environ["EBC_TC_TRACE"] = "3"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "eb-cadrl_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from ebc import synth
from ebc.actions import build_action_space
from ebc.engine import BatchedSim
WL = os.environ.get("WORKLOAD", "cfg2")
shape, cfg = bench.workload(WL); w, _ = bench.value_net_weights(fixture=bench.WORKLOADS[WL][3])
N = min(bench.WORKLOADS[WL][4], 4096)
sim = BatchedSim(cfg, N, shape.H, shape.Smax, shape.Rmax, 81, device="cuda:0")
sim.set_actions(build_action_space(shape.robot_v_pref, cfg.robot_kinematics)); sim.set_weights(w)
synth.load(sim, synth.generate(shape, np.arange(N)))
mode = sys.argv[1] if len(sys.argv) > 1 else "tc_fp16x2"
sim.set_value_mode(mode)
for _ in range(2): sim.decide()
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 4096)()
sim.be.lib.ebc_debug_trace(sim.h, buf, 4096)
raw = np.array(buf[:], dtype=np.int64)
np.save(os.path.join(ROOT, "gpurun_out", "trace_mma_raw_%s_%s.npy" % (WL, mode)), raw)
tm = raw[2048:]; tm = tm[tm > 0]
clk, tag, ks = tm >> 16, (tm >> 8) & 0xFF, tm & 0xFF
starts = np.where((tag == 0) & (ks == 0))[0]            # first k-step of every tile
per = int(np.diff(starts)[0])
print("workload", WL, "mode", mode, "k-steps per tile", per, "tiles traced", len(starts) - 1)
tiles = [(s, s + per) for s in starts[3:-1]]
d = np.stack([np.diff(np.concatenate([[clk[a - 1]], clk[a:b]])) for a, b in tiles]).astype(float)   # cycles since the previous issue
tg, kk = tag[tiles[0][0]:tiles[0][1]], ks[tiles[0][0]:tiles[0][1]]
print("tile period: %.0f cycles = %.2f us" % (d.sum(1).mean(), d.sum(1).mean() / 1965.0))
print("schedule entry: k-steps, mean cycles per k-step inside the entry (excluding its first), first k-step's gap, total")
for t in np.unique(tg):
    m = tg == t
    inner = d[:, m][:, 1:]
    print("  entry %d: %2d k-steps, inner mean %6.0f (max %6.0f), first gap mean %6.0f, total %7.0f cycles = %.2f us" % (
        t, m.sum(), inner.mean() if inner.size else 0, inner.max() if inner.size else 0, d[:, m][:, 0].mean(),
        d[:, m].sum(1).mean(), d[:, m].sum(1).mean() / 1965.0))
mid = len(tiles) // 2
print("one tile (entry:k-step gap in cycles):")
print(" ".join("%d:%d %d" % (a, b, c) for a, b, c in zip(tg, kk, d[mid].astype(int))))
