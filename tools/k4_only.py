"""K4 (value network) alone on the bench workload: a short command for ncu captures.
    python tools/k4_only.py [mode] [iterations]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "eb-cadrl_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from ebc import synth
from ebc.actions import build_action_space
from ebc.engine import BatchedSim
WL = os.environ.get("WORKLOAD", "cfg2")
shape, cfg = bench.workload(WL); w, _ = bench.value_net_weights(fixture=bench.WORKLOADS[WL][3])
N = int(os.environ.get("EPISODES", bench.WORKLOADS[WL][4]))
sim = BatchedSim(cfg, N, shape.H, shape.Smax, shape.Rmax, 81, device="cuda:0")
sim.set_actions(build_action_space(shape.robot_v_pref, cfg.robot_kinematics)); sim.set_weights(w)
synth.load(sim, synth.generate(shape, np.arange(N)))
sim.set_value_mode(sys.argv[1] if len(sys.argv) > 1 else "tc_fp16x2")
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
FUSED = os.environ.get("FUSED", "1") != "0"        # FUSED=0: the materialised input (K3 writes N*A*n*D floats, K4 reads them)
sim.orca(); sim.lookahead(build_inputs=not FUSED)
run = (lambda: sim.value(fused=True)) if FUSED else sim.value
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
run(); torch.cuda.synchronize()
a.record()
for _ in range(iters): run()
b.record(); torch.cuda.synchronize()
print("K4 %s %s N=%d %s input: %.3f ms per call" % (WL, sim.value_mode(), N, "fused" if FUSED else "materialised", a.elapsed_time(b) / iters))
