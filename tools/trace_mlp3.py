# clock64 trace of the mlp3 tensor-core kernel (EBC_TC_TRACE=2): crew thread 0 of CTA 0 stamps, per tile,
# X handed over, mlp3.0 / mlp3.2 / mlp3.4 complete, H written, score read; the MMA warp logs (time: k-steps cleared).
import os, sys, ctypes
os.environ["EBC_TC_TRACE"]="2"
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,os.path.join(ROOT,"eb-cadrl_b200")); sys.path.insert(0,os.path.join(ROOT,"tests")); sys.path.insert(0,ROOT)
import numpy as np, torch
import bench
from ebc import synth
from ebc.actions import build_action_space
from ebc.engine import BatchedSim
WL="cfg2"
shape,cfg=bench.workload(WL); w,_=bench.value_net_weights(fixture=bench.WORKLOADS[WL][3])
N=4096
sim=BatchedSim(cfg,N,shape.H,shape.Smax,shape.Rmax,81,device="cuda:0")
sim.set_actions(build_action_space(shape.robot_v_pref, cfg.robot_kinematics)); sim.set_weights(w)
synth.load(sim, synth.generate(shape,np.arange(N)))
sim.set_value_mode("tc_fp16x2")
for _ in range(2): sim.decide()
torch.cuda.synchronize()
buf=(ctypes.c_longlong*4096)()
sim.be.lib.ebc_debug_trace(sim.h, buf, 4096)
raw=np.array(buf[:], dtype=np.int64); t=raw[:2048]; t=t[t>0]; tm=raw[2048:]; tm=tm[tm>0]
np.set_printoptions(linewidth=200, precision=2, suppress=True)
print("stamps", len(t))
print("crew stamps (us) from start:", (t[:40]-t[0])/1965.0)
clk=tm>>16; b=tm&0xffff
print("MMA batches:", " ".join(("%.2f:%d"%((c-t[0])/1965.0,k) if k!=0xffff else "[done %.2f]"%((c-t[0])/1965.0)) for c,k in zip(clk[:60],b[:60])))



