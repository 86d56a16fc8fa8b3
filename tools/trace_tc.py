import os, sys, ctypes
os.environ["EBC_TC_TRACE"]="1"
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,os.path.join(ROOT,"eb-cadrl_b200")); sys.path.insert(0,os.path.join(ROOT,"tests")); sys.path.insert(0,ROOT)
import numpy as np, torch
import bench
from ebc import synth
from ebc.actions import build_action_space
from ebc.engine import BatchedSim
WL=os.environ.get("WORKLOAD","cfg2")
shape,cfg=bench.workload(WL); w,_=bench.value_net_weights(fixture=bench.WORKLOADS[WL][3])
N=min(bench.WORKLOADS[WL][4],4096)
sim=BatchedSim(cfg,N,shape.H,shape.Smax,shape.Rmax,81,device="cuda:0")
sim.set_actions(build_action_space(shape.robot_v_pref, cfg.robot_kinematics)); sim.set_weights(w)
synth.load(sim, synth.generate(shape,np.arange(N)))
mode = sys.argv[1] if len(sys.argv)>1 else "tc_fp32"
sim.set_value_mode(mode)
print("workload",WL,"episodes",N,"rows per state",shape.H+shape.Smax)
for _ in range(2): sim.decide()
torch.cuda.synchronize()
buf=(ctypes.c_longlong*4096)()
sim.be.lib.ebc_debug_trace(sim.h, buf, 4096)
raw=np.array(buf[:], dtype=np.int64)
os.makedirs(os.path.join(ROOT,"gpurun_out"),exist_ok=True); np.save(os.path.join(ROOT,"gpurun_out","trace_raw_%s_%s.npy"%(WL,mode)), raw)
t=raw[:2048]; t=t[t>0]; tm=raw[2048:]; tm=tm[tm>0]
# Stamps of CTA 0's crew thread 0.  The first tile has 13 (X, acc1, W0a, W0b, acc2, H1, G, T2, acc3, U, acc4, score, sync);
# every later one starts with the early hand-over of its input during the previous tile's tail and has 15.
wide=[v for k,v in w.items() if k.endswith("mlp1.0.weight")][0].shape[0]
two=((wide+15)//16*16)>208          # the wide layer is issued as two halves (one more stamp per tile)
t=t[(13 if two else 12):]
d=np.diff(t)
per=15 if two else 14
names=["tail of the previous tile (softmax, pooling, joint) under this tile's mlp1.0","crew sync",
       "wait mlp1.0 complete","epilogue wide half 0","epilogue wide half 1 (chases mlp1.2a)","wait mlp1.2 complete",
       "epilogue H1 (+ state means)","G operand (chases attention.0 local)","epilogue T2 (chases attention.0 global)",
       "crew sync + wait attention.0 complete","epilogue U (chases mlp2.2)","wait attention.2 complete",
       "score epilogue","crew sync","next tile's X staged and handed over"]
if not two: names.remove("epilogue wide half 1 (chases mlp1.2a)")
ntile=(len(t)-1)//per
arr=d[:ntile*per].reshape(ntile,per).astype(float)
mhz=1965.0
print("mode",mode,"tiles traced",ntile)
for i,nm in enumerate(names): print("%-48s mean %8.0f cyc = %6.2f us   (min %7.0f max %7.0f)"%(nm,arr[2:,i].mean(),arr[2:,i].mean()/mhz,arr[2:,i].min(),arr[2:,i].max()))
print("per tile total %.1f us"%(arr[2:].sum(1).mean()/mhz))
# MMA-warp entries (upper half): (clock << 16) | k-steps issued in that batch
if len(tm) > 0:
    clk = tm >> 16; batch = tm & 0xffff
    for tile in (20,):
        base = t[tile*per]; end = t[(tile+1)*per]
        print("tile %d, crew stamps (us): X, tail, sync, acc1, W0a,%s acc2, H1, G, T2, acc3, U, acc4, score, sync, next X" % (tile, " W0b," if two else ""))
        print("  ", np.round((t[tile*per:(tile+1)*per+1] - base)/mhz, 2))
        sel = (clk >= base - 200) & (clk <= end)
        print("  MMA warp batches (us: k-steps):", " ".join("%.2f:%d" % ((c - base)/mhz, b) for c, b in zip(clk[sel], batch[sel])))
