import os, sys, ctypes
os.environ["EBC_TC_TRACE"]="1"
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,os.path.join(ROOT,"eb-cadrl_b200")); sys.path.insert(0,os.path.join(ROOT,"tests")); sys.path.insert(0,ROOT)
import numpy as np, torch
import bench
from ebc import synth
from ebc.actions import build_action_space
from ebc.engine import BatchedSim
shape,cfg=bench.workload(); w,_=bench.value_net_weights()
N=4096
sim=BatchedSim(cfg,N,shape.H,shape.Smax,shape.Rmax,81,device="cuda:0")
sim.set_actions(build_action_space(shape.robot_v_pref)); sim.set_weights(w)
synth.load(sim, synth.generate(shape,np.arange(N)))
mode = sys.argv[1] if len(sys.argv)>1 else "tc_fp32"
sim.set_value_mode(mode)
for _ in range(2): sim.decide()
torch.cuda.synchronize()
buf=(ctypes.c_longlong*4096)()
sim.be.lib.ebc_debug_trace(sim.h, buf, 4096)
raw=np.array(buf[:], dtype=np.int64); t=raw[:2048]; t=t[t>0]; tm=raw[2048:]; tm=tm[tm>0]
d=np.diff(t)
per=12
names=["MMA L0 (sigX->acc)","epi wide0","MMA L1A","epi wide1","MMA L1B","epi H1","MMA L2+L4 (G/GV overlapped)","epi T2","MMA L3","epi U","MMA L5","tail (score,softmax,pool,joint)+X stage"]
ntile=(len(t)-1)//per
arr=d[:ntile*per].reshape(ntile,per).astype(float)
mhz=1965.0
print("mode",mode,"tiles traced",ntile)
for i,nm in enumerate(names): print("%-45s mean %8.0f cyc = %6.2f us   (min %7.0f max %7.0f)"%(nm,arr[2:,i].mean(),arr[2:,i].mean()/mhz,arr[2:,i].min(),arr[2:,i].max()))
print("per tile total %.1f us"%(arr[2:].sum(1).mean()/mhz))

if len(tm) >= 60:
    # MMA-thread stamps: (kbar[0] passed, kbar[last] passed) for L1A, L1B, L2 of every tile = 6 per tile
    tile = 20
    base = t[tile*per]            # crew: signal_a(X) of that tile
    print("tile %d timeline (us from X signal):" % tile)
    crew = (t[tile*per:(tile+1)*per] - base)/mhz
    print("  crew stamps:", np.round(crew,2))
    print("  mma  stamps:", np.round((tm[tile*6:(tile+1)*6] - base)/mhz,2), "(kbar first/last passed for L1A, L1B, L2)")
