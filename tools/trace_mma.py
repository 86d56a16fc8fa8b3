import os, sys, ctypes
os.environ["EBC_TC_TRACE"]="1"
ROOT="/root/repo"
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
mode = sys.argv[1] if len(sys.argv)>1 else "tc_fp16x2"
sim.set_value_mode(mode)
for _ in range(2): sim.decide()
torch.cuda.synchronize()
buf=(ctypes.c_longlong*4096)()
sim.be.lib.ebc_debug_trace(sim.h, buf, 4096)
raw=np.array(buf[:], dtype=np.int64); t=raw[:2048]; t=t[t>0]; tm=raw[2048:]; tm=tm[tm!=0]
clk=tm>>16; idle=(tm>>8)&0xff; m=tm&0xff
per=11; mhz=1965.0
tile=12
base=t[tile*per]; end=t[(tile+1)*per]
print("crew stamps:", np.round((t[tile*per:(tile+1)*per+1]-base)/mhz,2))
sel=(clk>=base-200)&(clk<=end)
ks=0
for c,i,mm in zip(clk[sel],idle[sel],m[sel]):
    both=mm&(mm>>4)&0xf
    print("%7.2f us  idle %3d  kbar %s full %s"%((c-base)/mhz,i,format(mm&0xf,'04b')[::-1],format((mm>>4)&0xf,'04b')[::-1]))
