import sys, torch, numpy as np
sys.path.insert(0,'/root/repo')
import python_motionplanning_b200 as mp
from python_motionplanning_b200 import workloads as wl
eng=mp.Engine(0); w=wl.config3_lattice()
px,py,obs=eng.dev(w["px"]),eng.dev(w["py"]),eng.dev(w["obstacles"])
tr=eng.path_trig(w["pyaw"], w["px"].shape[1])
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
for k in range(4):
    if k==3: e0.record()
    r=eng.collision_check_batch(px,py,None,obs,w["offsets"],w["radii"],want_clearance=True,trig=tr)
e1.record(); torch.cuda.synchronize()
print("min-clearance kernel path: %.3f ms"%e0.elapsed_time(e1))
