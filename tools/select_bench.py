import sys, time, torch, numpy as np
sys.path.insert(0,'/root/repo')
import python_motionplanning_b200 as mp
from python_motionplanning_b200 import workloads as wl
eng=mp.Engine(0)
w=wl.config3_lattice()
free=eng.collision_check_batch(w["px"],w["py"],w["pyaw"],w["obstacles"],w["offsets"],w["radii"])
ex,ey=eng.dev(w["px"][:,-1].copy()),eng.dev(w["py"][:,-1].copy())
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
import ctypes as C
best=eng.empty(1,dtype=torch.int32)
for k in range(5):
    if k==4: e0.record()
    eng.lib.b200mp_select_best_f64(0, eng._stream(), 4096, C.c_void_p(ex.data_ptr()), C.c_void_p(ey.data_ptr()), C.c_void_p(free.data_ptr()), 50.0, 50.0, 10.0, 1, None, C.c_void_p(best.data_ptr()))
e1.record(); torch.cuda.synchronize(); print("select_best kernels", e0.elapsed_time(e1), "ms", int(best.item()))
