import os, sys, torch
sys.path.insert(0, '/root/repo')
import python_motionplanning_b200 as mp
from python_motionplanning_b200 import workloads as wl
eng = mp.Engine(0); p = mp.VehicleParameters(); p.DFL = p.DFR = p.DRL = p.DRR = 1.0; eng.set_params(p)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for mode in ("auto", "closed_form"):
    eng.set_friction_mode(mode)
    for B, N, ns in ((65536, 500, 16), (65536, 500, 1), (65536, 100, 16), (262144, 100, 16)):
        st0, wps = wl.tracking_fleet(V=B, n_sets=ns)
        s, w = eng.dev(st0), eng.dev(wps)
        for k in range(3):
            if k == 2: e0.record()
            r = eng.track_closed_loop(s, w, 1e-4, N, 25.0, vehicles_per_set=-(-B // ns))
        e1.record(); torch.cuda.synchronize()
        print(mode, B, N, ns, f"{e0.elapsed_time(e1):.3f} ms")
