import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import python_motionplanning_b200 as mp
from python_motionplanning_b200 import workloads as wl
from python_motionplanning_b200.host_numerics import host_norm2_mode
eng = mp.Engine(0)
p = mp.VehicleParameters(); p.DFL = p.DFR = p.DRL = p.DRR = 1.0
eng.set_params(p)
n_sets, vps, N = 12, 4096, 200
V = n_sets * vps
st0, wps = wl.tracking_fleet(V, n_sets)
mode = host_norm2_mode()
for N in (40, 80, 200):
    big = eng.track_closed_loop(st0, wps, wl.DT, N, 25.0, vehicles_per_set=vps, norm_mode=mode, want_target_idx=True)
    parts = [eng.track_closed_loop(st0[:, s*vps:(s+4)*vps], wps[s:s+4], wl.DT, N, 25.0, vehicles_per_set=vps, norm_mode=mode, want_target_idx=True) for s in range(0, n_sets, 4)]
    a = big.state_end.cpu().numpy(); b = torch.cat([q.state_end for q in parts], 1).cpu().numpy()
    d = a != b
    print("N", N, "nan", np.isnan(a).sum(), np.isnan(b).sum(), "mismatch entries", d.sum(), "vehicles", d.any(0).sum(), "first vehicles", np.where(d.any(0))[0][:10], "max abs diff", np.nanmax(np.abs(a - b)))
    ti = (big.target_idx.cpu().numpy() != torch.cat([q.target_idx for q in parts], 1).cpu().numpy())
    print("   target idx mismatches", ti.sum(), "first update with mismatch", np.where(ti.any(1))[0][:5])
    ce = (big.ctrl_end.cpu().numpy() != torch.cat([q.ctrl_end for q in parts], 1).cpu().numpy())
    print("   ctrl_end mismatches per row", ce.sum(1))
    if d.any():
        v = np.where(d.any(0))[0][0]
        print("   vehicle", v, "rows", np.where(d[:, v])[0], a[:, v] - b[:, v])
