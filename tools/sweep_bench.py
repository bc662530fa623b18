#!/usr/bin/env python
"""Config 5 (parameter sweep: 256 tyre sets x 4,096 manoeuvres, set-major) and per-rollout mu_max on the FP64 generic
kernel: tabulated (set-uniform blocks) vs closed form -- timing and agreement (development tool)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import python_motionplanning_b200 as mp  # noqa: E402
from python_motionplanning_b200 import workloads as wl  # noqa: E402

n_sets = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n_man = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
N = int(sys.argv[3]) if len(sys.argv) > 3 else 100
eng = mp.Engine(0)
sets, st, dl, tq, ps = wl.config5_sweep(n_sets=n_sets, n_man=n_man)
plist = mp.VehicleParameters()
for w in ("FL", "FR", "RL", "RR"):
    setattr(plist, "B" + w, sets[:, 0])
    setattr(plist, "C" + w, sets[:, 1])
    setattr(plist, "D" + w, sets[:, 2])
t0 = time.perf_counter()
eng.set_params(plist)
print(f"set_params({n_sets} sets): {time.perf_counter() - t0:.2f} s")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a, b, c, s = eng.dev(st), eng.dev(dl), eng.dev(tq), eng.dev(ps, torch.int32)
B = a.shape[1]
res = {}
for mode in ("auto", "closed_form"):
    eng.set_friction_mode(mode)
    for k in range(4):
        if k == 3:
            e0.record()
        r = eng.rollout(a, b, c, wl.DT, N, hold=N, param_set=s)
    e1.record()
    torch.cuda.synchronize()
    res[mode] = r.state_end.cpu().numpy()
    print(f"param sweep {mode:12s} B={B} N={N}: {e0.elapsed_time(e1):.3f} ms  {B * N / e0.elapsed_time(e1) * 1e3:.3e} steps/s")
e = np.abs(res["auto"] - res["closed_form"]) / np.maximum(np.abs(res["closed_form"]), 1.0)
print("   tabulated vs closed form after", N, "steps: max rel", e.max())
# shuffled sets (blocks not uniform -> closed-form step inside the same kernel)
perm = torch.randperm(B, device=a.device)
eng.set_friction_mode("auto")
r2 = eng.rollout(a[:, perm].contiguous(), b[:, :, perm].contiguous(), c[:, :, perm].contiguous(), wl.DT, N, hold=N, param_set=s[perm].contiguous())
e = np.abs(r2.state_end.cpu().numpy() - res["closed_form"][:, perm.cpu().numpy()])
print("   shuffled sets (mixed blocks) vs closed form: max abs", e.max())
# one set, per-rollout mu_max
mu = eng.dev(np.random.default_rng(1).uniform(0.3, 1.1, (4, B)))
for mode in ("auto", "closed_form"):
    eng.set_friction_mode(mode)
    for k in range(4):
        if k == 3:
            e0.record()
        r = eng.rollout(a, b, c, wl.DT, N, hold=N, mu=mu)
    e1.record()
    torch.cuda.synchronize()
    res[mode] = r.state_end.cpu().numpy()
    print(f"mu_max      {mode:12s} B={B} N={N}: {e0.elapsed_time(e1):.3f} ms  {B * N / e0.elapsed_time(e1) * 1e3:.3e} steps/s")
e = np.abs(res["auto"] - res["closed_form"]) / np.maximum(np.abs(res["closed_form"]), 1.0)
print("   tabulated vs closed form after", N, "steps: max rel", e.max())
