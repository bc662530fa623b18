#!/usr/bin/env python
"""Randomised check of the tracker's nearest-waypoint search + look-ahead walk against the C oracle (development tool):
random smooth, kinked, self-approaching and noisy waypoint lists of random length, vehicles at random offsets; one control
update per vehicle with a hint-less search, then a second launch continuing from the first (hinted search).

    python tools/track_fuzz.py [n_cases] [seed]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import python_motionplanning_b200 as mp  # noqa: E402
from oracle import c_oracle, planar_numpy as pn  # noqa: E402
from python_motionplanning_b200.host_numerics import host_norm2_mode  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
eng = mp.Engine(0)
p = mp.VehicleParameters()
p.DFL = p.DFR = p.DRL = p.DRR = 1.0
eng.set_params(p)
par = c_oracle.make_params(pn.VehicleParams())
par[0].D[:] = (1.0,) * 4
mode = host_norm2_mode()
bad = 0
for case in range(n_cases):
    W = int(rng.choice([1, 7, 8, 9, 63, 64, 65, 511, 512, 513, 1000, 3000, 4097, 6000, 12000]))
    ds = float(rng.choice([0.01, 0.05, 0.3]))
    kind = case % 4
    s = np.arange(W) * ds
    if kind == 0:      # clothoid-like
        th = rng.uniform(-3, 3) + rng.uniform(-0.05, 0.05) * s + rng.uniform(-0.002, 0.002) * s ** 2
    elif kind == 1:    # kinks
        th = rng.uniform(-3, 3) + np.cumsum(np.where(rng.random(W) < 0.01, rng.uniform(-1.5, 1.5, W), 0.0))
    elif kind == 2:    # tight loops that come back over themselves
        th = rng.uniform(-3, 3) + rng.uniform(0.05, 0.5) * s
    else:              # noisy heading
        th = rng.uniform(-3, 3) + np.cumsum(rng.normal(0, 0.05, W))
    wp = np.stack([np.concatenate([[0.0], np.cumsum(np.cos(th[:-1]) * ds)]), np.concatenate([[0.0], np.cumsum(np.sin(th[:-1]) * ds)])], 1)
    wp += rng.uniform(-100, 100, 2)
    if kind == 3 and W > 20:
        wp[rng.integers(0, W, 3)] = wp[rng.integers(0, W, 3)]          # duplicates
    V = 1024
    i0 = rng.integers(0, W, V)
    off = rng.choice([0.0, 0.01, 0.3, 3.0, 30.0], V) * rng.normal(0, 1, (2, V))
    xs, ys = wp[i0, 0] + off[0], wp[i0, 1] + off[1]
    s0 = np.zeros((12, V))
    s0[0] = 20.0
    s0[3:7] = 20.0 / 0.308309813617345
    s0[7] = rng.uniform(-3, 3, V)
    s0[8], s0[9] = xs, ys
    r1 = eng.track_closed_loop(s0, wp, 1e-4, 20, 25.0, ctrl_every=10, want_target_idx=True, vehicles_per_set=V, norm_mode=mode)
    c0 = np.zeros((3, V))
    c0[2] = s0[0]
    ref = c_oracle.track_loop(s0, c0, wp, None, par, 1e-4, 20, 25.0, c_oracle.track_gains(), mode, ctrl_every=10, vehicles_per_set=V)
    got = r1.target_idx.cpu().numpy()
    mism = int((got != ref["target_idx"]).sum())
    bad += mism
    print(f"case {case:2d} kind {kind} W={W:5d} ds={ds}: {mism} index mismatches of {got.size}")
print("TOTAL MISMATCHES", bad)
sys.exit(1 if bad else 0)
