#!/usr/bin/env python
"""BASELINE config 5: 256 tyre-coefficient sets x 4,096 manoeuvres = 1,048,576 rollouts x 500 steps in FP64 and in
FP32; reports, per state component, the max and 99th percentile of |f32 - f64| / max(|f64|, 1) at steps 1, 10, 100
and 500 (the stated, measured FP32 drift bound) and both kernels' throughput.  Writes one JSON document.

    python tools/drift_report.py [out.json] [--sets 256 --man 4096]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import python_motionplanning_b200 as mp  # noqa: E402
from python_motionplanning_b200 import workloads as wl  # noqa: E402

NAMES = ["U", "V", "wz", "wFL", "wFR", "wRL", "wRR", "yaw", "x", "y", "ax_prev", "ay_prev"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out", nargs="?", default=None)
    ap.add_argument("--sets", type=int, default=256)
    ap.add_argument("--man", type=int, default=4096)
    a = ap.parse_args()
    eng = mp.Engine(0)
    sets, s0, d, t, pset = wl.config5_sweep(n_sets=a.sets, n_man=a.man)
    p = mp.VehicleParameters()
    for w in ("FL", "FR", "RL", "RR"):
        setattr(p, "B" + w, sets[:, 0])
        setattr(p, "C" + w, sets[:, 1])
        setattr(p, "D" + w, sets[:, 2])
    eng.set_params(p)
    B = s0.shape[1]
    checkpoints = [1, 10, 100, 500]
    states, timing = {}, {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for dt in ("f64", "f32"):
        td = torch.float64 if dt == "f64" else torch.float32
        s, dl, tq = eng.dev(s0, td), eng.dev(d, td), eng.dev(t, td)
        ps = eng.dev(pset, torch.int32)
        done, out = 0, {}
        for n in checkpoints:                       # rollouts are resumable: chain the carried state
            r = eng.rollout(s, dl, tq, wl.DT, n - done, hold=500, param_set=ps, dtype=dt, step0=done)
            s, done = r.state_end, n
            out[n] = r.state_end.double().cpu().numpy()
        states[dt] = out
        s = eng.dev(s0, td)
        for k in range(3):
            if k == 2:
                e0.record()
            eng.rollout(s, dl, tq, wl.DT, 500, hold=500, param_set=ps, dtype=dt)
        e1.record()
        torch.cuda.synchronize()
        timing[dt] = {"ms": e0.elapsed_time(e1), "rollout_steps_per_s": B * 500 / (e0.elapsed_time(e1) * 1e-3)}
    rep = {"workload": f"config5: {a.sets} tyre sets (B~U[8,25], C~U[1.2,1.9], D~U[0.3,1.2]) x {a.man} manoeuvres = {B} rollouts, "
                       "500 steps, dt 1e-4, step steer + constant torque",
           "metric": "|f32 - f64| / max(|f64|, 1) per carried component", "components": NAMES, "timing": timing, "steps": {}}
    for n in checkpoints:
        e = np.abs(states["f32"][n] - states["f64"][n]) / np.maximum(np.abs(states["f64"][n]), 1.0)
        finite = np.isfinite(e).all(axis=0)
        e = e[:, finite]
        rep["steps"][str(n)] = {"max": [float(v) for v in e.max(axis=1)], "p99": [float(v) for v in np.percentile(e, 99, axis=1)],
                                "max_over_states": float(e[:10].max()), "p99_over_states": float(np.percentile(e[:10], 99)),
                                "rollouts_compared": int(finite.sum())}
    txt = json.dumps(rep, indent=1)
    if a.out:
        with open(a.out, "w") as f:
            f.write(txt)
    for n in checkpoints:
        r = rep["steps"][str(n)]
        print(f"step {n:4d}: max over states {r['max_over_states']:.3e}  p99 {r['p99_over_states']:.3e}  (ax/ay max {max(r['max'][10:]):.3e})")
    print("timing", timing)


if __name__ == "__main__":
    main()
