"""Times the UNMODIFIED reference's own CPU path on this host.  TEST / BENCH
INFRASTRUCTURE: executed only by ``bench.py``'s ``cpu_baseline`` leg (as a subprocess, so no CUDA context is forked)
and by ``tests/``.

    python -m oracle.literal_baseline [--rollouts-per-proc 16] [--paths-per-proc 2] [--procs N]

* rollouts: ``VehicleModel.planar_model_RK4`` (``libs/vehicle_model/vehicle_model.py:427-445``) called the way
  ``Car.drive`` calls it (``drive.py:141-143``), over the first rollouts of config 2, fanned out with
  ``multiprocessing.Pool(os.cpu_count())`` -- the reference's own parallel idiom (``local_planner.py:15, 369-374``);
* collision: ``CollisionChecker.collision_check`` (``libs/motionplanner/collision_checker.py:32-117``) on the first
  config-3 paths against the 10,000 obstacle points, one path per task in the same kind of pool.
The reference is imported from ``/root/reference`` or the staged verbatim copy ``baseline/_ref``.
Prints one JSON object.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from python_motionplanning_b200 import workloads as wl  # noqa: E402


_REF = None


def _init_worker():
    """Pool initializer: import the reference once per worker, outside the timed region."""
    global _REF
    _REF = ref_loader.load()


def _rollout_task(args):
    state0, delta, torque, n_steps, hold = args
    ref = _REF or ref_loader.load()
    VM = ref.vehicle_model
    vm = VM.VehicleModel(2.906, np.deg2rad(30), wl.DT)
    p = VM.VehicleParameters()
    t0 = time.perf_counter()
    end = np.zeros((state0.shape[1], 12))
    for k in range(state0.shape[1]):
        st = list(state0[:10, k])
        ax, ay = state0[10, k], state0[11, k]
        for n in range(n_steps):
            d, t = delta[n // hold, 0, k], torque[n // hold, 0, k]
            r = vm.planar_model_RK4(st, [t, t, t, t], [1.0, 1.0, 1.0, 1.0], [d, d, 0, 0], p, ax, ay)
            st, ax, ay = r[0], r[7], r[8]
        end[k, :10], end[k, 10], end[k, 11] = st, ax, ay
    return end, time.perf_counter() - t0


def _collision_task(args):
    path, obstacles = args
    ref = _REF or ref_loader.load()
    cc = ref.collision_checker.CollisionChecker(list(wl.CIRCLE_OFFSETS), list(wl.CIRCLE_RADII), wl.PATH_SELECT_WEIGHT)
    return bool(cc.collision_check(path, obstacles))


def run(rollouts_per_proc: int = 16, paths_per_proc: int = 2, procs: int | None = None, n_steps: int = 500, repeats: int = 1):
    procs = procs or os.cpu_count() or 1
    out = {"procs": procs, "reference_root": ref_loader.REFERENCE_ROOT}
    nb = procs * rollouts_per_proc
    if nb:
        s0, d, t = wl.config2_rollouts(B=65536, n_steps=n_steps)
        tasks = [(s0[:, k::procs][:, :rollouts_per_proc].copy(), d[:, :, k::procs][:, :, :rollouts_per_proc].copy(),
                  t[:, :, k::procs][:, :, :rollouts_per_proc].copy(), n_steps, wl.HOLD) for k in range(procs)]
        with mp.Pool(procs, initializer=_init_worker) as pool:
            pool.map(abs, range(procs))                 # workers forked and warm before the clock starts
            secs = []
            for _ in range(max(1, repeats)):
                t0 = time.perf_counter()
                res = pool.map(_rollout_task, tasks, chunksize=1)
                secs.append(time.perf_counter() - t0)
            el = sum(secs) / len(secs)
        out["rollout"] = {"value": nb * n_steps / el, "unit": "rollout-steps/s", "seconds": el, "seconds_each": secs, "rollouts": nb, "n_steps": n_steps,
                          "sample": f"config 2 rollouts k, k+{procs}, ... ({rollouts_per_proc} per process) x {n_steps} steps, "
                                    f"unmodified planar_model_RK4 under multiprocessing.Pool({procs})",
                          "end_state_checksum": float(np.sum([r[0] for r in res]))}
    npaths = procs * paths_per_proc
    if npaths:
        w = wl.config3_lattice()
        obstacles = w["obstacles"].tolist()
        paths = [[w["px"][i].tolist(), w["py"][i].tolist(), w["pyaw"][i].tolist()] for i in range(npaths)]
        with mp.Pool(procs, initializer=_init_worker) as pool:
            pool.map(abs, range(procs))
            t0 = time.perf_counter()
            flags = pool.map(_collision_task, [(p, obstacles) for p in paths], chunksize=1)
            el = time.perf_counter() - t0
        M, n = len(obstacles), len(paths[0][0])
        out["collision"] = {"value": npaths * n * 3 * M / el, "unit": "circle-point tests/s (nominal P*49*3*M)", "seconds": el,
                            "paths": npaths, "paths_per_s": npaths / el, "free": [bool(f) for f in flags],
                            "sample": f"first {npaths} config-3 paths vs {M} obstacle points, unmodified collision_check, one path "
                                      f"per task under multiprocessing.Pool({procs})"}
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rollouts-per-proc", type=int, default=16)
    ap.add_argument("--paths-per-proc", type=int, default=2)
    ap.add_argument("--procs", type=int, default=None)
    ap.add_argument("--n-steps", type=int, default=500)
    ap.add_argument("--repeats", type=int, default=1, help="timed repetitions of the rollout sample (one pool)")
    a = ap.parse_args()
    if not ref_loader.available():
        print(json.dumps({"unavailable": f"no reference under {ref_loader.REFERENCE_ROOT} or baseline/_ref"}))
        sys.exit(0)
    print(json.dumps(run(a.rollouts_per_proc, a.paths_per_proc, a.procs, a.n_steps, a.repeats)))
