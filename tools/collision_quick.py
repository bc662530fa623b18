"""Kernel times of the config-3 collision paths (flags / min-clearance), broad-phase statistics; B200MP_NO_OBS_SORT=1 keeps the caller's obstacle order (needs a library built with -DB200MP_DEV_TUNABLES=1)."""
import ctypes as C, os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import python_motionplanning_b200 as mp
from python_motionplanning_b200 import workloads as wl
eng = mp.Engine(0); w = wl.config3_lattice()
px, py, obs = eng.dev(w["px"]), eng.dev(w["py"]), eng.dev(w["obstacles"])
tr = eng.path_trig(w["pyaw"], w["px"].shape[1])
rng = np.random.default_rng(0)
shuf = eng.dev(w["obstacles"][rng.permutation(len(w["obstacles"]))])
def ms(fn, reps=10):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps + 3)]
    for a, b in evs:
        a.record(); r = fn(); b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in evs[3:]), r
for name, ob in (("outline order", obs), ("shuffled", shuf)):
    t, fr = ms(lambda: eng.collision_check_batch(px, py, None, ob, w["offsets"], w["radii"], trig=tr))
    st = (C.c_ulonglong * 2)(); eng.lib.b200mp_collision_stats(eng.device, eng._stream(), len(w["obstacles"]), st)
    t2, r2 = ms(lambda: eng.collision_check_batch(px, py, None, ob, w["offsets"], w["radii"], trig=tr, want_clearance=True))
    print(f"{name:14s} flags {t:.3f} ms (warp-chunks screened {st[0]}, rechecked {st[1]}, free {int(fr.sum())})   min-clearance {t2:.3f} ms   sort={'off' if os.environ.get('B200MP_NO_OBS_SORT') else 'on'}")
