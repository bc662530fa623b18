"""CPU-only checks: the C-ABI library loads and exports what ``include/b200mp.h`` declares, struct
layouts agree between the header and the ctypes mirror, and the host-side logic (pool seam, install,
numerics probe, sharding, workloads) behaves like the reference's.  No compute call is made."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import python_motionplanning_b200 as mp
from python_motionplanning_b200 import _lib, distributed as D, workloads as wl
from python_motionplanning_b200.host_numerics import host_norm2_mode, norm2_closed_form

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200mp.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200mp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = mp.load_library()
    names = _declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200mp.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert sorted(_lib.PROTOTYPES) == names
    assert lib.b200mp_version() == 100


def test_struct_layouts_match_header(tmp_path):
    """Compile the header as plain C and compare sizeof/offsetof with the ctypes mirrors."""
    fields_p = [f[0] for f in _lib.VehicleParamsC._fields_]
    fields_r = [f[0] for f in _lib.RolloutArgsC._fields_]
    fields_t = [f[0] for f in _lib.TrackArgsC._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){',
            'printf("%zu\\n", sizeof(B200mpVehicleParams));', 'printf("%zu\\n", sizeof(B200mpRolloutArgs));',
            'printf("%zu\\n", sizeof(B200mpTrackArgs));']
    prog += [f'printf("%zu\\n", offsetof(B200mpVehicleParams, {f}));' for f in fields_p]
    prog += [f'printf("%zu\\n", offsetof(B200mpRolloutArgs, {f}));' for f in fields_r]
    prog += [f'printf("%zu\\n", offsetof(B200mpTrackArgs, {f}));' for f in fields_t]
    prog += ['return 0;}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", str(src), "-o", str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert out[0] == C.sizeof(_lib.VehicleParamsC) and out[1] == C.sizeof(_lib.RolloutArgsC)
    assert out[2] == C.sizeof(_lib.TrackArgsC)
    want = [getattr(_lib.VehicleParamsC, f).offset for f in fields_p] + [getattr(_lib.RolloutArgsC, f).offset for f in fields_r]
    want += [getattr(_lib.TrackArgsC, f).offset for f in fields_t]
    assert out[3:] == want


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mp.B200mpError):
        mp.Engine(0)
    with pytest.raises(mp.B200mpError):
        mp.VehicleModel(2.906, 0.52, 1e-4).planar_model_RK4([25, 0, 0, 81, 81, 81, 81, 0, 0, 0], [0] * 4, [1.0] * 4,
                                                            [0, 0, 0, 0], mp.VehicleParameters(), 0, 0)
    with pytest.raises(mp.B200mpError):
        mp.CollisionChecker([-1, 1, 3], [1.5] * 3, 10).collision_check([[1.0], [0.0], [0.0]], [[1.0, 0.0]])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "python_motionplanning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text, f


def test_vehicle_parameters_surface():
    p = mp.VehicleParameters()
    assert p.m == 987.89 + 869.93 and p.rw == 0.308309813617345 and p.b == 1.5708108108108108
    assert p.a == 1.3351891891891894 and p.Izz == 1948.2304506781593 and p.wL == p.wR == 0.768
    assert (p.BRR, p.CRL, p.DFR) == (20.6357, 1.5047, 1.1233) and p.E == [0.0376, 0.0376, 0, 0]
    q = mp.VehicleParameters(mf=1000, BFL=10.0)
    assert q.m == 1000 + 869.93 and q.BRL == 10.0
    arr = mp.pack_params(p)
    assert len(arr) == 1 and arr[0].rw == p.rw and list(arr[0].B) == [20.6357] * 4
    p.BFL = p.BFR = p.BRL = p.BRR = np.array([8.0, 9.0, 10.0])
    arr = mp.pack_params(p)
    assert len(arr) == 3 and [a.B[2] for a in arr] == [8.0, 9.0, 10.0] and arr[2].C[0] == 1.5047


class _FakeChecker:
    """Stands in for CollisionChecker so the seam logic can be tested without a GPU."""

    def __init__(self):
        self.batches = []

    def collision_check_paths(self, paths, obstacles):
        self.batches.append(len(paths))
        return [len(p[0]) % 2 == 0 for p in paths]

    def collision_check(self, path, obstacles):
        raise AssertionError("the seam must batch, not call per path")


def test_pool_seam_contract():
    import itertools
    fc = _FakeChecker()
    paths = [[[0.0] * n, [0.0] * n, [0.0] * n] for n in (1, 2, 3, 4, 5, 6, 7)]
    obstacle = [[1.0, 2.0]]
    out = mp.ThreadPool(processes=len(paths)).starmap(fc.collision_check, zip(paths, itertools.repeat(obstacle)))
    assert out == [False, True, False, True, False, True, False] and fc.batches == [7]
    with pytest.raises(ValueError):
        mp.ThreadPool(processes=0)           # multiprocessing.Pool semantics the planner relies on (:373-374)
    assert mp.ThreadPool(3).starmap(pow, [(2, 3), (3, 2)]) == [8, 9]
    assert mp.ThreadPool(3).map(abs, [-1, 2]) == [1, 2]
    with mp.ThreadPool(2) as pool:
        assert pool.starmap(fc.collision_check, []) == []


def test_install_rebinds_reference_seams():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference checkout not present (GPU box)")
    ref = ref_loader.load()
    orig = (ref.drive.VehicleModel, ref.local_planner.ThreadPool, ref.collision_checker.CollisionChecker)
    try:
        done = mp.install()
        assert len(done) == 3
        assert ref.drive.VehicleModel is mp.VehicleModel
        assert ref.local_planner.ThreadPool is mp.ThreadPool
        assert ref.collision_checker.CollisionChecker is mp.CollisionChecker
        # the unmodified LocalPlanner now constructs the GPU-backed checker
        lp = ref.local_planner.LocalPlanner(30, 7, 2, [-1.0, 1.0, 3.0], [1.5] * 3, 10, 1.0, 1.5, 2.0, 3.5)
        assert isinstance(lp._collision_checker, mp.CollisionChecker)
    finally:
        mp.uninstall()
    assert (ref.drive.VehicleModel, ref.local_planner.ThreadPool, ref.collision_checker.CollisionChecker) == orig


def test_norm2_probe_matches_numpy():
    mode = host_norm2_mode()
    rng = np.random.default_rng(3)
    for a, b in rng.uniform(-30, 30, (200, 2)):
        assert norm2_closed_form(float(a), float(b), mode) == float(np.linalg.norm([a, b]))


def test_shard_ranges_cover_exactly():
    for n in (0, 1, 7, 4096, 1 << 20, 1000003):
        for ws in (1, 2, 3, 4, 8):
            spans = [D.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_workloads_are_deterministic_and_shaped():
    s0, d, t = wl.config2_rollouts(B=1024, n_steps=500)
    s0b, _, _ = wl.config2_rollouts(B=1024, n_steps=500)
    assert np.array_equal(s0, s0b) and s0.shape == (12, 1024) and d.shape == (50, 1, 1024) and t.shape == d.shape
    assert s0[0].min() >= 5 and s0[0].max() <= 40 and np.all(s0[10:] == 0)
    w = wl.config3_lattice(P=64, M=500)
    assert w["px"].shape == (64, 49) and w["pyaw"].shape == (64, 49) and w["obstacles"].shape == (500, 2)
    x, y, yaw = wl.sample_spirals(np.array([0.01]), np.array([-0.02]), np.array([30.0]))
    assert x.shape == (1, 49) and yaw.shape == (1, 50)        # the reference's 49/50 length quirk
    sets, st, dl, tq, ps = wl.config5_sweep(n_sets=4, n_man=8)
    assert st.shape == (12, 32) and ps.tolist() == sorted(ps.tolist()) and dl.shape == (1, 1, 32)
    c4 = wl.config4_mpc(B=16, n_steps=10)
    assert c4["cost_ref"].shape == (10, 2) and c4["state0"].shape == (12,)
