"""Development check without a GPU: the *device* headers (csrc/vehicle_rhs.cuh), compiled for the host
by g++, reproduce the oracle.  This validates the re-derived algebra of the CUDA kernels (hoisted
constants, shared reciprocals, running-sum RK4, front-steer fast path) before spending GPU time; the
MUFU/Newton math paths themselves are only exercised by the ``-m gpu`` tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, rel_err
from oracle import c_oracle, planar_numpy as pn
from python_motionplanning_b200 import workloads as wl


@pytest.fixture(scope="module")
def hostsim(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostsim") / "libhostsim.so"
    # HOSTSIM_FLAGS: extra -D tunables of the device headers (e.g. -DB200MP_MU_DEG5=1) for A/B checks on the CPU
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", *os.environ.get("HOSTSIM_FLAGS", "").split(),
                    os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp"), "-o", str(out)], check=True)
    return C.CDLL(str(out))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _run(lib, f32, s0, d, t, par, n, hold, stride, aux=False, mu=None, pset=None):
    B = s0.shape[1]
    n_out = n // stride
    traj, end = np.zeros((n_out, 10, B)), np.zeros((12, B))
    ax = np.zeros((n_out, 28, B)) if aux else None
    ps = None if pset is None else pset.ctypes.data_as(C.POINTER(C.c_int))
    lib.hostsim_rollout(f32, B, n, C.c_double(1e-4), hold, _ptr(s0), _ptr(d), d.shape[1], _ptr(t), t.shape[1], _ptr(mu),
                        par, ps, stride, _ptr(traj), _ptr(ax), _ptr(end))
    return traj, ax, end


def test_device_headers_match_oracle(hostsim):
    B, N = 256, 300
    s0, d, t = wl.config2_rollouts(B=B, n_steps=N)
    par = c_oracle.make_params(pn.VehicleParams())
    par[0].D[:] = (1.0,) * 4
    ref = c_oracle.rollout(s0, d, t, par, 1e-4, N, hold=10, store_stride=50, want_aux=True)
    traj, _, end = _run(hostsim, 0, s0, d, t, par, N, 10, 50)
    assert rel_err(traj, ref["traj"]).max() < 1e-13
    assert rel_err(end, ref["state_end"]).max() < 1e-11
    # generic 4-channel layout + logging outputs + per-rollout mu_max
    z = np.zeros_like(d)
    d4, t4 = np.ascontiguousarray(np.concatenate([d, d, z, z], 1)), np.ascontiguousarray(np.repeat(t, 4, 1))
    traj4, aux4, _ = _run(hostsim, 0, s0, d4, t4, par, N, 10, 50, aux=True, mu=np.ones((4, B)))
    assert rel_err(traj4, ref["traj"]).max() < 1e-13
    assert rel_err(aux4, ref["aux"]).max() < 1e-9
    # FP32 instantiation stays within the stated drift bound
    traj32, _, _ = _run(hostsim, 1, s0, d, t, par, N, 10, 50)
    assert rel_err(traj32, ref["traj"]).max() < 2e-3


def test_speculative_step_falls_back_out_of_range(hostsim):
    """Headings beyond the fast sincos range (|yaw| > 1e5) and yaw rates whose stage increments exceed the
    small-angle rotation (|h*wz| > 2^-10) must take the checked step and still match the oracle."""
    B, N = 64, 60
    s0, d, t = wl.config2_rollouts(B=B, n_steps=N)
    s0 = s0.copy()
    s0[7, :16] = 3.0e5 + np.arange(16)          # huge heading
    s0[2, 16:32] = 12.0                          # |h*wz| = 1.2e-3 > 2^-10
    s0[7, 32:40] = -2.5e5
    par = c_oracle.make_params(pn.VehicleParams())
    par[0].D[:] = (1.0,) * 4
    ref = c_oracle.rollout(s0, d, t, par, 1e-4, N, hold=10, store_stride=20)
    traj, _, end = _run(hostsim, 0, s0, d, t, par, N, 10, 20)
    assert rel_err(traj, ref["traj"]).max() < 1e-10      # sin/cos of 3e5: conditioning, not method
    assert rel_err(end, ref["state_end"]).max() < 1e-10


def test_tabulated_friction_path_matches_oracle(hostsim):
    """The fast-path kernels replace sqrt / reciprocal / atan / sin of the combined-slip friction by a host-built table of
    polynomials in 1 + (B s)^2 (vehicle_rhs.cuh: build_mu_table).  Compiled for the host, the same code must reproduce
    the oracle at rounding level, hand slips beyond the table to the closed form, and keep the zero-slip state exact."""
    hostsim.hostsim_set_table.restype = C.c_double
    B, N = 256, 300
    s0, d, t = wl.config2_rollouts(B=B, n_steps=N)
    s0 = s0.copy()
    s0[3:7, :8] *= 3.5                      # wheels spinning at 3.5x: slip 2.5 is beyond the table (B s ~ 52)
    s0[3:7, 8:16] = 0.0                     # locked wheels: slip -1
    par = c_oracle.make_params(pn.VehicleParams())
    par[0].D[:] = (1.0,) * 4
    ref = c_oracle.rollout(s0, d, t, par, 1e-4, N, hold=10, store_stride=50)
    hostsim.hostsim_set_table(1)
    try:
        traj, _, end = _run(hostsim, 0, s0, d, t, par, N, 10, 50)
        err = hostsim.hostsim_set_table(1)
        assert 0.0 < err < 1e-15, err       # the builder's own audit against long double
        assert rel_err(traj, ref["traj"]).max() < 1e-12
        assert rel_err(end, ref["state_end"]).max() < 1e-11
        # zero-slip equilibrium stays exactly stationary (reference: state_dot = [0, .., 0, 25, 0])
        z = np.zeros((12, 1))
        z[0], z[3:7] = 25.0, 25.0 / 0.308309813617345
        zt, _, zend = _run(hostsim, 0, z, np.zeros((1, 1, 1)), np.zeros((1, 1, 1)), par, 10, 10, 10)
        assert zend[0, 0] == 25.0 and np.all(zend[1:3, 0] == 0.0) and np.all(zend[3:7, 0] == z[3, 0]) and zend[9, 0] == 0.0
        # per-wheel D (mu_max) and a parameter sweep with D != 1: the tables are normalised to D = 1, D scales the normal load
        mu = np.random.default_rng(2).uniform(0.4, 1.1, (4, B))
        refm = c_oracle.rollout(s0, d, t, par, 1e-4, N, hold=10, mu=mu, store_stride=50)
        trajm, _, endm = _run(hostsim, 0, s0, d, t, par, N, 10, 50, mu=mu)
        assert rel_err(trajm, refm["traj"]).max() < 1e-12
        # the FP32 twin of the table: audited to a few FP32 ulps, trajectories inside the FP32 drift bound
        sub = [np.ascontiguousarray(a[..., 16:]) for a in (s0, d, t)]        # without the rollouts sent beyond the table
        traj32, _, end32 = _run(hostsim, 1, sub[0], sub[1], sub[2], par, N, 10, 50)
        hostsim.hostsim_table_err_f32.restype = C.c_double
        assert 0.0 < hostsim.hostsim_table_err_f32() < 1e-6
        assert rel_err(end32[:10], ref["state_end"][:10, 16:]).max() < 2e-4
    finally:
        hostsim.hostsim_set_table(0)
