"""ctypes binding of ``libb200mp.so`` (the C ABI declared in ``include/b200mp.h``).

There is deliberately no fallback: if the shared library is missing and cannot be built, or a call
returns a non-zero code, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200mp.so")

E_ARG, E_PARAMS, E_NODEVICE = -1, -2, -3
NORM2_NOFMA, NORM2_FMA_V1, NORM2_FMA_V0 = 0, 1, 2


class B200mpError(RuntimeError):
    """A libb200mp entry point failed (CUDA error, or no CUDA device: there is no CPU fallback)."""


class VehicleParamsC(C.Structure):
    """``B200mpVehicleParams``"""
    _fields_ = [(n, C.c_double) for n in ("m", "a", "b", "Izz", "Jw", "hg", "T", "wL", "wR", "rw")] + [
        ("B", C.c_double * 4), ("C", C.c_double * 4), ("D", C.c_double * 4)]


class RolloutArgsC(C.Structure):
    """``B200mpRolloutArgs``"""
    _fields_ = [
        ("B", C.c_int), ("n_steps", C.c_int), ("step0", C.c_int), ("hold", C.c_int), ("dt", C.c_double),
        ("state0", C.c_void_p), ("delta", C.c_void_p), ("torque", C.c_void_p),
        ("delta_ch", C.c_int), ("torque_ch", C.c_int), ("ctrl_broadcast", C.c_int), ("store_stride", C.c_int),
        ("mu", C.c_void_p), ("param_set", C.c_void_p), ("traj", C.c_void_p), ("aux", C.c_void_p),
        ("state_end", C.c_void_p), ("cost", C.c_void_p), ("cost_in", C.c_void_p), ("cost_ref", C.c_void_p),
        ("w_u", C.c_double), ("u_ref", C.c_double), ("state_broadcast", C.c_int), ("friction_override", C.c_int),
    ]


class TrackArgsC(C.Structure):
    """``B200mpTrackArgs``"""
    _fields_ = [(n, C.c_int) for n in ("V", "n_steps", "step0", "ctrl_every", "store_stride", "n_sets", "w_max",
                                       "vehicles_per_set", "norm_mode")] + [
        (n, C.c_double) for n in ("dt", "target_vel", "k", "k_soft", "max_steer", "kp", "ki", "kd", "lookahead", "deadband",
                                  "steer_filter")] + [
        (n, C.c_void_p) for n in ("state0", "ctrl0", "waypoints", "wp_count", "traj", "log", "target_idx", "state_end",
                                  "ctrl_end")] + [("friction_override", C.c_int)]


# every symbol include/b200mp.h declares: name -> (restype, argtypes)
_vp, _i, _d, _ll, _ull = C.c_void_p, C.c_int, C.c_double, C.c_longlong, C.c_ulonglong
_dp = C.POINTER(C.c_double)
PROTOTYPES = {
    "b200mp_version": (_i, []),
    "b200mp_last_error": (C.c_char_p, []),
    "b200mp_device_count": (_i, []),
    "b200mp_set_params": (_i, [_i, C.POINTER(VehicleParamsC), _i]),
    "b200mp_rk4_rollout_f64": (_i, [_i, _vp, C.POINTER(RolloutArgsC)]),
    "b200mp_rk4_rollout_f32": (_i, [_i, _vp, C.POINTER(RolloutArgsC)]),
    "b200mp_planar_model_f64": (_i, [_i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200mp_mpc_sample_controls_f64": (_i, [_i, _vp, _i, _i, _ull, _ll, _d, _d, _d, _d, _d, _vp, _vp]),
    "b200mp_argmin_f64": (_i, [_i, _vp, _ll, _vp, _ll, _vp, _vp]),
    "b200mp_mpc_winner_f64": (_i, [_i, _vp, _ll, _i, _vp, _vp, _vp, _ll, _vp]),
    "b200mp_collision_check_f64": (_i, [_i, _vp, _i, _i, _i, _dp, _dp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "b200mp_collision_check_yaw_f64": (_i, [_i, _vp, _i, _i, _i, _dp, _dp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _i]),
    "b200mp_collision_resolve_f64": (_i, [_i, _vp, _i, _vp, _vp, _i, _i, _i, _dp, _dp, _vp, _vp, _i, _vp, _vp]),
    "b200mp_set_friction_mode": (_i, [_i]),
    "b200mp_set_collision_mode": (_i, [_i]),
    "b200mp_collision_stats": (_i, [_i, _vp, _i, C.POINTER(C.c_ulonglong)]),
    "b200mp_select_best_f64": (_i, [_i, _vp, _i, _vp, _vp, _vp, _d, _d, _d, _i, _vp, _vp]),
    "b200mp_track_closed_loop_f64": (_i, [_i, _vp, C.POINTER(TrackArgsC)]),
    "b200mp_sample_lattice_f64": (_i, [_i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200mp_optimize_spirals_f64": (_i, [_i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200mp_fma_peak": (_i, [_i, _i, _i, _dp]),
    "b200mp_shutdown": (_i, []),
}

_lib = None


def load(build_if_missing: bool = True):
    """Load (once) and return the ctypes handle with every prototype set."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise B200mpError(f"{LIB_PATH} is missing; run `python -m python_motionplanning_b200.build`")
        from . import build as _build
        try:
            _build.build()
        except Exception as exc:  # pragma: no cover - depends on toolchain
            raise B200mpError(f"libb200mp.so is missing and could not be built: {exc}") from exc
    try:
        import torch  # noqa: F401  (loads the CUDA runtime the library shares with torch)
    except Exception:  # pragma: no cover
        pass
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError here = the .so does not match include/b200mp.h
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    msg = load().b200mp_last_error().decode("utf-8", "replace")
    kind = {E_ARG: "argument error", E_PARAMS: "parameter-table error", E_NODEVICE: "no CUDA device"}.get(
        rc, f"CUDA error {rc}" if rc > 0 else f"error {rc}")
    if rc == E_ARG:
        raise ValueError(f"{what}: {msg}")
    raise B200mpError(f"{what}: {kind}: {msg}")
