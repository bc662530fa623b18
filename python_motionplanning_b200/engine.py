"""Batch entry points of the B200 engine: thin Python over the C ABI (``include/b200mp.h``).

PyTorch appears here only as plumbing: CUDA tensors are the device buffers (``.data_ptr()``), the
current torch stream is the launch stream, pinned tensors are the host staging areas.  Every result is
computed by the hand-written kernels in ``csrc/``; there is no CPU fallback.

Array layouts are structure-of-arrays with the rollout / path index last (fastest):
``state[12, B]``, ``delta[n_seg, ch, B]``, ``traj[n_out, 10, B]``, ``px[P, n_pts]`` ...
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import B200mpError, RolloutArgsC, TrackArgsC, VehicleParamsC, check
from .host_numerics import host_norm2_mode

_PARAM_SCALARS = ("m", "a", "b", "Izz", "Jw", "hg", "T", "wL", "wR", "rw")
_WHEELS = ("FL", "FR", "RL", "RR")


def pack_params(p, n_sets: Optional[int] = None):
    """Pack objects carrying the reference's ``VehicleParameters`` attribute names into the C struct array.

    ``p`` may be one object (its ``B**/C**/D**`` scalars, or ``[n_sets]`` arrays for a tyre sweep) or a
    sequence of objects (one set each).
    """
    if isinstance(p, (list, tuple)):
        arr = (VehicleParamsC * len(p))()
        for k, q in enumerate(p):
            arr[k] = pack_params(q)[0]
        return arr
    cols = {}
    n = 1
    for w in _WHEELS:
        for c in "BCD":
            v = np.atleast_1d(np.asarray(getattr(p, c + w), dtype=np.float64))
            cols[c + w] = v
            n = max(n, len(v))
    n = n_sets or n
    arr = (VehicleParamsC * n)()
    for s in range(n):
        q = arr[s]
        for name in _PARAM_SCALARS:
            setattr(q, name, float(getattr(p, name)))
        for i, w in enumerate(_WHEELS):
            for c in "BCD":
                v = cols[c + w]
                getattr(q, c)[i] = float(v[s] if len(v) > 1 else v[0])
    return arr


# The parameter table is library state per DEVICE (not per Engine object): remember what was uploaded
# last so that several Engine / VehicleModel objects on one device never trust a stale table.
_UPLOADED = {}      # device -> (signature bytes, upload counter)


def uploaded_token(device: int) -> int:
    """Counter that changes whenever a new parameter table is uploaded to ``device``."""
    return _UPLOADED.get(device, (None, 0, 0))[1]


@dataclass
class TrackGains:
    """Controller constants with the reference's values: Stanley ``k, k_soft, max_steer`` (drive.py:71-74, :57),
    PID ``kp, ki, kd`` (drive.py:82-84), ``lookahead`` / ``deadband`` (stanley_controller.py:44-45) and the
    steering filter coefficient ``1e-5 / (2*0.001)`` (drive.py:137)."""
    k: float = 100.0
    k_soft: float = 1.0
    max_steer: float = float(np.deg2rad(30))
    kp: float = 1000.0
    ki: float = 100.0
    kd: float = 0.0
    lookahead: float = 5.0
    deadband: float = 0.01
    steer_filter: float = 1e-5 / (2 * 0.001)


@dataclass
class TrackResult:
    state_end: torch.Tensor                      # [12, V]
    ctrl_end: torch.Tensor                       # [3, V]  x_del, integral of speed error, previous speed
    traj: Optional[torch.Tensor] = None          # [n_out, 10, V]
    log: Optional[torch.Tensor] = None           # [n_out, 45, V]  the DataLog rows of drive.py:145-151
    target_idx: Optional[torch.Tensor] = None    # [n_ctrl, V] int32


@dataclass
class RolloutResult:
    state_end: torch.Tensor                 # [12, B]
    traj: Optional[torch.Tensor] = None     # [n_out, 10, B]
    aux: Optional[torch.Tensor] = None      # [n_out, 28, B]  state_dot(10) + outputs(18)
    cost: Optional[torch.Tensor] = None     # [B]


class Engine:
    """One CUDA device's view of libb200mp.  Construct one per GPU (one process per GPU under torchrun)."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        n = self.lib.b200mp_device_count()
        if n <= 0 or not torch.cuda.is_available():
            raise B200mpError("no CUDA device is visible: python_motionplanning_b200 has no CPU fallback "
                              f"(b200mp_device_count() = {n})")
        if not 0 <= device < n:
            raise ValueError(f"device {device} out of range (have {n})")
        self.device = int(device)
        self.tdev = torch.device("cuda", self.device)
        self._norm2_mode = None

    # ------------------------------------------------------------------ plumbing
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.tdev).cuda_stream

    def dev(self, x, dtype=torch.float64) -> torch.Tensor:
        """A contiguous device tensor of ``dtype`` holding ``x`` (numpy / list / tensor); no copy when possible."""
        if isinstance(x, torch.Tensor):
            if x.device != self.tdev or x.dtype != dtype or not x.is_contiguous():
                x = x.to(device=self.tdev, dtype=dtype).contiguous()
            return x
        a = np.ascontiguousarray(x, dtype={torch.float64: np.float64, torch.float32: np.float32,
                                           torch.int32: np.int32, torch.uint8: np.uint8}[dtype])
        return torch.from_numpy(a).to(self.tdev)

    def empty(self, *shape, dtype=torch.float64) -> torch.Tensor:
        return torch.empty(*shape, dtype=dtype, device=self.tdev)

    @staticmethod
    def _ptr(t: Optional[torch.Tensor]):
        return None if t is None else C.c_void_p(t.data_ptr())

    # ---------------------------------------------------------------- parameters
    def set_params(self, p) -> int:
        """Upload parameter set(s) (reference ``VehicleParameters`` objects); returns the number of sets."""
        arr = p if isinstance(p, C.Array) else pack_params(p)
        sig = bytes(arr)
        old_sig, token = _UPLOADED.get(self.device, (None, 0, 0))[:2]
        if sig != old_sig:
            check(self.lib.b200mp_set_params(self.device, arr, len(arr)), "b200mp_set_params")
            _UPLOADED[self.device] = (sig, token + 1, len(arr))
        return len(arr)

    # ------------------------------------------------------------------ rollouts
    def rollout(self, state0, delta, torque, dt: float, n_steps: int, hold: int = 1, mu=None, param_set=None,
                store_stride: int = 0, want_aux: bool = False, dtype: str = "f64", ctrl_broadcast: bool = False,
                cost_ref=None, cost_in=None, w_u: float = 0.1, u_ref: float = 25.0, step0: int = 0,
                traj_out: Optional[torch.Tensor] = None, state_out: Optional[torch.Tensor] = None,
                aux_out: Optional[torch.Tensor] = None, state_broadcast: bool = False, friction: Optional[str] = None,
                cost_out: Optional[torch.Tensor] = None) -> RolloutResult:
        """Batched open-loop RK4 rollouts (``b200mp_rk4_rollout_f64/_f32``), asynchronous on the current stream.

        state0 ``[12,B]`` (or ``[10,B]``: ax_prev = ay_prev = 0, drive.py:60-61); delta ``[n_seg,1|4,B]``;
        torque ``[n_seg,1|4,B]`` (``[n_seg,ch,1]`` with ``ctrl_broadcast``); segment of step n = (step0+n)//hold.
        ``state_broadcast``: state0 is ``[12,1]``, one start state for all B rollouts (B is taken from the controls).
        ``friction`` = ``"auto" | "closed_form"`` for THIS launch (default: the process-wide mode).
        """
        td = torch.float64 if dtype == "f64" else torch.float32
        fn = self.lib.b200mp_rk4_rollout_f64 if dtype == "f64" else self.lib.b200mp_rk4_rollout_f32
        s0 = self.dev(state0, td)
        if s0.dim() != 2 or s0.shape[0] not in (10, 12):
            raise ValueError(f"state0 must be [12,B] or [10,B], got {tuple(s0.shape)}")
        B = s0.shape[1]
        if s0.shape[0] == 10:
            s0 = torch.cat([s0, torch.zeros(2, B, dtype=td, device=self.tdev)])
        dl, tq = self.dev(delta, td), self.dev(torque, td)
        if dl.dim() != 3 or tq.dim() != 3:
            raise ValueError("delta and torque must be [n_seg, channels, B]")
        if state_broadcast:
            if B != 1 or ctrl_broadcast:
                raise ValueError("state_broadcast takes state0 [12,1] and per-rollout controls")
            B = dl.shape[2]
        if friction not in (None, "auto", "closed_form"):
            raise ValueError("friction must be 'auto' or 'closed_form'")
        need_seg = -(-(step0 + n_steps) // hold) if n_steps else 0
        cb = 1 if ctrl_broadcast else B
        if dl.shape[0] < need_seg or tq.shape[0] < need_seg or dl.shape[2] != cb or tq.shape[2] != cb:
            raise ValueError(f"controls must cover {need_seg} segments x {cb} rollouts; got delta {tuple(dl.shape)}, "
                             f"torque {tuple(tq.shape)}")
        n_out = n_steps // store_stride if store_stride else 0
        traj = aux = None
        if n_out:
            if traj_out is not None or not (want_aux and aux_out is not None):
                traj = traj_out if traj_out is not None else self.empty(n_out, 10, B, dtype=td)
                if traj.shape != (n_out, 10, B) or traj.dtype != td or not traj.is_contiguous():
                    raise ValueError("traj_out has the wrong shape/dtype")
            if want_aux:
                aux = aux_out if aux_out is not None else self.empty(n_out, 28, B, dtype=td)
                if aux.shape != (n_out, 28, B) or aux.dtype != td or not aux.is_contiguous():
                    raise ValueError("aux_out has the wrong shape/dtype")
        end = state_out if state_out is not None else self.empty(12, B, dtype=td)
        mu_t = None if mu is None else self.dev(mu, td)
        self._check_param_set(param_set)
        ps_t = None if param_set is None else self.dev(param_set, torch.int32)
        cref = cost = cin = None
        if cost_ref is not None:
            cref = self.dev(cost_ref, td)
            if cref.shape[0] < step0 + n_steps:
                raise ValueError("cost_ref must have one (x, y) row per step")
            cost = cost_out if cost_out is not None else self.empty(B, dtype=td)
            if cost.shape != (B,) or cost.dtype != td or not cost.is_contiguous():
                raise ValueError("cost_out has the wrong shape/dtype")
            cin = None if cost_in is None else self.dev(cost_in, td)
        a = RolloutArgsC(B=B, n_steps=int(n_steps), step0=int(step0), hold=int(hold), dt=float(dt),
                         state0=s0.data_ptr(), delta=dl.data_ptr(), torque=tq.data_ptr(),
                         delta_ch=int(dl.shape[1]), torque_ch=int(tq.shape[1]), ctrl_broadcast=int(bool(ctrl_broadcast)),
                         store_stride=int(store_stride if n_out else 0),
                         mu=None if mu_t is None else mu_t.data_ptr(),
                         param_set=None if ps_t is None else ps_t.data_ptr(),
                         traj=None if traj is None else traj.data_ptr(), aux=None if aux is None else aux.data_ptr(),
                         state_end=end.data_ptr(), cost=None if cost is None else cost.data_ptr(),
                         cost_in=None if cin is None else cin.data_ptr(),
                         cost_ref=None if cref is None else cref.data_ptr(), w_u=float(w_u), u_ref=float(u_ref),
                         state_broadcast=int(bool(state_broadcast)),
                         friction_override={None: 0, "auto": 1, "closed_form": 2}[friction])
        check(fn(self.device, self._stream(), C.byref(a)), "b200mp_rk4_rollout_" + dtype)
        return RolloutResult(state_end=end, traj=traj, aux=aux, cost=cost)

    def _check_param_set(self, param_set):
        """Host-side range check of per-rollout parameter-set indices given as host data (a device tensor is not read back:
        the kernels clamp indices to the uploaded table, so an out-of-range entry can never read out of bounds)."""
        if param_set is None or isinstance(param_set, torch.Tensor) and param_set.is_cuda:
            return
        a = np.asarray(param_set)
        n = _UPLOADED.get(self.device, (None, 0, 0))[2]
        if a.size and n and (a.min() < 0 or a.max() >= n):
            raise ValueError(f"param_set entries must be in [0, {n}) (the uploaded table has {n} sets); got [{a.min()}, {a.max()}]")

    def rollout_to_host(self, state0_host: torch.Tensor, delta_host: torch.Tensor, torque_host: torch.Tensor,
                        dt: float, n_steps: int, hold: int, traj_host: torch.Tensor, chunk_steps: int = 50,
                        dtype: str = "f64", state_end_host: Optional[torch.Tensor] = None, n_slabs: int = 2):
        """End-to-end rollout with HOST buffers: H2D of the inputs, time-chunked kernels, D2H of the full
        trajectory overlapped with the next chunks (``n_slabs`` device slabs in a ring, two streams).

        ``*_host`` are pinned CPU tensors; ``traj_host`` is ``[n_steps, 10, B]``.  Rollouts are resumable
        (``state_end`` of one chunk is ``state0`` of the next), which is what makes the time split exact.
        Returns after the last byte has landed in ``traj_host``.
        """
        td = torch.float64 if dtype == "f64" else torch.float32
        B = state0_host.shape[1]
        if chunk_steps % hold != 0 and n_steps > chunk_steps:
            raise ValueError("chunk_steps must be a multiple of hold")
        compute = torch.cuda.current_stream(self.tdev)
        copy = self._copy_stream()
        s = state0_host.to(self.tdev, non_blocking=True)
        dl = delta_host.to(self.tdev, non_blocking=True)
        tq = torque_host.to(self.tdev, non_blocking=True)
        n_slabs = max(2, int(n_slabs))
        slabs = self._slabs(min(chunk_steps, n_steps), B, td, n_slabs)
        slab_free = [None] * n_slabs
        n0, k = 0, 0
        while n0 < n_steps:
            nc = min(chunk_steps, n_steps - n0)
            slab = slabs[k % n_slabs][:nc]
            if slab_free[k % n_slabs] is not None:
                compute.wait_event(slab_free[k % n_slabs])      # D2H of the chunk that used this slab is done
            res = self.rollout(s, dl, tq, dt, nc, hold=hold, store_stride=1, dtype=dtype, step0=n0, traj_out=slab)
            s = res.state_end
            done = torch.cuda.Event()
            done.record(compute)
            copy.wait_event(done)
            with torch.cuda.stream(copy):
                traj_host[n0:n0 + nc].copy_(slab, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            slab_free[k % n_slabs] = ev
            n0 += nc
            k += 1
        if state_end_host is not None:
            state_end_host.copy_(s, non_blocking=True)
        copy.synchronize()
        compute.synchronize()
        return s

    def rollout_endstate_to_host(self, state0_host: torch.Tensor, delta_host: torch.Tensor, torque_host: torch.Tensor,
                                 dt: float, n_steps: int, hold: int, state_end_host: torch.Tensor, chunk_steps: int = 100,
                                 dtype: str = "f64"):
        """End-to-end rollout with HOST buffers when only the final states are wanted (no trajectory readback):
        the controls are uploaded in time-chunks on the copy stream while the previous chunk computes, so the H2D of
        the inputs (the only sizeable transfer left) hides behind the kernels.  ``*_host`` are pinned CPU tensors;
        ``state_end_host`` is ``[12, B]``.  Returns after the end states have landed in ``state_end_host``."""
        if chunk_steps % hold != 0 and n_steps > chunk_steps:
            raise ValueError("chunk_steps must be a multiple of hold")
        compute = torch.cuda.current_stream(self.tdev)
        copy = self._copy_stream()
        seg_per_chunk = max(chunk_steps // hold, 1)
        s = state0_host.to(self.tdev, non_blocking=True)
        n_seg = delta_host.shape[0]
        dl = torch.empty(delta_host.shape, dtype=delta_host.dtype, device=self.tdev)
        tq = torch.empty(torque_host.shape, dtype=torque_host.dtype, device=self.tdev)
        start = torch.cuda.Event()
        start.record(compute)
        copy.wait_event(start)
        ready = []
        with torch.cuda.stream(copy):
            for s0 in range(0, n_seg, seg_per_chunk):
                s1 = min(n_seg, s0 + seg_per_chunk)
                dl[s0:s1].copy_(delta_host[s0:s1], non_blocking=True)
                tq[s0:s1].copy_(torque_host[s0:s1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
                ready.append(ev)
        n0 = 0
        while n0 < n_steps:
            nc = min(chunk_steps, n_steps - n0)
            # the launch reads control segments up to ceil((n0 + nc) / hold) - 1: wait for the upload that covers the last one
            last_seg = -(-(n0 + nc) // hold) - 1
            compute.wait_event(ready[min(last_seg // seg_per_chunk, len(ready) - 1)])
            s = self.rollout(s, dl, tq, dt, nc, hold=hold, store_stride=0, dtype=dtype, step0=n0).state_end
            n0 += nc
        state_end_host.copy_(s, non_blocking=True)
        compute.synchronize()
        copy.synchronize()          # uploads of segments beyond the last step must not outlive dl / tq
        return s

    def _copy_stream(self):
        if not hasattr(self, "_cstream"):
            self._cstream = torch.cuda.Stream(self.tdev)
        return self._cstream

    def _slabs(self, steps, B, td, n=2):
        key = (steps, B, td, n)
        if getattr(self, "_slab_key", None) != key:
            self._slab = None
            self._slab = [self.empty(steps, 10, B, dtype=td) for _ in range(n)]
            self._slab_key = key
        return self._slab

    def pinned_empty(self, *shape, dtype=torch.float64, numa_local: bool = True) -> torch.Tensor:
        """Page-locked host tensor for the ``*_to_host`` calls, placed on this GPU's NUMA node when the host allows it."""
        from . import hostmem
        return hostmem.pinned_empty(*shape, dtype=dtype, numa_node=hostmem.gpu_numa_node(self.device) if numa_local else -1)

    def planar_model_batch(self, state, torque, mu, delta, ax_prev, ay_prev, param_set=None, axay=None, out=None):
        """Batched ``VehicleModel.planar_model``: returns ``(state_dot[10,B], misc[6,B], outputs[18,B])``.
        ``axay`` (a ready ``[2,B]`` device tensor, instead of ``ax_prev, ay_prev``) and ``out`` (three preallocated
        device tensors) let a caller in a latency-bound loop skip the staging work."""
        st = self.dev(state)
        B = st.shape[1]
        tq, dl = self.dev(torque), self.dev(delta)
        mu_t = None if mu is None else self.dev(mu)
        if axay is None:
            axay = torch.stack([self.dev(ax_prev).reshape(B), self.dev(ay_prev).reshape(B)]).contiguous()
        self._check_param_set(param_set)
        ps = None if param_set is None else self.dev(param_set, torch.int32)
        sd, misc, out = out if out is not None else (self.empty(10, B), self.empty(6, B), self.empty(18, B))
        check(self.lib.b200mp_planar_model_f64(self.device, self._stream(), B, self._ptr(st), self._ptr(tq),
                                               self._ptr(mu_t), self._ptr(dl), self._ptr(axay), self._ptr(ps),
                                               self._ptr(sd), self._ptr(misc), self._ptr(out)), "b200mp_planar_model_f64")
        return sd, misc, out

    # ------------------------------------------------------------ closed loop
    def track_closed_loop(self, state0, waypoints, dt: float, n_steps: int, target_vel: float = 25.0,
                          gains: Optional[TrackGains] = None, ctrl0=None, wp_count=None, vehicles_per_set: Optional[int] = None,
                          ctrl_every: int = 10, store_stride: int = 0, want_log: bool = False, want_target_idx: bool = False,
                          step0: int = 0, norm_mode: Optional[int] = None, friction: Optional[str] = None) -> TrackResult:
        """Batched Stanley + PID closed loop around the RK4 step (``b200mp_track_closed_loop_f64``): what
        ``Car.drive`` does between two planner calls (drive.py:126-151), for V vehicles at once.

        state0 ``[12,V]`` (or ``[10,V]``); waypoints ``[n_sets,W,2]`` or ``[W,2]`` (one shared list), ``wp_count``
        the used length per set; vehicle r tracks set ``r // vehicles_per_set``.  ``ctrl0 [3,V]`` = steering-filter
        state, integral of the speed error, previous speed (default ``[0, 0, U0]``, drive.py:47-54).  The uploaded
        parameter set's ``D`` plays ``mu_max`` (set it to 1.0 for drive.py:142).  Asynchronous on the current stream."""
        g = gains or TrackGains()
        s0 = self.dev(state0)
        if s0.dim() != 2 or s0.shape[0] not in (10, 12):
            raise ValueError(f"state0 must be [12,V] or [10,V], got {tuple(s0.shape)}")
        V = s0.shape[1]
        if s0.shape[0] == 10:
            s0 = torch.cat([s0, torch.zeros(2, V, dtype=torch.float64, device=self.tdev)])
        wp = self.dev(waypoints)
        if wp.dim() == 2:
            wp = wp[None]
        if wp.dim() != 3 or wp.shape[2] != 2:
            raise ValueError("waypoints must be [n_sets, W, 2] or [W, 2]")
        wp = wp.contiguous()
        n_sets, W = wp.shape[0], wp.shape[1]
        cnt = self.dev(wp_count if wp_count is not None else [W] * n_sets, torch.int32)
        if cnt.numel() != n_sets:
            raise ValueError("wp_count must have one entry per waypoint set")
        if wp_count is not None and not isinstance(wp_count, torch.Tensor):   # host data: check it here (the kernel clamps)
            c = np.asarray(wp_count)
            if c.size and (c.min() < 0 or c.max() > W):
                raise ValueError(f"wp_count entries must lie in [0, {W}]")
        vps = int(vehicles_per_set or max(1, -(-V // n_sets)))
        if ctrl0 is None:
            c0 = torch.zeros(3, V, dtype=torch.float64, device=self.tdev)
            c0[2] = s0[0]
        else:
            c0 = self.dev(ctrl0)
            if c0.shape != (3, V):
                raise ValueError("ctrl0 must be [3, V]")
        n_out = n_steps // store_stride if store_stride else 0
        traj = self.empty(n_out, 10, V) if n_out else None
        log = self.empty(n_out, 45, V) if (n_out and want_log) else None
        n_ctrl = -(-n_steps // ctrl_every) if n_steps else 0
        tid = self.empty(n_ctrl, V, dtype=torch.int32) if (want_target_idx and n_ctrl) else None
        end, cend = self.empty(12, V), self.empty(3, V)
        if norm_mode is None:
            if self._norm2_mode is None:
                self._norm2_mode = host_norm2_mode()
            norm_mode = self._norm2_mode
        p = lambda t: None if t is None else t.data_ptr()
        a = TrackArgsC(V=V, n_steps=int(n_steps), step0=int(step0), ctrl_every=int(ctrl_every),
                       store_stride=int(store_stride if n_out else 0), n_sets=n_sets, w_max=W, vehicles_per_set=vps,
                       norm_mode=int(norm_mode), dt=float(dt), target_vel=float(target_vel), k=g.k, k_soft=g.k_soft,
                       max_steer=g.max_steer, kp=g.kp, ki=g.ki, kd=g.kd, lookahead=g.lookahead, deadband=g.deadband,
                       steer_filter=g.steer_filter, state0=p(s0), ctrl0=p(c0), waypoints=p(wp), wp_count=p(cnt),
                       traj=p(traj), log=p(log), target_idx=p(tid), state_end=p(end), ctrl_end=p(cend),
                       friction_override={None: 0, "auto": 1, "closed_form": 2}[friction])
        check(self.lib.b200mp_track_closed_loop_f64(self.device, self._stream(), C.byref(a)), "b200mp_track_closed_loop_f64")
        return TrackResult(state_end=end, ctrl_end=cend, traj=traj, log=log, target_idx=tid)

    # ------------------------------------------------------------- sampling MPC
    def mpc_sample_controls(self, B: int, n_seg: int, seed: int, rollout0: int = 0, delta_mean=0.0, delta_sigma=0.02,
                            delta_clip=0.5235987755982988, torque_mean=0.0, torque_sigma=50.0):
        delta, torque = self.empty(n_seg, 1, B), self.empty(n_seg, 1, B)
        check(self.lib.b200mp_mpc_sample_controls_f64(self.device, self._stream(), B, n_seg, int(seed), int(rollout0),
                                                      delta_mean, delta_sigma, delta_clip, torque_mean, torque_sigma,
                                                      self._ptr(delta), self._ptr(torque)), "b200mp_mpc_sample_controls_f64")
        return delta, torque

    def mpc_sample_controls_into(self, delta: torch.Tensor, torque: torch.Tensor, seed: int, rollout0: int = 0, delta_mean=0.0,
                                 delta_sigma=0.02, delta_clip=0.5235987755982988, torque_mean=0.0, torque_sigma=50.0):
        """``mpc_sample_controls`` into caller-owned ``[n_seg,1,B]`` buffers (no allocation: resident MPC planners)."""
        n_seg, _, B = delta.shape
        check(self.lib.b200mp_mpc_sample_controls_f64(self.device, self._stream(), B, n_seg, int(seed), int(rollout0),
                                                      delta_mean, delta_sigma, delta_clip, torque_mean, torque_sigma,
                                                      self._ptr(delta), self._ptr(torque)), "b200mp_mpc_sample_controls_f64")

    def mpc_winner(self, cost: torch.Tensor, delta: torch.Tensor, torque: torch.Tensor, index_offset: int = 0,
                   record_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Winner record ``[min cost, global index, delta[n_seg], torque[n_seg]]`` of a sampled batch in one call
        (``b200mp_mpc_winner_f64``: lowest-index argmin + gather of the winner's controls); device tensor, async."""
        n_seg, B = delta.shape[0], delta.shape[-1]
        rec = record_out if record_out is not None else self.empty(2 + 2 * n_seg)
        check(self.lib.b200mp_mpc_winner_f64(self.device, self._stream(), B, n_seg, self._ptr(cost), self._ptr(delta),
                                             self._ptr(torque), int(index_offset), self._ptr(rec)), "b200mp_mpc_winner_f64")
        return rec

    def argmin(self, cost: torch.Tensor, index_offset: int = 0):
        """Device-side lowest-index argmin; returns device tensors ``(min[1] f64, idx[1] i64)`` (async)."""
        cost = self.dev(cost)
        mn, ix = self.empty(1), self.empty(1, dtype=torch.int64)
        check(self.lib.b200mp_argmin_f64(self.device, self._stream(), cost.numel(), self._ptr(cost), int(index_offset),
                                         self._ptr(mn), self._ptr(ix)), "b200mp_argmin_f64")
        return mn, ix

    # ---------------------------------------------------------------- collision
    def path_trig(self, pyaw, n: int):
        """cos / sin of the first ``n`` yaws of every path, evaluated on the HOST with numpy (what the reference
        does, collision_checker.py:88-89) and uploaded: pass the pair as ``trig=`` to reuse it across calls."""
        if isinstance(pyaw, torch.Tensor):
            pyaw = pyaw.detach().cpu().numpy()
        yaw = np.asarray(pyaw, dtype=np.float64)[:, :n]
        return self.dev(np.cos(yaw)), self.dev(np.sin(yaw))

    def _clearance_from_yaws(self, px, py, pyaw, obstacles, offsets, radii, mode):
        """Exact minimum clearance ``[P]`` with only the CANDIDATE yaws evaluated on the host (see ``collision_check_batch``);
        ``None`` when the shortcut does not apply (non-finite values, no obstacles, too many candidates).

        The kernels are the existing ones, called on one-point "paths": (1) every path point with the device's ``sincos``;
        (2) with ``delta_j`` a bound on |device-trig clearance - host-trig clearance| of point j (the centre moves by at most a
        few ulp of its coordinates), ``U = min_j (c_j + delta_j)`` bounds the exact path minimum from above and a point with
        ``c_j - delta_j > U`` cannot attain it; (3) the remaining points are re-evaluated with numpy ``cos`` / ``sin`` of their
        yaws, which is the oracle's own arithmetic; (4) per-path minimum of those."""
        pxt, pyt, yaw_t = self.dev(px), self.dev(py), self.dev(pyaw)
        P, n = pxt.shape
        obs = self.dev(np.asarray(obstacles, dtype=np.float64).reshape(-1, 2) if not isinstance(obstacles, torch.Tensor)
                       else obstacles)
        if P == 0 or n == 0 or obs.shape[0] == 0 or yaw_t.dim() != 2 or yaw_t.shape[0] != P or yaw_t.shape[1] < n:
            return None
        yaw = yaw_t[:, :n].contiguous()
        flat = lambda t: t.contiguous().reshape(-1, 1)
        _, cpt = self.collision_check_batch(flat(pxt), flat(pyt), flat(yaw), obs, offsets, radii, want_clearance=True,
                                            device_trig=True, mode=mode)
        cpt = cpt.view(P, n)
        scale = pxt.abs() + pyt.abs() + (obs.abs().max() + max(abs(float(o)) for o in offsets))   # stays on the device
        delta = 1.5e-14 * scale                      # ~64 ulp of the coordinates involved: generous, and still ~1e-12 m
        upper = (cpt + delta).amin(dim=1, keepdim=True)
        cand = (cpt - delta) <= upper
        finite = (torch.isfinite(cpt).all() & torch.isfinite(delta).all()).to(torch.float64).view(1)
        idx = cand.nonzero()                         # (synchronises: the candidate count)
        K = idx.shape[0]
        if K == 0 or K > 4 * P + 1024:
            return None
        ip, ij = idx[:, 0], idx[:, 1]
        host = torch.cat([yaw[ip, ij], finite]).cpu().numpy()     # one transfer: the candidate yaws and the finiteness flag
        if host[-1] != 1.0:
            return None
        yc = host[:-1]
        trig = (self.dev(np.cos(yc)).view(-1, 1), self.dev(np.sin(yc)).view(-1, 1))   # numpy on the host: the oracle's values
        _, cex = self.collision_check_batch(pxt[ip, ij].view(-1, 1), pyt[ip, ij].view(-1, 1), None, obs, offsets, radii,
                                            want_clearance=True, trig=trig, mode=mode)
        out = torch.full((P,), float("inf"), dtype=torch.float64, device=self.tdev)
        out.scatter_reduce_(0, ip, cex, reduce="amin")
        self.last_clearance_candidates = int(K)
        return out

    _UNDECIDED_CAPACITY = 8192
    _COLLISION_MODES = {None: -1, "auto": 0, "fp64": 1, "screen": 2}

    def collision_check_batch(self, px, py, pyaw, obstacles, offsets: Sequence[float], radii: Sequence[float],
                              want_clearance: bool = False, device_trig: bool = False, trig=None, host_trig: bool = False,
                              mode: Optional[str] = None, clearance_trig: str = "auto"):
        """``free[P]`` (uint8, 1 = collision-free) for P paths at once.

        px, py ``[P,n]``; pyaw ``[P,>=n]`` (first n used).  Default: the yaws go to the device
        (``b200mp_collision_check_yaw_f64``); every verdict the kernel writes is proven to equal the reference's, whose
        circle centres carry numpy's ``cos`` / ``sin`` roundings (collision_checker.py:88-89), and the few path points it
        cannot prove -- an obstacle point within ~1e-13 m of a circle -- are decided with host-evaluated numpy
        ``cos`` / ``sin`` of exactly those yaws (``b200mp_collision_resolve_f64``).  The flags are bit-exact and only a
        4-byte count crosses the bus on the way back.  ``trig=(cos, sin)`` uses the caller's values as they are;
        ``host_trig=True`` evaluates all yaws with numpy on the host (the former default; what ``want_clearance`` uses, because
        the minimum clearance is a double that depends on every centre's last bit); ``device_trig=True`` is the unproven
        device-only path (kept for A/B).  ``mode`` = ``"auto" | "screen" | "fp64"`` for this call (default: the
        process-wide mode).  ``self.last_collision_undecided`` = number of host-resolved path points of the last call.

        ``want_clearance=True`` also returns the minimum clearance ``[P]`` (an extension: the reference returns booleans only;
        the oracle defines it with numpy's ``cos`` / ``sin``, and it is a double that depends on the last bit of every centre).
        ``clearance_trig="auto"`` (default): the clearance of every path POINT is first evaluated with the device's ``sincos``
        (within a proven ``delta`` of the host-trig value), only the points that can still hold their path's minimum -- usually
        one per path -- get numpy ``cos`` / ``sin`` on the host and an exact re-evaluation, and the result equals the oracle's
        double bit for bit; non-finite inputs or too many candidates fall back to ``"host"``.  ``"host"`` evaluates every yaw
        with numpy (2.9 ms on config 3 instead of ~1 ms).  ``"device"`` keeps everything on the device: the clearance then
        differs from the oracle's by at most a few 1e-14 m (the circle centres move by an ulp).  The FLAGS are the proven
        bit-exact ones in all three.
        """
        if want_clearance and clearance_trig not in ("auto", "host", "device"):
            raise ValueError("clearance_trig must be 'auto', 'host' or 'device'")
        if want_clearance and clearance_trig != "host" and trig is None and not host_trig and not device_trig:
            free = self.collision_check_batch(px, py, pyaw, obstacles, offsets, radii, mode=mode)
            undecided = self.last_collision_undecided          # of the flags call: the calls below must not overwrite it
            if clearance_trig == "device":
                _, clr = self.collision_check_batch(px, py, pyaw, obstacles, offsets, radii, want_clearance=True, device_trig=True,
                                                    mode=mode)
            else:
                clr = self._clearance_from_yaws(px, py, pyaw, obstacles, offsets, radii, mode)
            if clr is not None:
                self.last_collision_undecided = undecided
                return free, clr
        pxt, pyt = self.dev(px), self.dev(py)
        if pxt.dim() != 2:
            raise ValueError("px, py must be [P, n_pts]")
        P, n = pxt.shape
        obs = self.dev(np.asarray(obstacles, dtype=np.float64).reshape(-1, 2) if not isinstance(obstacles, torch.Tensor)
                       else obstacles)
        M = obs.shape[0]
        off = (C.c_double * len(offsets))(*[float(v) for v in offsets])
        rad = (C.c_double * len(radii))(*[float(v) for v in radii])
        if len(offsets) != len(radii):
            raise ValueError("circle_offsets and circle_radii must have the same length")
        if mode not in self._COLLISION_MODES:
            raise ValueError(f"collision mode must be one of {sorted(k for k in self._COLLISION_MODES if k)}")
        self.last_collision_undecided = 0
        free = self.empty(P, dtype=torch.uint8)
        exact_yaw = trig is None and not host_trig and not device_trig and not want_clearance
        if exact_yaw:
            yaw_t = self.dev(pyaw)
            if yaw_t.dim() != 2 or yaw_t.shape[0] != P or yaw_t.shape[1] < n:
                raise ValueError("pyaw must be [P, >= n_pts]")
            cap = self._UNDECIDED_CAPACITY
            if getattr(self, "_und", None) is None:
                self._und = self.empty(1 + cap, dtype=torch.int32)
                self._und_host = torch.empty(1, dtype=torch.int32).pin_memory()
            check(self.lib.b200mp_collision_check_yaw_f64(self.device, self._stream(), P, n, len(offsets), off, rad,
                                                          self._ptr(pxt), self._ptr(pyt), self._ptr(yaw_t), yaw_t.shape[1], M,
                                                          self._ptr(obs), self._ptr(free), self._ptr(self._und), cap,
                                                          self._COLLISION_MODES[mode]), "b200mp_collision_check_yaw_f64")
            self._und_host.copy_(self._und[:1], non_blocking=True)
            torch.cuda.current_stream(self.tdev).synchronize()
            count = int(self._und_host[0])
            self.last_collision_undecided = count
            if count == 0:
                return free
            if count <= cap:
                items = torch.unique(self._und[1:1 + count])
                jj, pp = items % n, items // n
                yaws = yaw_t[pp.long(), jj.long()].cpu().numpy()
                cs_host = np.stack([np.cos(yaws), np.sin(yaws)])          # numpy on the host: the reference's own values
                cs_dev = self.dev(cs_host)
                check(self.lib.b200mp_collision_resolve_f64(self.device, self._stream(), int(items.numel()), self._ptr(items),
                                                            self._ptr(cs_dev), P, n, len(offsets), off, rad, self._ptr(pxt),
                                                            self._ptr(pyt), M, self._ptr(obs), self._ptr(free)),
                      "b200mp_collision_resolve_f64")
                return free
            host_trig = True        # more undecided points than the list holds: evaluate every yaw on the host
        pc = ps = yaw_t = None
        stride = 0
        if trig is not None:
            pc, ps = self.dev(trig[0]), self.dev(trig[1])
            if pc.shape != (P, n) or ps.shape != (P, n):
                raise ValueError("trig must be (cos[P,n], sin[P,n])")
        elif device_trig and not host_trig:
            yaw_t = self.dev(pyaw)
            stride = yaw_t.shape[1]
        else:
            pc, ps = self.path_trig(pyaw, n)
        clr = self.empty(P) if want_clearance else None
        prev = None
        if mode is not None:
            prev = self.set_collision_mode(mode)
        try:
            check(self.lib.b200mp_collision_check_f64(self.device, self._stream(), P, n, len(offsets), off, rad,
                                                      self._ptr(pxt), self._ptr(pyt), self._ptr(pc), self._ptr(ps),
                                                      self._ptr(yaw_t), stride, M, self._ptr(obs), self._ptr(free),
                                                      self._ptr(clr)), "b200mp_collision_check_f64")
        finally:
            if prev is not None:
                self.set_collision_mode(prev)
        return (free, clr) if want_clearance else free

    def sample_lattice(self, kappa1, kappa2, sf, ego=None, n_samples: int = 50, want_trig: bool = True,
                       want_end: bool = True):
        """Spiral sampling + frame transform on the device (``b200mp_sample_lattice_f64``): for P optimisation
        parameter triples ``[p1, p2, sf]`` what ``PathOptimizer.sample_spiral`` (path_optimizer.py:131-174) followed by
        ``transform_paths`` (local_planner.py:424-470) return.  ``ego`` = ``(x, y, yaw)`` scalars (one pose for the
        whole lattice), ``[3,P]`` per path, or ``None`` for ego-frame output.  Returns a dict of device tensors
        ``px, py, pyaw [P, n_samples-1]`` (+ ``pcos, psin`` and ``end_xy [2,P]``), ready for
        ``collision_check_batch(px, py, None, obstacles, ..., trig=(pcos, psin))``."""
        k1, k2, s_f = self.dev(kappa1).reshape(-1), self.dev(kappa2).reshape(-1), self.dev(sf).reshape(-1)
        P = k1.numel()
        if k2.numel() != P or s_f.numel() != P:
            raise ValueError("kappa1, kappa2 and sf must have the same length")
        exs = eys = eyw = None
        bcast = 0
        if ego is not None:
            e = self.dev(np.asarray(ego, dtype=np.float64) if not isinstance(ego, torch.Tensor) else ego)
            if e.numel() == 3:
                e = e.reshape(3, 1)
                bcast = 1
            if e.shape[0] != 3 or e.shape[1] not in (1, P):
                raise ValueError("ego must be (x, y, yaw) or [3, P]")
            e = e.contiguous()
            exs, eys, eyw = e[0], e[1], e[2]
        n = n_samples - 1
        out = {"px": self.empty(P, n), "py": self.empty(P, n), "pyaw": self.empty(P, n)}
        if want_trig:
            out["pcos"], out["psin"] = self.empty(P, n), self.empty(P, n)
        if want_end:
            out["end_xy"] = self.empty(2, P)
        check(self.lib.b200mp_sample_lattice_f64(self.device, self._stream(), P, int(n_samples), self._ptr(k1), self._ptr(k2),
                                                 self._ptr(s_f), self._ptr(exs), self._ptr(eys), self._ptr(eyw), bcast,
                                                 self._ptr(out["px"]), self._ptr(out["py"]), self._ptr(out["pyaw"]),
                                                 self._ptr(out.get("pcos")), self._ptr(out.get("psin")),
                                                 self._ptr(out.get("end_xy"))), "b200mp_sample_lattice_f64")
        return out

    def optimize_spirals(self, xf, yf, tf, n_samples: int = 50):
        """Batched ``PathOptimizer.optimize_spiral`` (path_optimizer.py:31-88) for P goal states in the vehicle frame
        (``b200mp_optimize_spirals_f64``).  Returns device tensors ``p [3,P]`` (p1, p2, sf), ``objective [P]``,
        ``iterations [P]`` (int32) and ``valid [P]`` (uint8: the planner's acceptance test, local_planner.py:317-323)."""
        x, y, t = self.dev(xf).reshape(-1), self.dev(yf).reshape(-1), self.dev(tf).reshape(-1)
        P = x.numel()
        if y.numel() != P or t.numel() != P:
            raise ValueError("xf, yf and tf must have the same length")
        out = {"p": self.empty(3, P), "objective": self.empty(P), "iterations": self.empty(P, dtype=torch.int32),
               "valid": self.empty(P, dtype=torch.uint8)}
        check(self.lib.b200mp_optimize_spirals_f64(self.device, self._stream(), P, int(n_samples), self._ptr(x), self._ptr(y),
                                                   self._ptr(t), self._ptr(out["p"]), self._ptr(out["objective"]),
                                                   self._ptr(out["iterations"]), self._ptr(out["valid"])),
              "b200mp_optimize_spirals_f64")
        return out

    def plan_lattice(self, goals_local, ego, obstacles, offsets: Sequence[float], radii: Sequence[float], goal_xy,
                     weight: float, n_samples: int = 50):
        """The planner's numeric core on the device, end to end (local_planner.py:367-379): optimise one spiral per
        goal state (vehicle frame ``goals_local [3,P]`` = xf, yf, tf), sample and transform the spirals with ``ego`` =
        (x, y, yaw), test them against ``obstacles [M,2]`` and select the best path index.

        A path whose optimisation fails the planner's acceptance test is DROPPED, as ``plan_paths`` does
        (local_planner.py:317-323): it is neither a candidate nor a penalty term of ``select_best_path_index``
        (``free`` value 2 = excluded).  Returns ``(best, out)``: ``best`` indexes the FULL goal list (the device arrays in
        ``out`` are full-size); ``out["best_filtered"]`` is the same path's index in the reference's filtered ``paths``
        list (what ``MotionPlanner`` returns, :379-421); both ``None`` when nothing is free."""
        g = self.dev(goals_local)
        if g.dim() != 2 or g.shape[0] != 3:
            raise ValueError("goals_local must be [3, P]")
        opt = self.optimize_spirals(g[0], g[1], g[2], n_samples)
        lat = self.sample_lattice(opt["p"][0], opt["p"][1], opt["p"][2], ego=ego, n_samples=n_samples, want_trig=False)
        free = self.collision_check_batch(lat["px"], lat["py"], lat["pyaw"], obstacles, offsets, radii)
        state = torch.where(opt["valid"] != 0, free, torch.full_like(free, 2))
        best = self.select_best_path_index_batch(lat["end_xy"][0], lat["end_xy"][1], state, goal_xy, weight)
        lat.update(opt)
        lat["free"] = free & opt["valid"]
        lat["select_state"] = state
        lat["best_filtered"] = None if best is None else int(opt["valid"][:best].sum().item())
        return best, lat

    def set_friction_mode(self, mode: str) -> str:
        """``"auto"`` (default: the fast-path FP64 kernels evaluate the combined-slip friction from the host-built
        polynomial table of the uploaded tyre) or ``"closed_form"`` (always sqrt / atan / sin).  Process-wide; returns
        the previous mode."""
        names = {"auto": 0, "closed_form": 1}
        if mode not in names:
            raise ValueError(f"friction mode must be one of {sorted(names)}")
        prev = self.lib.b200mp_set_friction_mode(names[mode])
        if prev < 0:
            check(prev, "b200mp_set_friction_mode")
        return "closed_form" if prev == 1 else "auto"

    def set_collision_mode(self, mode: str) -> str:
        """``"auto"`` (bounding-box broad phase over 32-point obstacle chunks, FP32 screen, exact FP64 recheck of
        undecided pairs; default), ``"screen"`` (the same without the broad phase) or ``"fp64"`` (all-FP64 kernel).
        All give bit-identical flags; process-wide.  Returns the previous mode."""
        names = {"auto": 0, "fp64": 1, "screen": 2}
        if mode not in names:
            raise ValueError(f"collision mode must be one of {sorted(names)}")
        prev = self.lib.b200mp_set_collision_mode(names[mode])
        if prev < 0:
            check(prev, "b200mp_set_collision_mode")
        return {0: "auto", 1: "fp64", 2: "screen"}[prev]

    def select_best_path_index_batch(self, end_x, end_y, free, goal_xy, weight: float, norm_mode: Optional[int] = None,
                                     want_scores: bool = False):
        """``select_best_path_index`` on end points; returns ``int`` or ``None`` (synchronises).  ``free``: bool, or
        uint8 with 1 = free, 0 = colliding, 2 = excluded (dropped by the planner: no candidate, no penalty)."""
        ex, ey = self.dev(end_x), self.dev(end_y)
        fr = self.dev(free, torch.uint8) if not (isinstance(free, torch.Tensor) and free.dtype == torch.bool) \
            else free.to(self.tdev).to(torch.uint8)
        P = ex.numel()
        if norm_mode is None:
            if self._norm2_mode is None:
                self._norm2_mode = host_norm2_mode()
            norm_mode = self._norm2_mode
        best = self.empty(1, dtype=torch.int32)
        scores = self.empty(P) if want_scores else None
        check(self.lib.b200mp_select_best_f64(self.device, self._stream(), P, self._ptr(ex), self._ptr(ey), self._ptr(fr),
                                              float(goal_xy[0]), float(goal_xy[1]), float(weight), int(norm_mode),
                                              self._ptr(scores), self._ptr(best)), "b200mp_select_best_f64")
        b = int(best.item())
        b = None if b < 0 else b
        return (b, scores) if want_scores else b

    # -------------------------------------------------------------- measurement
    def fma_peak(self, dtype_bits: int = 64, reps: int = 5) -> float:
        """Measured FMA-pipe peak in TFLOP/s (register-resident chains), the rollout roofline denominator."""
        out = C.c_double(0.0)
        check(self.lib.b200mp_fma_peak(self.device, dtype_bits, reps, C.byref(out)), "b200mp_fma_peak")
        return out.value
