#!/usr/bin/env python
"""bench.py -- 7-DoF RK4 rollout-steps/s (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch: BASELINE.json config 2 on every GPU -- 65,536
open-loop rollouts x 500 RK4 steps, FP64, the full trajectory (2.62 GB) written to HBM.  Rollouts are
independent, so N GPUs run N such batches with no data-path collective (weak scaling).

One JSON line is printed by rank 0.  ``value`` is timed with CUDA events around each step with the inputs
resident in HBM; ``e2e`` goes through the public host-buffer API (H2D of the inputs, time-chunked kernels,
D2H of the whole trajectory, overlapped) and is the number to hold against the CPU reference arm.
``--impl reference`` times the reference's own CPU path on the box's host cores on a bounded sample of the same
workload: the UNMODIFIED Python (``planar_model_RK4`` under ``multiprocessing.Pool(os.cpu_count())``) from the verbatim
copy staged under ``baseline/_ref`` (``baseline/stage_reference.py``), with the plain-C/OpenMP port (``oracle/csrc``)
timed beside it; the port alone when no copy is staged.  The lattice half of the metric (collision checks/s, config 3)
is reported in the same line: ``collision``, ``roofline_collision``, ``cpu_baseline_collision``.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "7-DoF RK4 rollout-steps/sec"
UNIT = "rollout-steps/s"
B, N_STEPS, HOLD, DT = 65536, 500, 10, 1e-4
ALG_FLOP_PER_STEP = 854          # 806 add/mul/div/sqrt flops + 48 transcendental values
ALG_BYTES_PER_STEP = 80          # 10 states x 8 B written per rollout-step
ALG_FLOP_PER_TEST = 8            # one circle-vs-point test
WORKLOAD = f"config2: {B} rollouts x {N_STEPS} steps, FP64, dt=1e-4, ZOH-{HOLD} controls, full trajectory stored"


def bench_config(n_gpus: int) -> dict:
    """The `config` object of BOTH arms (the reference arm times a bounded sample of this workload; what the sample
    was is stated in its `cpu_baseline.sample`)."""
    return {"workload": WORKLOAD, "per_gpu_batch": B, "n_steps": N_STEPS,
            "parallelism": f"dp{n_gpus} (independent rollouts, no data-path collective)",
            "l2": "256 MiB buffer written between timed iterations (untimed); each step also streams 2.62 GB of output through the 126 MB L2"}


K1_SOURCES = ("b200mp_math.cuh", "rollout_kernels.cuh", "rollout_kernels_f64.cu", "slice_sched.cuh", "vehicle_rhs.cuh")


def kernel_source_sha16() -> str:
    """Hash of the CUDA sources the ncu-derived numbers in profiles/roofline_inputs.json were captured from."""
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "python_motionplanning_b200", "csrc")
    for name in K1_SOURCES:         # everything the headline kernel is compiled from
        with open(os.path.join(csrc, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def _host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------- CPU reference
def cpu_port_run(threads: int, rollouts_per_thread: int = 2048, repeats: int = 1):
    """Time the plain-C port of the reference's RK4 path on `threads` host threads (bounded sample)."""
    import numpy as np

    from oracle import c_oracle, planar_numpy
    from python_motionplanning_b200 import workloads as wl
    nb = min(B, threads * rollouts_per_thread)
    s0, d, t = wl.config2_rollouts(B=B, n_steps=N_STEPS)
    s0, d, t = (np.ascontiguousarray(a[..., :nb]) for a in (s0, d, t))
    par = c_oracle.make_params(planar_numpy.VehicleParams())
    par[0].D[:] = (1.0,) * 4
    c_oracle.rollout(s0[:, :64], d[:, :, :64], t[:, :, :64], par, DT, 50, hold=HOLD, nthreads=threads)   # warm
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        c_oracle.rollout(s0, d, t, par, DT, N_STEPS, hold=HOLD, store_stride=1, nthreads=threads)
        times.append(time.perf_counter() - t0)
    sample = f"first {nb} rollouts of config 2 x {N_STEPS} steps, full trajectory, {threads} OpenMP threads"
    return nb * N_STEPS, times, sample


def literal_run(rollouts_per_proc: int = 16, paths_per_proc: int = 0, repeats: int = 1, timeout: int = 900):
    """The UNMODIFIED reference timed in a subprocess (no CUDA context is forked): oracle/literal_baseline.py.
    Returns its JSON object, or {"unavailable": why}."""
    cmd = [sys.executable, "-m", "oracle.literal_baseline", "--rollouts-per-proc", str(rollouts_per_proc),
           "--paths-per-proc", str(paths_per_proc), "--repeats", str(repeats)]
    try:
        res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
        for ln in reversed(res.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"unavailable": f"literal baseline printed no JSON (rc={res.returncode}): {res.stderr[-200:]}"}
    except Exception as exc:  # noqa: BLE001
        return {"unavailable": f"{type(exc).__name__}: {exc}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = _host_threads()
    # the C port, all host threads (a far stronger CPU implementation than the reference's own Python)
    cpu_port_run(threads, rollouts_per_thread=256)
    p_units, p_times, p_sample = cpu_port_run(threads, repeats=max(1, min(args.steps, 3)))
    port = {"value": p_units * len(p_times) / sum(p_times), "unit": UNIT, "cores": threads, "kind": "port", "sample": p_sample,
            "note": "plain-C/OpenMP restatement of vehicle_model.py:220-445 (oracle/csrc/oracle.c)"}
    # the unmodified reference: every step = procs x R rollouts x 500 steps, R sized so that warmup + steps end in ~2 min
    total_steps = args.steps + args.warmup
    R = max(1, min(16, 360 // max(total_steps, 1)))
    lit = literal_run(rollouts_per_proc=R, repeats=total_steps)
    times, units, procs, sample = [], 0, None, None
    if "rollout" in lit:
        times = lit["rollout"]["seconds_each"][args.warmup:]
        units = lit["rollout"]["rollouts"] * lit["rollout"]["n_steps"] * len(times)
        procs, sample = lit["procs"], lit["rollout"]["sample"]
    if times:
        value, ms = units / sum(times), 1e3 * sum(times) / len(times)
        cpu = {"value": value, "unit": UNIT, "cores": procs, "kind": "reference", "sample": sample + f"; one such sample per step ({len(times)} timed)",
               "reference_root": os.path.relpath(lit.get("reference_root", ""), ROOT) if lit.get("reference_root", "").startswith(ROOT) else lit.get("reference_root"),
               "port": port}
    else:
        value, ms = port["value"], 1e3 * sum(p_times) / len(p_times)
        cpu = dict(port, literal_unavailable=lit.get("unavailable"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args.gpus),
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock, power and throttle reasons of one GPU while the timed region runs: NVML in a
    thread (2 ms period, so even a 60 ms region gets ~30 samples), nvidia-smi polling as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.tmp = None
        self.rows = []
        self.thread = None
        self.stop_flag = threading.Event()
        self.how = None

    def _nvml_loop(self, nv, h, max_mhz):
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mw = nv.nvmlDeviceGetPowerUsage(h)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((float(mhz), float(max_mhz), mw / 1000.0, [k for k, b in bits.items() if r & b]))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical order; map through the UUID of the CUDA device when visible-devices remaps
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            h = None
            for i in range(nv.nvmlDeviceGetCount()):
                cand = nv.nvmlDeviceGetHandleByIndex(i)
                u = nv.nvmlDeviceGetUUID(cand)
                u = u.decode() if isinstance(u, bytes) else u
                if uuid in u:
                    h = cand
                    break
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h, max_mhz), daemon=True)
            self.thread.start()
            self.how = "nvml thread, 2 ms period"
            return
        except Exception:
            self.thread = None
        try:
            self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
            self.how = "nvidia-smi -lms 20"
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "how": self.how}
        rows = []
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            rows = [(a, b, c, d) for a, b, c, d in self.rows]
            names = None
        elif self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            self.tmp.flush()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            with open(self.tmp.name) as f:
                for ln in f:
                    parts = [p.strip() for p in ln.split(",")]
                    if len(parts) >= 7:
                        try:
                            rows.append((float(parts[0]), float(parts[1]), float(parts[2]),
                                         [names[i] for i in range(4) if parts[3 + i].lower().startswith("active")]))
                        except ValueError:
                            pass
            os.unlink(self.tmp.name)
        if not rows:
            return out
        loaded = [r for r in rows if r[2] > 0.5 * max(x[2] for x in rows)] or rows
        reasons = sorted({k for r in rows for k in r[3]})
        out.update(sm_mhz=statistics.median(r[0] for r in loaded), sm_max_mhz=max(r[1] for r in rows), reasons=reasons,
                   samples=len(rows), power_w_max=max(r[2] for r in rows))
        return out


# ---------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import python_motionplanning_b200 as mp
    from python_motionplanning_b200 import workloads as wl

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = mp.Engine(local_rank)
    dev = eng.tdev
    p = mp.VehicleParameters()
    p.DFL = p.DFR = p.DRL = p.DRR = 1.0          # mu_max = [1, 1, 1, 1] (drive.py:142) folded into the set
    eng.set_params(p)

    # each rank gets its own seeded batch (weak scaling; same distribution)
    s0_h, d_h, t_h = wl.config2_rollouts(B=B, n_steps=N_STEPS, seed=wl.SEED + rank)
    s0, dl, tq = eng.dev(s0_h), eng.dev(d_h), eng.dev(t_h)
    traj = eng.empty(N_STEPS, 10, B)
    end = eng.empty(12, B)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2

    def step():
        eng.rollout(s0, dl, tq, DT, N_STEPS, hold=HOLD, store_stride=1, traj_out=traj, state_out=end)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    def timed_region():
        sampler = ClockSampler(local_rank)
        sampler.start()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        t0 = time.perf_counter()
        for e0, e1 in evs:
            flush.fill_(1)                      # evict L2 between timed iterations (untimed)
            e0.record()
            step()
            e1.record()
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop()
        ms = [e0.elapsed_time(e1) for e0, e1 in evs]
        return ms, wall, clocks

    ms, wall, clocks = timed_region()
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    remeasured = False
    if bad & set(clocks["reasons"]):
        remeasured = True
        ms, wall, clocks = timed_region()
    dev_ms = sum(ms)
    t_all = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    max_ms = float(t_all.item())
    units_per_step = B * N_STEPS
    value = world * units_per_step * args.steps / (max_ms * 1e-3)

    # ---- end to end through the public host-buffer API (pinned host <-> device copies inside the timed region)
    hs, hd, ht = (torch.from_numpy(a).pin_memory() for a in (s0_h, d_h, t_h))
    traj_host = torch.empty(N_STEPS, 10, B, dtype=torch.float64).pin_memory()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(2):
        eng.rollout_to_host(hs, hd, ht, DT, N_STEPS, HOLD, traj_host, chunk_steps=50)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.rollout_to_host(hs, hd, ht, DT, N_STEPS, HOLD, traj_host, chunk_steps=50)
    barrier()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * units_per_step * e2e_steps / float(t_e.item())
    h2d = int(hs.numel() + hd.numel() + ht.numel()) * 8
    d2h = int(traj_host.numel()) * 8
    # what the box allows for that read-back: every rank copies one 50-step slab (262 MB) device -> pinned host, all ranks at
    # once, nothing else running.  The step time is a max over ranks, so the bound on the aggregate is N x the SLOWEST rank's rate.
    cp = torch.cuda.Stream(dev)
    barrier()
    with torch.cuda.stream(cp):
        for _ in range(2):
            traj_host[:50].copy_(traj[:50], non_blocking=True)
        cp.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(10):
            traj_host[:50].copy_(traj[:50], non_blocking=True)
        cp.synchronize()
    my_rate = torch.tensor([10 * traj[:50].numel() * 8 / (time.perf_counter() - t0) / 1e9], dtype=torch.float64, device=dev)
    rates = [my_rate.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(rates, my_rate)
    rates = [float(r.item()) for r in rates]
    e2e_gbs = e2e_value * ALG_BYTES_PER_STEP / 1e9
    d2h_ceiling = {"gbs_per_gpu_all_ranks_at_once": [round(r, 2) for r in rates], "gbs_aggregate": round(sum(rates), 2),
                   "gbs_bound_by_slowest_rank": round(world * min(rates), 2), "e2e_d2h_gbs": round(e2e_gbs, 2),
                   "frac_of_d2h_ceiling": e2e_gbs / (world * min(rates)),
                   "note": "concurrent cudaMemcpyAsync device -> pinned host of one 262 MB slab per rank, measured in this run; a "
                           "single B200 reads back at ~56 GB/s alone (PCIe Gen5 x16), eight at once share the host's write path "
                           "(profiles/r02_d2h_ceiling.md)"}
    # supplementary: the same host-buffer call when only the final states are read back (no trajectory D2H); the
    # controls are uploaded chunk by chunk behind the kernels
    end_host = torch.empty(12, B, dtype=torch.float64).pin_memory()
    for _ in range(2):
        eng.rollout_endstate_to_host(hs, hd, ht, DT, N_STEPS, HOLD, end_host, chunk_steps=100)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.rollout_endstate_to_host(hs, hd, ht, DT, N_STEPS, HOLD, end_host, chunk_steps=100)
    barrier()
    t_e2 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2, op=dist.ReduceOp.MAX)
    e2e_end_value = world * units_per_step * e2e_steps / float(t_e2.item())

    # ---- config 4 as BASELINE.json states it: 1,048,576 control sequences x 100 steps SHARDED over the ranks (strong
    # scaling), each plan = sample -> rollout with cost -> local argmin -> one all-gather of per-rank
    # winner records (cost, index, control sequence; NCCL); device-timed, max over ranks
    mpc_sharded = None
    if world > 1:
        from python_motionplanning_b200 import distributed as D
        cfg4 = wl.config4_mpc(B=1 << 20)
        for _ in range(3):
            plan = D.mpc_plan(eng, cfg4)
        barrier()
        n_plans = 10
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        for _ in range(n_plans):
            plan = D.mpc_plan(eng, cfg4)
        m1.record()
        barrier()
        t_m = torch.tensor([m0.elapsed_time(m1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t_m, op=dist.ReduceOp.MAX)
        ms_plan = float(t_m.item()) / n_plans
        # the kernels of one rank's shard alone (sample + rollout with cost + winner record; no collective, no host read)
        pl = eng._mpc_planners[next(iter(eng._mpc_planners))]
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(n_plans):
            eng.mpc_sample_controls_into(pl.delta, pl.torque, pl.seed, rollout0=pl.lo, delta_mean=cfg4["delta_mean"],
                                         delta_sigma=cfg4["delta_sigma"], delta_clip=cfg4["delta_clip"],
                                         torque_mean=cfg4["torque_mean"], torque_sigma=cfg4["torque_sigma"])
            eng.rollout(pl.state0, pl.delta, pl.torque, pl.dt, pl.n_steps, hold=1, cost_ref=pl.cost_ref, w_u=cfg4["w_u"],
                        u_ref=cfg4["u_ref"], state_broadcast=True, cost_out=pl.cost, state_out=pl.state_end)
            eng.mpc_winner(pl.cost, pl.delta, pl.torque, index_offset=pl.lo, record_out=pl.rec)
        k1.record()
        barrier()
        t_k = torch.tensor([k0.elapsed_time(k1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t_k, op=dist.ReduceOp.MAX)
        mpc_sharded = {"value": cfg4["B"] * cfg4["n_steps"] / (ms_plan * 1e-3), "unit": UNIT, "ms_per_plan": ms_plan,
                       "kernels_only_ms_per_plan": float(t_k.item()) / n_plans,
                       "sequences_total": cfg4["B"], "horizon": cfg4["n_steps"], "scaling": "strong", "n_gpus": world,
                       "best_index": int(plan["index"]), "best_cost": float(plan["cost"]), "owner_rank": int(plan["owner"]),
                       "includes": "resident MpcPlanner: sampling + rollout with cost (broadcast start state) + winner-record kernels, ONE "
                                   "NCCL all_gather_into_tensor of per-rank winner records (cost, index, control sequence) and one "
                                   "pinned device->host read of the records, every plan synchronised on the host"}

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel (rk4_rollout_kernel<double,...>), measured live on this GPU
        peak64 = eng.fma_peak(64, reps=5)
        kernel_ms = statistics.mean(ms)
        steps_per_s_gpu = units_per_step / (kernel_ms * 1e-3)
        achieved_tf = steps_per_s_gpu * ALG_FLOP_PER_STEP * 1e-12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = None
        prof = {}
        src_sha = kernel_source_sha16()
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "roofline_inputs.json")))
        except Exception:
            pass
        # ncu-derived numbers are carried only while the CUDA sources are the ones that were profiled
        prof_fresh = prof.get("csrc_sha16") == src_sha
        if not prof_fresh:
            prof = {"stale": True}
        traffic = prof.get("rk4_rollout_f64_dram_bytes_per_launch")
        roofline = {
            "bound": "fp64", "kernel": "rk4_rollout_kernel<double,front-steer>", "achieved": achieved_tf, "peak": peak64,
            "unit": "TFLOP/s", "frac": achieved_tf / peak64 if peak64 else None, "traffic": traffic,
            "peak_source": "measured live: register-resident DFMA chains (b200mp_fma_peak), FMA = 2 flop; "
                           "MEASURED_PEAKS.json has no FP64 entry",
            "algorithmic_flop_per_step": ALG_FLOP_PER_STEP, "kernel_ms": kernel_ms,
            "fp64_pipe_util_ncu": prof.get("rk4_rollout_f64_fp64_pipe_pct"),
            "smem_wavefront_util_ncu": prof.get("rk4_rollout_f64_smem_wavefront_pct"),
            "issue_active_ncu": prof.get("rk4_rollout_f64_issue_active_pct"),
            "ncu_source": prof.get("rk4_rollout_f64_source_file"), "csrc_sha16": src_sha,
            "ncu_numbers_match_this_source": prof_fresh,
            "note": "friction from a host-built polynomial table in shared memory (no sqrt/atan/sin in the kernel); traffic and the "
                    "*_ncu fields come from profiles/roofline_inputs.json and are dropped (null) when its csrc_sha16 differs from "
                    "the sources this run was built from",
            "hbm": {"achieved": steps_per_s_gpu * ALG_BYTES_PER_STEP * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": steps_per_s_gpu * ALG_BYTES_PER_STEP * 1e-9 / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"},
        }
        cpu_baseline = None
        col = {}
        if world == 1:
            threads = _host_threads()
            units, times, sample = cpu_port_run(threads)
            cpu_baseline = {"value": units / times[0], "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
            lit = literal_run(rollouts_per_proc=16, paths_per_proc=2)
            if "rollout" in lit:
                cpu_baseline["literal"] = {"value": lit["rollout"]["value"], "unit": UNIT, "cores": lit["procs"], "kind": "reference",
                                           "sample": lit["rollout"]["sample"], "seconds": lit["rollout"]["seconds"]}
            else:
                cpu_baseline["literal"] = {"unavailable": lit.get("unavailable", "no rollout leg")}
            col = collision_metrics(eng, wl, np, torch, lit)
        extras = secondary_metrics(eng, wl, np, torch) if world == 1 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": bench_config(world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "Engine.rollout_to_host: pinned host inputs -> H2D, 10 time-chunks of 50 steps, D2H of the full trajectory overlapped on a copy stream",
                    "bound": "device -> host read-back of 2.62 GB of trajectory per step and GPU", "d2h_ceiling": d2h_ceiling,
                    "frac_of_d2h_ceiling": d2h_ceiling["frac_of_d2h_ceiling"],
                    "endstate_only": {"value": e2e_end_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12 * B * 8,
                                      "api": "Engine.rollout_endstate_to_host: same host inputs, controls uploaded in 5 time-chunks behind "
                                             "the kernels, only the [12, B] end states read back"}},
            "gpu_launches": args.steps * world,
            "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                       "samples": clocks["samples"], "how": clocks.get("how"), "power_w_max": clocks.get("power_w_max"),
                       "remeasured": remeasured},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "wall_s_timed_region": wall,
        }
        line.update(col)          # collision, roofline_collision, cpu_baseline_collision (N = 1)
        if extras:
            line["secondary"] = extras
        if mpc_sharded:
            line["secondary"] = {"mpc_plan_sharded": mpc_sharded}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


def collision_metrics(eng, wl, np, torch, lit):
    """The metric's second half -- lattice collision checks/s on BASELINE config 3 (4,096 paths x 49 points x 3 circles vs
    10,000 obstacle points) -- as top-level objects: `collision` (kernel-resident value, the public-API number with host
    buffers, the variants), `roofline_collision` (the SHIPPED broad-phase kernel on executed and on nominal tests) and
    `cpu_baseline_collision` (C port on all host threads + the unmodified collision_check on a subsample)."""
    w = wl.config3_lattice()
    P, n = w["px"].shape
    M = len(w["obstacles"])
    nc = len(w["offsets"])
    tests = P * n * nc * M
    unit = "circle-point tests/s (nominal P*49*3*M)"
    px, py, yaw = eng.dev(w["px"]), eng.dev(w["py"]), eng.dev(w["pyaw"])
    obs = eng.dev(w["obstacles"])
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def dev_ms(fn, reps=10, warm=3):
        evs = [(ev(), ev()) for _ in range(reps + warm)]
        for e0, e1 in evs:
            e0.record()
            r = fn()
            e1.record()
        torch.cuda.synchronize()
        return statistics.median(e0.elapsed_time(e1) for e0, e1 in evs[warm:]), r

    def wall_ms(fn, reps=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            r = fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return statistics.median(ts) * 1e3, r

    out = {}
    # ---- shipped path, inputs resident: yaws on the device, proven verdicts, host-resolved leftovers (count read back)
    ms_res, free = wall_ms(lambda: eng.collision_check_batch(px, py, yaw, obs, w["offsets"], w["radii"]))
    undecided = eng.last_collision_undecided
    st2 = (C.c_ulonglong * 2)()
    eng.lib.b200mp_collision_stats(eng.device, eng._stream(), M, st2)
    n_warps = -(-P * n // 32)
    executed = int(st2[0]) * 32 * 32 * nc
    # ---- public API with HOST buffers (numpy in, numpy out): H2D of px, py, pyaw, obstacles; D2H of the flags
    hx, hy, hyaw, hobs = w["px"], w["py"], w["pyaw"], w["obstacles"]
    ms_e2e, _ = wall_ms(lambda: eng.collision_check_batch(hx, hy, hyaw, hobs, w["offsets"], w["radii"]).cpu())
    h2d = int(hx.nbytes + hy.nbytes + hyaw.nbytes + hobs.nbytes)
    # ---- the former default (numpy cos / sin of all 200k yaws on the host) for comparison
    ms_host_trig, _ = wall_ms(lambda: eng.collision_check_batch(px, py, w["pyaw"], obs, w["offsets"], w["radii"], host_trig=True), reps=5)
    # ---- kernels only, CUDA events, caller-supplied trig (no host work in the call), all three arithmetic modes; the
    # shipped scene and the same scene with the obstacles out of reach (nothing collides: no early exit, nothing culled by verdict)
    trig = eng.path_trig(w["pyaw"], n)
    far = eng.dev(w["obstacles"] + np.array([400.0, 0.0]))
    peak32 = eng.fma_peak(32, reps=3)
    peak64 = eng.fma_peak(64, reps=3)
    variants = {}
    for name, ob in (("shipped_scene", obs), ("no_early_exit", far)):
        for mode in ("auto", "screen", "fp64"):
            ms, fr = dev_ms(lambda: eng.collision_check_batch(px, py, None, ob, w["offsets"], w["radii"], trig=trig, mode=mode))
            variants[f"{name}_{mode}"] = {"ms": ms, "value": tests / (ms * 1e-3), "free_fraction": float(fr.float().mean().item())}
    ms_k = variants["shipped_scene_auto"]["ms"]
    ms_clear, _ = dev_ms(lambda: eng.collision_check_batch(px, py, None, obs, w["offsets"], w["radii"], trig=trig, want_clearance=True))
    ms_clear_api, _ = wall_ms(lambda: eng.collision_check_batch(px, py, w["pyaw"], obs, w["offsets"], w["radii"], want_clearance=True,
                                                                 clearance_trig="host"), reps=5)
    ms_clear_api_auto, _ = wall_ms(lambda: eng.collision_check_batch(px, py, w["pyaw"], obs, w["offsets"], w["radii"], want_clearance=True), reps=5)
    ms_clear_api_dev, _ = wall_ms(lambda: eng.collision_check_batch(px, py, w["pyaw"], obs, w["offsets"], w["radii"], want_clearance=True,
                                                                     clearance_trig="device"), reps=5)
    out["collision"] = {
        "metric": "lattice collision checks/sec", "unit": unit, "workload": f"config3: {P} paths x {n} points x {nc} circles vs {M} obstacle points, bit-exact flags",
        "value": tests / (ms_k * 1e-3), "ms": ms_k, "paths_per_s": P / (ms_k * 1e-3),
        "value_note": "memset + obstacle_prepare_kernel + collision_cull_kernel<3>, CUDA events, paths / trig / obstacles resident",
        "free_fraction": float(free.float().mean().item()),
        "e2e": {"value": tests / (ms_e2e * 1e-3), "unit": unit, "ms": ms_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": P + 4,
                "api": "Engine.collision_check_batch(numpy px, py, pyaw, obstacles) -> flags on the host: yaws evaluated on the device with "
                       "proven verdicts, undecided path points resolved with host numpy cos/sin (b200mp_collision_check_yaw_f64 + _resolve_)",
                "path_points_resolved_on_host": undecided},
        "api_inputs_resident": {"ms": ms_res, "value": tests / (ms_res * 1e-3),
                                "includes": "kernels + the 4-byte undecided count read back + stream sync"},
        "api_host_trig_all_yaws": {"ms": ms_host_trig, "note": "the former default: numpy cos/sin of 200k yaws on the host + their H2D"},
        "min_clearance": {"kernel_ms": ms_clear, "value": tests / (ms_clear * 1e-3), "api_ms_host_trig": ms_clear_api,
                          "api_ms_device_trig": ms_clear_api_dev, "api_ms": ms_clear_api_auto,
                          "api": "default (clearance_trig='auto'): device-trig clearance of every path point, numpy cos/sin and an exact "
                                 "re-evaluation only for the points that can hold their path's minimum; bit-identical to the oracle",
                          "candidates_resolved_on_host": getattr(eng, "last_clearance_candidates", None)},
        "variants": variants,
    }
    # ---- roofline of the shipped kernel: 4 FP32 lane-operations per executed test (2 subtractions, 1 multiply, 1 FMA = 5 flop)
    k_far = variants["no_early_exit_screen"]
    out["roofline_collision"] = {
        "kernel": "collision_cull_kernel<3> (+ obstacle_prepare_kernel)", "bound": "fp32",
        "achieved": executed * 5 / (ms_k * 1e-3) * 1e-12, "peak": peak32, "unit": "TFLOP/s",
        "frac": executed * 5 / (ms_k * 1e-3) * 1e-12 / peak32,
        "basis": "EXECUTED tests (b200mp_collision_stats: warp-chunks that survive the bounding-box broad phase x 32 path points x 32 "
                 "obstacle points x 3 circles) x 5 FP32 flop, over the whole launch",
        "tests_executed": executed, "tests_nominal": tests, "executed_fraction": executed / tests,
        "warp_chunks_screened": int(st2[0]), "warp_chunks_total": n_warps * (-(-M // 32)), "thread_chunks_rechecked_fp64": int(st2[1]),
        "nominal": {"achieved": tests * ALG_FLOP_PER_TEST / (ms_k * 1e-3) * 1e-12, "peak": peak64, "unit": "TFLOP/s",
                    "frac": tests * ALG_FLOP_PER_TEST / (ms_k * 1e-3) * 1e-12 / peak64,
                    "basis": "NOMINAL tests x 8 FP64 flop (SURVEY.md 8d) against the measured FP64 peak: above 1 because culled and "
                             "FP32-screened pairs never reach the FP64 pipe (flags proven bit-exact)"},
        "screen_every_pair": {"ms": k_far["ms"], "achieved": k_far["value"] * 5 * 1e-12, "peak": peak32, "frac": k_far["value"] * 5 * 1e-12 / peak32,
                              "frac_of_fp32_lane_issue": k_far["value"] * 4 / (peak32 * 1e12 / 2),
                              "basis": "collision_filter_kernel<3>, obstacles out of reach: every nominal test executed in packed FP32"},
        "peak_source": "measured live: register-resident FFMA / DFMA chains (b200mp_fma_peak)", "traffic": None,
        "traffic_note": "inputs 4.9 MB, outputs 4 KB per launch: not memory-bound (profiles/r02_collision_cull.md)",
    }
    # ---- CPU: the C port on all host threads (full config 3, early exit like the reference) + the literal reference
    from oracle import c_oracle
    threads = _host_threads()
    c_oracle.collision_check(w["px"][:64], w["py"][:64], w["pyaw"][:64], w["obstacles"], w["offsets"], w["radii"], nthreads=threads)
    t0 = time.perf_counter()
    ref_free, _, c_tests = c_oracle.collision_check(w["px"], w["py"], w["pyaw"], w["obstacles"], w["offsets"], w["radii"], nthreads=threads)
    c_s = time.perf_counter() - t0
    same = bool(np.array_equal(ref_free, free.cpu().numpy().astype(bool)))
    cb = {"value": tests / c_s, "unit": unit, "cores": threads, "kind": "port", "seconds": c_s, "paths_per_s": P / c_s,
          "tests_executed": int(c_tests), "flags_equal_gpu": same,
          "sample": f"all {P} config-3 paths, plain-C restatement of collision_checker.py:32-117 with the reference's early exit, {threads} OpenMP threads"}
    if "collision" in lit:
        lc = lit["collision"]
        cb["literal"] = {"value": lc["value"], "unit": unit, "cores": lit["procs"], "kind": "reference", "paths_per_s": lc["paths_per_s"],
                         "seconds": lc["seconds"], "sample": lc["sample"],
                         "flags_equal_gpu": bool(np.array_equal(np.array(lc["free"], bool), free.cpu().numpy().astype(bool)[:lc["paths"]]))}
    else:
        cb["literal"] = {"unavailable": lit.get("unavailable", "no collision leg")}
    out["cpu_baseline_collision"] = cb
    return out


def secondary_metrics(eng, wl, np, torch):
    """The other configs and kernels, N = 1 only (the collision half of the metric is `collision_metrics`)."""
    out = {}
    w = wl.config3_lattice()
    P, n = w["px"].shape
    M = len(w["obstacles"])
    tests = P * n * 3 * M
    px, py = eng.dev(w["px"]), eng.dev(w["py"])
    obs = eng.dev(w["obstacles"])
    free = eng.collision_check_batch(px, py, w["pyaw"], obs, w["offsets"], w["radii"])
    peak32 = eng.fma_peak(32, reps=3)
    ex_d, ey_d = px[:, -1].contiguous(), py[:, -1].contiguous()
    eng.select_best_path_index_batch(ex_d, ey_d, free, w["goal"], w["weight"])   # first call probes the host's norm closed form
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    best = eng.select_best_path_index_batch(ex_d, ey_d, free, w["goal"], w["weight"])
    out["select_best"] = {"P": P, "ms": (time.perf_counter() - t0) * 1e3, "best_index": best,
                          "includes": "score + argmin kernels and the device->host read of the index (the call returns a host int)"}
    # FP32 twin of the headline kernel on the same batch
    s0_h, d_h, t_h = wl.config2_rollouts(B=B, n_steps=N_STEPS)
    s32, d32, t32 = (eng.dev(a, torch.float32) for a in (s0_h, d_h, t_h))
    tr32 = eng.empty(N_STEPS, 10, B, dtype=torch.float32)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(4):
        if k == 3:
            ev0.record()
        eng.rollout(s32, d32, t32, DT, N_STEPS, hold=HOLD, store_stride=1, dtype="f32", traj_out=tr32)
    ev1.record()
    torch.cuda.synchronize()
    ms32 = ev0.elapsed_time(ev1)
    out["rollout_f32"] = {"value": B * N_STEPS / (ms32 * 1e-3), "unit": UNIT, "ms": ms32, "fp32_peak_tflops": peak32,
                          "frac_of_fp32_peak": B * N_STEPS / (ms32 * 1e-3) * ALG_FLOP_PER_STEP * 1e-12 / peak32}
    del tr32
    # device-resident lattice pipeline: 4,096 goal states -> optimised spirals -> sampled and
    # transformed paths -> collision flags -> best index; nothing but the goal states and obstacles is uploaded
    par = w["spiral_params"]
    k1d, k2d, sfd, egod = eng.dev(par[0]), eng.dev(par[1]), eng.dev(par[2]), eng.dev(w["ego"])
    lat0 = eng.sample_lattice(k1d, k2d, sfd, ego=None, want_trig=False)          # ego-frame end states as goal states
    tf_goal = eng.dev(3.0 * (par[0] + par[1]) * par[2] / 8.0)
    def pipeline(optimise):
        if optimise:
            o = eng.optimize_spirals(lat0["end_xy"][0], lat0["end_xy"][1], tf_goal)
            a1, a2, a3 = o["p"][0], o["p"][1], o["p"][2]
        else:
            a1, a2, a3 = k1d, k2d, sfd
        lt = eng.sample_lattice(a1, a2, a3, ego=egod, want_trig=False)
        fr = eng.collision_check_batch(lt["px"], lt["py"], lt["pyaw"], obs, w["offsets"], w["radii"])
        return eng.select_best_path_index_batch(lt["end_xy"][0], lt["end_xy"][1], fr, w["goal"], w["weight"]), fr
    for name, optimise in (("lattice_pipeline_sample_check_select", False), ("lattice_pipeline_optimise_sample_check_select", True)):
        for _ in range(3):
            pipeline(optimise)
        ts = []
        for _ in range(10):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            bi, fr = pipeline(optimise)
            ts.append(time.perf_counter() - t0)             # select_best returns a host int: the call synchronises
        sec = statistics.median(ts)
        out[name] = {"ms": sec * 1e3, "paths_per_s": P / sec, "value": tests / sec, "unit": "circle-point tests/s (nominal)",
                     "best_index": bi, "free_fraction": float(fr.float().mean().item()),
                     "includes": "all kernels + the undecided count and the chosen index read back; no host trig on the bulk, no path upload"}
    # closed-loop tracking: 65,536 vehicles on 16 waypoint lists of 3,000 points, 500 sub-steps =
    # 50 Stanley/PID updates each; same RK4 work per step as the headline metric plus the controllers
    st0, wps = wl.tracking_fleet(V=B, n_sets=16)
    st0_d, wps_d = eng.dev(st0), eng.dev(wps)
    for name, kw in (("closed_loop_tracking", {}), ("closed_loop_tracking_datalog_stride10", {"store_stride": 10, "want_log": True})):
        for k in range(3):
            if k == 2:
                ev0.record()
            eng.track_closed_loop(st0_d, wps_d, DT, N_STEPS, 25.0, vehicles_per_set=B // 16, **kw)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        out[name] = {"value": B * N_STEPS / (ms * 1e-3), "unit": "closed-loop rollout-steps/s", "ms": ms,
                     "vehicles": B, "waypoints_per_list": int(wps.shape[1]), "control_updates": N_STEPS // 10}
    # the same launch with the closed-form friction (sqrt / atan / sin in the kernel)
    eng.set_friction_mode("closed_form")
    trj = eng.empty(N_STEPS, 10, B)
    s0c, dlc, tqc = eng.dev(s0_h), eng.dev(d_h), eng.dev(t_h)
    for k in range(4):
        if k == 3:
            ev0.record()
        eng.rollout(s0c, dlc, tqc, DT, N_STEPS, hold=HOLD, store_stride=1, traj_out=trj)
    ev1.record()
    torch.cuda.synchronize()
    eng.set_friction_mode("auto")
    out["rollout_f64_closed_form"] = {"value": B * N_STEPS / (ev0.elapsed_time(ev1) * 1e-3), "unit": UNIT, "ms": ev0.elapsed_time(ev1)}
    del trj
    # end-state-only FP64 (no trajectory writeback): separates compute from writeback
    s0, dl, tq = eng.dev(s0_h), eng.dev(d_h), eng.dev(t_h)
    for k in range(4):
        if k == 3:
            ev0.record()
        eng.rollout(s0, dl, tq, DT, N_STEPS, hold=HOLD, store_stride=0)
    ev1.record()
    torch.cuda.synchronize()
    out["rollout_f64_endstate_only"] = {"value": B * N_STEPS / (ev0.elapsed_time(ev1) * 1e-3), "unit": UNIT,
                                        "ms": ev0.elapsed_time(ev1)}
    del s0, dl, tq
    # config 1's inner call: the scalar drop-in VehicleModel.planar_model_RK4 (one vehicle, one step per call, host lists
    # in and out, as drive.py:141-143 calls it) -- launch- and copy-latency bound by construction
    import python_motionplanning_b200 as mp
    vm = mp.VehicleModel(dt=DT, engine=eng)
    pv = mp.VehicleParameters()
    st1 = [20.0, 0.0, 0.0, 20.0 / pv.rw, 20.0 / pv.rw, 20.0 / pv.rw, 20.0 / pv.rw, 0.0, 0.0, 0.0]
    axp = ayp = 0.0
    for k in range(230):
        if k == 30:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        r1 = vm.planar_model_RK4(st1, [50.0] * 4, [1.0] * 4, [0.01, 0.01, 0.0, 0.0], pv, axp, ayp)
        st1, axp, ayp = list(r1[0]), r1[7], r1[8]
    out["scalar_drop_in_planar_model_RK4"] = {"us_per_call": (time.perf_counter() - t0) / 200 * 1e6, "calls": 200,
                                              "note": "one vehicle, one RK4 step per call through the reference's own method signature "
                                                      "(H2D of 24 doubles, one launch, D2H of 40 doubles, stream sync); the literal "
                                                      "Python reference needs ~1 ms per call (tests/golden/rollout_cfg2_sub.npz)"}
    eng.set_params(pv)                     # D = mu_max = 1 on the four wheels, the set the headline runs with
    # config 4 (one GPU's view): 1,048,576 sampled control sequences x 100 steps from one start state, running
    # cost, lowest-index argmin; the controls are drawn on the device (Philox keyed by the global rollout index)
    cfg = wl.config4_mpc(B=1 << 20)
    Bm, Nm = cfg["B"], cfg["n_steps"]
    from python_motionplanning_b200 import distributed as D
    pl = D.MpcPlanner(eng, cfg)
    for _ in range(3):
        plan = pl.plan()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        plan = pl.plan()
    ms_plan = (time.perf_counter() - t0) / 5 * 1e3
    out["mpc_sample_rollout_cost_argmin"] = {"value": Bm * Nm / (ms_plan * 1e-3), "unit": UNIT, "ms": ms_plan,
                                             "sequences": Bm, "horizon": Nm, "best_index": int(plan["index"]), "best_cost": float(plan["cost"]),
                                             "includes": "resident MpcPlanner.plan(): control sampling + rollout with running cost (broadcast start "
                                                         "state) + winner record kernels + the pinned read of the record, host-synchronised per plan"}
    del pl, plan
    # config 5 (FP64 leg): 256 tyre-coefficient sets x 4,096 manoeuvres, per-rollout parameter sets (set-major), 100 steps
    sets, st5, dl5, tq5, ps5 = wl.config5_sweep()
    p5 = mp.VehicleParameters()
    for wname in ("FL", "FR", "RL", "RR"):
        setattr(p5, "B" + wname, sets[:, 0])
        setattr(p5, "C" + wname, sets[:, 1])
        setattr(p5, "D" + wname, sets[:, 2])
    eng.set_params(p5)
    a5, b5, c5, s5 = eng.dev(st5), eng.dev(dl5), eng.dev(tq5), eng.dev(ps5, torch.int32)
    B5, N5 = a5.shape[1], 100
    for mode in ("auto", "closed_form"):
        eng.set_friction_mode(mode)
        for k in range(4):
            if k == 3:
                ev0.record()
            eng.rollout(a5, b5, c5, DT, N5, hold=N5, param_set=s5)
        ev1.record()
        torch.cuda.synchronize()
        out["param_sweep_f64" + ("" if mode == "auto" else "_closed_form")] = {
            "value": B5 * N5 / (ev0.elapsed_time(ev1) * 1e-3), "unit": UNIT, "ms": ev0.elapsed_time(ev1), "sets": int(len(sets)),
            "manoeuvres": int(B5 // len(sets)), "n_steps": N5,
            "path": "generic kernel; blocks of 64 rollouts that share a set take the tabulated step (per-set D = 1 tables)"
                    if mode == "auto" else "generic kernel, closed-form friction"}
    eng.set_friction_mode("auto")
    for k in range(4):
        if k == 3:
            ev0.record()
        eng.rollout(a5, b5, c5, DT, N5, hold=N5, param_set=s5, dtype="f32")
    ev1.record()
    torch.cuda.synchronize()
    out["param_sweep_f32"] = {"value": B5 * N5 / (ev0.elapsed_time(ev1) * 1e-3), "unit": UNIT, "ms": ev0.elapsed_time(ev1),
                              "drift_report": "profiles/r01_fp32_drift_cfg5.json", "path": "generic kernel; set-uniform blocks take the tabulated FP32 step"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.steps < 1:
        raise SystemExit("--steps must be >= 1")
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
