// K1 / K1f / K1c: batched 7-DoF RK4 rollouts, one thread per rollout (sm_100a).
//
// Replaces the per-vehicle Python loop of the reference (drive.py:141-143 around
// VehicleModel.planar_model_RK4, vehicle_model.py:427-445).  Design:
//   * thread r owns rollout r for all n_steps; its 10 states + ax_prev/ay_prev (+ running cost) stay
//     in registers across the four RK4 stages and across all steps of the launch;
//   * arrays are structure-of-arrays [component][B], so a warp's load/store of one component is 32
//     consecutive elements (256 B for FP64): trajectory writeback is fully coalesced;
//   * the parameter set travels as a __grid_constant__ kernel argument -> constant-bank operands
//     (GENERIC = false); per-rollout parameter sets / mu_max (GENERIC = true) are gathered once into
//     registers;
//   * controls are zero-order-hold segments: the steer sincos is evaluated once per segment;
//   * bound: the FP64 (or FP32) CUDA-core pipe; no tensor cores (nothing is a contraction), HBM only
//     for the 80 B/step trajectory writeback.
#pragma once
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "b200mp_internal.h"
#include "slice_sched.cuh"

namespace b200mp {

__device__ __forceinline__ void prefetch_l1(const void *p)
{
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// Trajectory stores: written once, never read by the kernel.  B200MP_TRAJ_STORE = 1 (default): streaming (evict-first)
// stores, 2: .wt, 0: plain stores.  Measured with tools/kbench (profiles/r02_k1_launch_shape.md): config 2 1.683 -> 1.670 ms,
// eight waves 2.073e10 -> 2.096e10 steps/s with either hint -- the 2.6 GB of output no longer displaces the carried state
// and the control lines in L2.
#ifndef B200MP_TRAJ_STORE
#define B200MP_TRAJ_STORE 1
#endif
template <typename R> __device__ __forceinline__ void traj_store(R *p, R v)
{
#if B200MP_TRAJ_STORE == 1
    __stcs(p, v);
#elif B200MP_TRAJ_STORE == 2
    __stwt(p, v);
#else
    *p = v;
#endif
}

// Development A/B (tools/kbench): a CTA-wide arrival-counting barrier in front of every control segment keeps the warps of
// a CTA in the same part of the step loop (the closed-loop kernel needs it for its instruction-cache footprint; whether the
// plain rollout gains from it is what the macro measures -- see profiles/r02_k1_launch_shape.md).
#ifndef B200MP_ROLLOUT_SYNC
#define B200MP_ROLLOUT_SYNC 0
#endif
__device__ __forceinline__ void rollout_cta_rendezvous()
{
#if B200MP_ROLLOUT_SYNC
    asm volatile("barrier.sync 0;" ::: "memory");
#endif
}

// Phase staggering.  The CTAs of a launch start together and do identical work, so the (two) warps that share a scheduler
// run the same part of an RK4 step at the same time: both want the FP64 pipe in the wheel / polynomial sections and both
// leave it idle in the serial sections (one wave of 37,888 rollouts x 500 steps runs at 1.93e10 steps/s, eight waves -- whose
// CTAs start at scattered times -- at 2.08e10).  With STAGGER each CTA delays its start by a fraction of a step that depends on
// its order of arrival on its SM (0, 1/2, 1/4, 3/4 of B200MP_STAGGER cycles), which costs < 2 us per launch.
#ifndef B200MP_STAGGER
#define B200MP_STAGGER 0
#endif
#if B200MP_STAGGER > 0
__device__ unsigned int g_sm_arrivals[256];   // never reset: only the order modulo 4 matters
__device__ __forceinline__ void stagger_start()
{
    __shared__ int s_delay;
    if (threadIdx.x == 0) {
        unsigned smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        const unsigned k = atomicAdd(&g_sm_arrivals[smid & 255], 1u) & 3u;
        s_delay = (int)(((k & 1u) * 2u + (k >> 1)) * (B200MP_STAGGER / 4));     // k = 0,1,2,3 -> 0, 1/2, 1/4, 3/4
    }
    __syncthreads();
    const long long t0 = clock64();
    const int d = s_delay;
    while (clock64() - t0 < d) {}
}
#else
__device__ __forceinline__ void stagger_start() {}
#endif
#ifndef B200MP_SLICE_NANOSLEEP
#define B200MP_SLICE_NANOSLEEP 100
#endif
#ifdef B200MP_SLICE_PROFILE
// development only (tools/kbench.cu): where the time of a time-sliced launch goes besides the steps themselves
__device__ unsigned long long g_slice_prof[8];   // spin, prologue, epilogue cycles (thread 0 of every CTA), items, spins > 0
#define SLICE_PROF_T(var) const long long var = clock64()
#define SLICE_PROF_ADD(i, v) do { if (threadIdx.x == 0) atomicAdd(&g_slice_prof[i], (unsigned long long)(v)); } while (0)
#else
#define SLICE_PROF_T(var)
#define SLICE_PROF_ADD(i, v)
#endif

template <typename R> struct RolloutDev {
    int B, n_steps, step0, hold, store_stride, torque_ch;
    R dt;
    const R *state0, *delta, *torque, *mu;
    size_t ctrl_bstride;  // B, or 0 when one control sequence is broadcast to all rollouts
    int state_broadcast;  // 1: state0 is [12][1], shared by every rollout
    const int *param_set;
    int n_sets;           // parameter sets in `table`: a param_set entry outside [0, n_sets) is clamped (never read out of bounds)
    const DevParams<R> *table;
    R *traj, *aux, *state_end, *cost;
    const R *cost_in, *cost_ref;
    R w_u, u_ref;
    const void *mu_table;     // friction table of set 0 in the precision of R (TAB kernels), B^2 beside it
    double mu_B2;
    const void *set_tables;   // [n_sets] D = 1 tables in the precision of R and [n_sets] B^2 (0 = none): GENERIC TAB kernels
    const double *set_B2;
};

// Launch shape.  65,536 rollouts (config 2) are 2,048 warps = 13.8 per SM: with <= 144 registers per
// thread 14 warps fit on an SM and the batch is exactly one wave (no tail); above that the second wave
// runs at 15 % occupancy.  Tunables are macros so tools/kbench can sweep them.
#ifndef B200MP_ROLLOUT_BLOCK
#if B200MP_MU_DEG5
#define B200MP_ROLLOUT_BLOCK 128   /* 96 KB friction table per CTA: two CTAs per SM */
#else
#define B200MP_ROLLOUT_BLOCK 64
#endif
#endif
#ifndef B200MP_ROLLOUT_MAXNREG
#define B200MP_ROLLOUT_MAXNREG 0
#endif
constexpr int kRolloutBlock = B200MP_ROLLOUT_BLOCK;
#ifndef B200MP_RK4_SPECULATIVE
#define B200MP_RK4_SPECULATIVE 1
#endif
constexpr bool kRolloutSpeculative = B200MP_RK4_SPECULATIVE != 0;
#ifndef B200MP_MU_CACHE_STEPS
#define B200MP_MU_CACHE_STEPS 0   /* measured: keeping the rows live across steps costs 255 registers and 13 % */
#endif
constexpr bool kCacheAcrossSteps = B200MP_MU_CACHE_STEPS != 0;
// NOTE (measured, profiles/r02_k1_launch_shape.md): __maxnreg__ REPLACES __launch_bounds__, and a kernel compiled without
// __launch_bounds__(64) is 12 % slower even with a cap of 255 registers -- register-budget experiments must use
// B200MP_ROLLOUT_MINBLOCKS (__launch_bounds__(block, min blocks per SM)), not B200MP_ROLLOUT_MAXNREG.
#ifndef B200MP_ROLLOUT_MINBLOCKS
#define B200MP_ROLLOUT_MINBLOCKS 0
#endif
#if B200MP_ROLLOUT_MAXNREG > 0
#define B200MP_ROLLOUT_BOUNDS __maxnreg__(B200MP_ROLLOUT_MAXNREG)
#elif B200MP_ROLLOUT_MINBLOCKS > 0
#define B200MP_ROLLOUT_BOUNDS __launch_bounds__(B200MP_ROLLOUT_BLOCK, B200MP_ROLLOUT_MINBLOCKS)
#else
#define B200MP_ROLLOUT_BOUNDS __launch_bounds__(B200MP_ROLLOUT_BLOCK)
#endif

// SLICED = false: one CTA = one rollout block for the whole launch (no queue code in the kernel at all).
// COST = false compiles the running-cost code out (it is otherwise carried as predicated-off instructions
// through every step of a plain rollout).
// H1 = true (front-steer fast path with hold = 1, i.e. new controls every step: sampling MPC, config 4): the controls of
// step n + 1 are turned into steer sin/cos inside the straight-line block of step n (branch-free form, checked after the
// step) and those of step n + 2 are requested, instead of a load -> sincos dependency chain in front of every step.
template <typename R, bool REAR0, bool GENERIC, bool AUX, bool SLICED, bool TAB, bool COST, bool H1 = false>
__global__ void B200MP_ROLLOUT_BOUNDS
rk4_rollout_kernel(const __grid_constant__ RolloutDev<R> a, const __grid_constant__ DevParams<R> P0,
                   const __grid_constant__ SliceSched sc)
{
    __shared__ int s_item, s_set, s_cur_set;
    constexpr int kTabWords = MuTab<R>::kBytes / 8;
    // the friction table is DYNAMIC shared memory (48 KB in FP64: with the three words above it is past the static limit);
    // launches pass MuTab<R>::kBytes for TAB kernels and 0 otherwise
    extern __shared__ __align__(16) unsigned char s_dyn[];
    double *s_mu = reinterpret_cast<double *>(s_dyn);
    MuTableView T;
    T.c = s_mu;
    T.B2 = a.mu_B2;
    if (TAB && !GENERIC) {   // one tyre for the whole launch: its table (D folded in) is staged once per CTA, 16 bytes per load
        const double2 *src = static_cast<const double2 *>(a.mu_table);
        double2 *dst = reinterpret_cast<double2 *>(s_mu);
#pragma unroll 8
        for (int i = threadIdx.x; i < kTabWords / 2; i += kRolloutBlock) dst[i] = src[i];
        __syncthreads();
    }
    if (TAB && GENERIC && threadIdx.x == 0) s_cur_set = -1;   // ordered by the barriers of the first item
    if (!GENERIC && !AUX) stagger_start();
#ifdef B200MP_SLICE_PROFILE
    unsigned long long gt0 = 0;
    const long long ck0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
#endif
    const size_t B = (size_t)a.B;
    int item = blockIdx.x;
    for (;;) {
        if (SLICED) {
            // persistent CTAs (one grid of resident CTAs, the friction table is staged once per CTA): (block, chunk)
            // items are claimed through a ticket, so an item's predecessor was always claimed earlier, by a CTA
            // that is running or done (dispatch order is not relied on) and that never waits on a later ticket
            if (threadIdx.x == 0) s_item = atomicAdd(sc.counter, 1);
            __syncthreads();
            item = s_item;
            __syncthreads();   // s_item is rewritten by the next round
            if (item >= sc.n_blocks * sc.n_chunks) break;
        }
        const int chunk_idx = SLICED ? item / sc.n_blocks : 0;
        const int blk = SLICED ? item - chunk_idx * sc.n_blocks : item;
        const int n_begin = SLICED ? chunk_idx * sc.chunk : 0;
        const int n_end = SLICED ? min(a.n_steps, n_begin + sc.chunk) : a.n_steps;
        SLICE_PROF_T(t_claim);
        if (SLICED && chunk_idx > 0) {   // wait for this block's previous time-chunk
            if (threadIdx.x == 0) {
#ifdef B200MP_SLICE_PROFILE
                if (ld_acquire(sc.done + blk) < chunk_idx) atomicAdd(&g_slice_prof[4], 1ULL);
#endif
                while (ld_acquire(sc.done + blk) < chunk_idx) __nanosleep(B200MP_SLICE_NANOSLEEP);
            }
            __syncthreads();
        }
        SLICE_PROF_T(t_ready);
        SLICE_PROF_ADD(0, t_ready - t_claim);
        SLICE_PROF_ADD(3, 1);
        const int r = blk * kRolloutBlock + threadIdx.x;
        // GENERIC + TAB: per-rollout parameter sets / mu_max.  When every rollout of this block uses the same set
        // (set-major parameter sweeps; any batch with one set and per-rollout mu_max) and that set has a friction table,
        // the table (normalised to D = 1; D scales the normal load) is staged and the block takes the tabulated step;
        // otherwise it takes the closed-form step.  The choice is uniform over the CTA.
        bool use_tab = TAB && !GENERIC;
        if (TAB && GENERIC) {
            const int myset = r < a.B ? (a.param_set ? min(max(a.param_set[r], 0), a.n_sets - 1) : 0) : -1;
            if (threadIdx.x == 0) s_set = myset;
            __syncthreads();
            const int set0 = s_set, cur = s_cur_set;
            const int uniform = __syncthreads_and(myset == set0 || myset < 0);
            const double b2 = a.set_B2[set0];
            use_tab = uniform && b2 > 0.0;
            if (use_tab) {
                if (cur != set0) {
                    const double *src = static_cast<const double *>(a.set_tables) + (size_t)set0 * kTabWords;
                    for (int i = threadIdx.x; i < kTabWords; i += kRolloutBlock) s_mu[i] = src[i];
                    if (threadIdx.x == 0) s_cur_set = set0;
                }
                T.B2 = b2;
            }
            __syncthreads();
        }
        if (r < a.B) {
            R y[10], ax, ay;
            if (!SLICED || chunk_idx == 0) {
                const size_t sB = a.state_broadcast ? 1 : B, sr = a.state_broadcast ? 0 : (size_t)r;
#pragma unroll
                for (int c = 0; c < 10; ++c) y[c] = a.state0[c * sB + sr];
                ax = a.state0[10 * sB + sr];
                ay = a.state0[11 * sB + sr];
            } else {   // carried state, written by another SM: read through L2
#pragma unroll
                for (int c = 0; c < 10; ++c) y[c] = __ldcg(a.state_end + c * B + r);
                ax = __ldcg(a.state_end + 10 * B + r);
                ay = __ldcg(a.state_end + 11 * B + r);
            }

            // parameters: constant bank (P0) or a per-rollout gather
            DevParams<R> Pl;
            R Dl[4];
            if (GENERIC) {
                Pl = a.param_set ? a.table[min(max(a.param_set[r], 0), a.n_sets - 1)] : P0;
#pragma unroll
                for (int i = 0; i < 4; ++i) Dl[i] = a.mu ? a.mu[i * B + r] : Pl.Dc[i];   // vehicle_model.py:232-235
            }
            const DevParams<R> &P = GENERIC ? Pl : P0;
            const R *D = GENERIC ? Dl : P0.Dc;

            const size_t cb = a.ctrl_bstride;
            const size_t rc = cb ? (size_t)r : 0;
            const size_t cB = cb ? B : 1;
            R J = (R)0;
            if (COST && a.cost) {
                if (SLICED && chunk_idx > 0)
                    J = __ldcg(a.cost + r);
                else if (a.cost_in)
                    J = a.cost_in[r];
            }

            const size_t out0 = a.store_stride > 0 ? (size_t)(n_begin / a.store_stride) : 0;
            R *tp = a.traj ? a.traj + out0 * 10 * B + r : nullptr;
            R *xp = (AUX && a.aux) ? a.aux + out0 * 28 * B + r : nullptr;
            int until_store = a.store_stride;

            WheelCtrl<R> c;
            MuRowCacheT<typename MuTab<R>::Row> rowc;
            rowc.k[0] = rowc.k[1] = rowc.k[2] = rowc.k[3] = -1;
            int n = n_begin;
            auto step_body = [&](int nn) {   // one RK4 step + running cost + trajectory / log store
                R sdot[AUX ? 10 : 1], outs[AUX ? 18 : 1];
                if constexpr (TAB && GENERIC) {
                    // a logging launch needs state_dot and the 18 outputs (combined slips included) only on the steps it
                    // stores: those take the closed-form logging step, every other step the tabulated one
                    const bool log_step = AUX && a.store_stride > 0 && until_store == 1;
                    if (use_tab && !log_step)
                        rk4_step<R, REAR0, false, false, true, true>(P, D, c, a.dt, y, ax, ay, sdot, outs, T);
                    else
                        rk4_step<R, REAR0, AUX, false, false, false>(P, D, c, a.dt, y, ax, ay, sdot, outs);
                } else {
                    rk4_step<R, REAR0, AUX, !GENERIC, (kRolloutSpeculative && !GENERIC && !AUX) || TAB, TAB>(P, D, c, a.dt, y, ax, ay, sdot, outs, T,
                                                                                                              (TAB && kCacheAcrossSteps) ? &rowc : nullptr);
                }
                if (COST && a.cost) {
                    const size_t g = (size_t)(a.step0 + nn);
                    const R ex = y[8] - a.cost_ref[2 * g], ey = y[9] - a.cost_ref[2 * g + 1], eu = y[0] - a.u_ref;
                    J = J + (ex * ex + ey * ey + a.w_u * (eu * eu));
                }
                if (a.store_stride > 0 && --until_store == 0) {
                    until_store = a.store_stride;
                    if (tp) {
#pragma unroll
                        for (int cidx = 0; cidx < 10; ++cidx) traj_store(tp + cidx * B, y[cidx]);
                        tp += 10 * B;
                    }
                    if (AUX && xp) {
#pragma unroll
                        for (int cidx = 0; cidx < 10; ++cidx) xp[cidx * B] = sdot[cidx];
#pragma unroll
                        for (int cidx = 0; cidx < 18; ++cidx) xp[(10 + cidx) * B] = outs[cidx];
                        xp += 28 * B;
                    }
                }
            };
            static_assert(!H1 || (REAR0 && !GENERIC && !AUX), "H1 is a variant of the front-steer fast path");
            if (H1) {
                const size_t g0 = (size_t)(a.step0 + n_begin), glast = (size_t)(a.step0 + n_end - 1);
                R dl_now = a.delta[g0 * cB + rc], tq_now = a.torque[g0 * cB + rc];
                set_steer<R, true>(c, &dl_now);
                c.tq[0] = c.tq[1] = c.tq[2] = c.tq[3] = tq_now * P.inv_Jw;
                const size_t g1 = g0 + 1 < glast ? g0 + 1 : glast;
                R dl_next = a.delta[g1 * cB + rc], tq_next = a.torque[g1 * cB + rc];
#pragma unroll 1
                for (; n < n_end; ++n) {
                    WheelCtrl<R> cn;
                    R sn, cs;
                    const R dl_used = dl_next;
                    const bool steer_ok = Math<R>::sincos_core(dl_used, &sn, &cs);
                    cn.sd[0] = cn.sd[1] = sn;
                    cn.cd[0] = cn.cd[1] = cs;
                    cn.sd[2] = cn.sd[3] = (R)0;
                    cn.cd[2] = cn.cd[3] = (R)1;
                    cn.tq[0] = cn.tq[1] = cn.tq[2] = cn.tq[3] = tq_next * P.inv_Jw;
                    const size_t g2 = (size_t)(a.step0 + n) + 2 < glast ? (size_t)(a.step0 + n) + 2 : glast;
                    dl_next = a.delta[g2 * cB + rc];
                    tq_next = a.torque[g2 * cB + rc];
                    step_body(n);
                    if (!steer_ok) set_steer<R, true>(cn, &dl_used);   // |delta| > 1e5 or NaN: the library path
                    c = cn;
                }
            }
            SLICE_PROF_T(t_loop);
            SLICE_PROF_ADD(1, t_loop - t_ready);
            while (!H1 && n < n_end) {
                rollout_cta_rendezvous();
                const int seg = (a.step0 + n) / a.hold;
                int seg_end = (seg + 1) * a.hold - a.step0;
                if (seg_end > n_end) seg_end = n_end;
                {   // controls of this segment
                    R dl[4];
                    if (REAR0) {
                        dl[0] = a.delta[(size_t)seg * cB + rc];
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) dl[i] = a.delta[((size_t)seg * 4 + i) * cB + rc];
                    }
                    set_steer<R, REAR0>(c, dl);
                    R tau[4];
                    if (a.torque_ch == 1) {
                        tau[0] = tau[1] = tau[2] = tau[3] = a.torque[(size_t)seg * cB + rc];
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) tau[i] = a.torque[((size_t)seg * 4 + i) * cB + rc];
                    }
                    set_torque(c, P, tau);
                }
                if (seg_end < n_end) {   // next segment's controls: start them towards L1 now, ten steps ahead of their use
                    const size_t nx = (size_t)(seg + 1);
                    if (REAR0) {
                        prefetch_l1(a.delta + nx * cB + rc);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) prefetch_l1(a.delta + (nx * 4 + i) * cB + rc);
                    }
                    if (a.torque_ch == 1) {
                        prefetch_l1(a.torque + nx * cB + rc);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) prefetch_l1(a.torque + (nx * 4 + i) * cB + rc);
                    }
                }
#pragma unroll 1
                for (; n < seg_end; ++n) {
                    step_body(n);
                }
            }
            SLICE_PROF_T(t_done);
#pragma unroll
            for (int cidx = 0; cidx < 10; ++cidx) a.state_end[cidx * B + r] = y[cidx];
            a.state_end[10 * B + r] = ax;
            a.state_end[11 * B + r] = ay;
            if (COST && a.cost) a.cost[r] = J;
#ifdef B200MP_SLICE_PROFILE
            __threadfence();
            __syncthreads();
            SLICE_PROF_ADD(2, clock64() - t_done);
            SLICE_PROF_ADD(5, t_done - t_loop);
#endif
#if B200MP_ROLLOUT_SYNC
        } else if (!H1) {   // no rollout: keep the rendezvous count of the block (one per control segment)
            for (int n = n_begin; n < n_end;) {
                rollout_cta_rendezvous();
                int seg_end = ((a.step0 + n) / a.hold + 1) * a.hold - a.step0;
                n = seg_end > n_end ? n_end : seg_end;
            }
#endif
        }
        if (SLICED && chunk_idx + 1 < sc.n_chunks) {   // publish the carried state of this block
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) st_release(sc.done + blk, chunk_idx + 1);
        }
        if (!SLICED) break;
    }
#ifdef B200MP_SLICE_PROFILE
    if (threadIdx.x == 0) {   // SM clock actually delivered to this CTA: cycles elapsed per nanosecond of global time
        unsigned long long gt1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
        atomicAdd(&g_slice_prof[6], (unsigned long long)(clock64() - ck0));
        atomicAdd(&g_slice_prof[7], gt1 - gt0);
    }
#endif
}

// zero-step launch with a broadcast start state: state_end[c][r] = state0[c] (the state "passes through")
template <typename R> __global__ void broadcast_state_kernel(int B, const R *__restrict__ s0, R *__restrict__ out)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B) return;
#pragma unroll
    for (int c = 0; c < 12; ++c) out[(size_t)c * B + r] = s0[c];
}

// Resident CTAs per device for one kernel instantiation and the opt-in for dynamic shared memory beyond the default 48 KB
// carve-out: both are queried / set ONCE per (function, shared-memory size) and cached -- the two driver calls cost tens of
// microseconds, which is per-launch latency for the scalar drop-in path and several per cent of a one-wave launch.
struct KernelInfo {
    const void *fn;
    size_t smem;
    int resident;
};
template <typename K> static int kernel_info(K kernel, size_t smem, int *resident_out)
{
    static std::mutex mu;
    static std::vector<KernelInfo> cache;
    int dev = 0;
    B200MP_CUDA(cudaGetDevice(&dev));
    const void *key = reinterpret_cast<const void *>(kernel);
    {
        std::lock_guard<std::mutex> lock(mu);
        for (const KernelInfo &k : cache)
            if (k.fn == key && k.smem == smem + ((size_t)dev << 48)) {
                *resident_out = k.resident;
                return 0;
            }
    }
    if (smem) B200MP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 0, occ = 0;
    B200MP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    B200MP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kRolloutBlock, smem));
    *resident_out = sms * (occ > 0 ? occ : 1);
    std::lock_guard<std::mutex> lock(mu);
    cache.push_back(KernelInfo{key, smem + ((size_t)dev << 48), *resident_out});
    return 0;
}

// Chooses the time slicing of a launch (see SliceSched) and starts the kernel.
template <typename R, typename K>
static int start_rollout(K kernel, K kernel_sliced, int device, cudaStream_t st, const RolloutDev<R> &a,
                         const DevParams<R> &P0, size_t smem)
{
    const int n_blocks = (a.B + kRolloutBlock - 1) / kRolloutBlock;
#ifdef B200MP_SLICE_PROFILE
    if (const char *ev = getenv("B200MP_SMEM_PAD")) smem += (size_t)atoi(ev);   // development: fewer resident CTAs per SM
#endif
    int resident = 0, resident_plain = 0;
    int rc = kernel_info(kernel, smem, &resident_plain);
    if (!rc) rc = kernel_info(kernel_sliced, smem, &resident);
    if (rc) return rc;
    SliceSched sc{nullptr, nullptr, n_blocks, 1, a.n_steps};
    // slice only when the batch is more than one wave but too few waves for the tail to vanish
    const int min_chunk = 20;
    const size_t sched_bytes = sizeof(int) * ((size_t)n_blocks + 1);
    if (n_blocks > resident && n_blocks < 8 * resident && a.n_steps >= 2 * min_chunk && sched_bytes <= kSchedSlotBytes) {
        long long want = (12LL * resident + n_blocks - 1) / n_blocks;      // ~12 rounds of items
        long long max_chunks = a.n_steps / min_chunk;
        if (want > max_chunks) want = max_chunks;
        int chunk = (int)((a.n_steps + want - 1) / want);
        if (a.hold > 1 && a.hold < chunk) chunk = (chunk + a.hold - 1) / a.hold * a.hold;   // whole ZOH segments
        if (a.store_stride > 1) chunk = (chunk + a.store_stride - 1) / a.store_stride * a.store_stride;
        const int n_chunks = (a.n_steps + chunk - 1) / chunk;
        if (n_chunks > 1) {
            sc.n_chunks = n_chunks;
            sc.chunk = chunk;
        }
    }
    int grid = n_blocks;
    void *sched_mem = nullptr;
    cudaEvent_t sched_done = nullptr;
    if (sc.n_chunks > 1) {
        rc = acquire_sched_slot(device, &sched_mem, &sched_done);
        if (rc) return rc;
        B200MP_CUDA(cudaMemsetAsync(sched_mem, 0, sched_bytes, st));
        sc.counter = (int *)sched_mem;
        sc.done = (int *)sched_mem + 1;
        grid = n_blocks * sc.n_chunks < resident ? n_blocks * sc.n_chunks : resident;
    }
    if (sc.n_chunks > 1)
        kernel_sliced<<<grid, kRolloutBlock, smem, st>>>(a, P0, sc);
    else
        kernel<<<grid, kRolloutBlock, smem, st>>>(a, P0, sc);
    cudaError_t e = cudaGetLastError();
    if (sched_mem) {
        cudaError_t e2 = cudaEventRecord(sched_done, st);
        if (e == cudaSuccess) e = e2;
    }
    if (e != cudaSuccess) return cuda_fail(e, "rk4_rollout_kernel launch");
    return 0;
}

template <typename R>
static int launch_rollout(int device, cudaStream_t st, const B200mpRolloutArgs &g)
{
    if (g.B < 0 || g.n_steps < 0 || g.hold < 1 || g.step0 < 0 || g.store_stride < 0) {
        set_error("rk4_rollout: bad sizes B=%d n_steps=%d step0=%d hold=%d store_stride=%d", g.B, g.n_steps, g.step0,
                  g.hold, g.store_stride);
        return B200MP_E_ARG;
    }
    if (!g.state0 || !g.state_end || !g.delta || !g.torque) {
        set_error("rk4_rollout: state0, state_end, delta and torque must be non-NULL");
        return B200MP_E_ARG;
    }
    if ((g.delta_ch != 1 && g.delta_ch != 4) || (g.torque_ch != 1 && g.torque_ch != 4)) {
        set_error("rk4_rollout: delta_ch and torque_ch must be 1 or 4 (got %d, %d)", g.delta_ch, g.torque_ch);
        return B200MP_E_ARG;
    }
    if (g.store_stride > 0 && g.step0 % g.store_stride != 0) {
        set_error("rk4_rollout: step0 must be a multiple of store_stride");
        return B200MP_E_ARG;
    }
    if (g.cost && !g.cost_ref) {
        set_error("rk4_rollout: cost requires cost_ref");
        return B200MP_E_ARG;
    }
    DeviceState &ds = dev_state(device);
    if (ds.n_sets < 1) {
        set_error("rk4_rollout: no parameter table on device %d (call b200mp_set_params first)", device);
        return B200MP_E_PARAMS;
    }
    if (g.friction_override < 0 || g.friction_override > 2 || (g.state_broadcast != 0 && g.state_broadcast != 1)) {
        set_error("rk4_rollout: friction_override must be 0..2 and state_broadcast 0 or 1 (got %d, %d)", g.friction_override,
                  g.state_broadcast);
        return B200MP_E_ARG;
    }
    if (g.B == 0 || g.n_steps == 0) {
        if (g.B > 0 && g.state_end != g.state0) {
            if (g.state_broadcast) {
                broadcast_state_kernel<R><<<(g.B + 255) / 256, 256, 0, st>>>(g.B, (const R *)g.state0, (R *)g.state_end);
                B200MP_CUDA(cudaGetLastError());
            } else
                B200MP_CUDA(cudaMemcpyAsync(g.state_end, g.state0, sizeof(R) * 12 * (size_t)g.B, cudaMemcpyDeviceToDevice, st));
        }
        return 0;
    }
    const int fmode = g.friction_override ? g.friction_override - 1 : friction_mode();
    RolloutDev<R> a;
    a.B = g.B;
    a.n_steps = g.n_steps;
    a.step0 = g.step0;
    a.hold = g.hold;
    a.store_stride = (g.traj || g.aux) ? g.store_stride : 0;
    a.torque_ch = g.torque_ch;
    a.dt = (R)g.dt;
    a.state0 = (const R *)g.state0;
    a.delta = (const R *)g.delta;
    a.torque = (const R *)g.torque;
    a.mu = (const R *)g.mu;
    a.ctrl_bstride = g.ctrl_broadcast ? 0 : (size_t)g.B;
    a.state_broadcast = g.state_broadcast;
    a.param_set = g.param_set;
    a.n_sets = ds.n_sets;
    a.table = sizeof(R) == 8 ? (const DevParams<R> *)ds.table64 : (const DevParams<R> *)ds.table32;
    a.traj = (R *)g.traj;
    a.aux = (R *)g.aux;
    a.state_end = (R *)g.state_end;
    a.cost = (R *)g.cost;
    a.cost_in = (const R *)g.cost_in;
    a.cost_ref = (const R *)g.cost_ref;
    a.w_u = (R)g.w_u;
    a.u_ref = (R)g.u_ref;
    const DevParams<R> P0 = derive_params<R>(ds.set0);

    const bool rear0 = g.delta_ch == 1;
    // the constant-bank fast path assumes one (B, C, D) triple for the four tyres (the reference's
    // VehicleParameters copies the FL values to every wheel, vehicle_model.py:41-54)
    bool uniform_tyres = true;
    for (int i = 1; i < 4; ++i)
        uniform_tyres = uniform_tyres && ds.set0.B[i] == ds.set0.B[0] && ds.set0.C[i] == ds.set0.C[0] &&
                        ds.set0.D[i] == ds.set0.D[0];
    const bool generic = g.mu != nullptr || g.param_set != nullptr || !uniform_tyres;
    const bool aux = g.aux != nullptr;
    // tabulated friction: FP64 fast path only, when set_params could build the table of set 0
    const bool have_table = sizeof(R) == 8 ? (ds.mu_table && ds.mu_table_B2 > 0.0) : ds.mu_table_f32_ok;
    const bool tab = !generic && !aux && have_table && fmode == B200MP_FRICTION_AUTO;
    a.mu_table = sizeof(R) == 8 ? (const void *)ds.mu_table : (const void *)ds.mu_table_f32;
    a.mu_B2 = ds.mu_table_B2;
    a.set_tables = sizeof(R) == 8 ? (const void *)ds.set_tables : (const void *)ds.set_tables_f32;
    a.set_B2 = ds.set_B2;
    // per-set tables: generic FP64 launches whose blocks turn out to be set-uniform take the tabulated step
    const bool set_tables_ok = ds.set_tables && ds.set_tables_f32 && ds.set_tables_n == ds.n_sets &&
                               fmode == B200MP_FRICTION_AUTO;
    const bool tabg = generic && !aux && set_tables_ok;
    // FP64 logging launches that store a subset of the steps: tabulated step between the stored ones
    const bool tab_aux = aux && sizeof(R) == 8 && set_tables_ok && g.store_stride > 1;
#define B200MP_START2(REAR0, GENERIC, AUX, TAB, COST) \
    start_rollout<R>(rk4_rollout_kernel<R, REAR0, GENERIC, AUX, false, TAB, COST>, rk4_rollout_kernel<R, REAR0, GENERIC, AUX, true, TAB, COST>, device, st, a, P0, \
                     (TAB) ? (size_t)MuTab<R>::kBytes : (size_t)0)
#define B200MP_START_H1(TAB, COST) \
    start_rollout<R>(rk4_rollout_kernel<R, true, false, false, false, TAB, COST, true>, rk4_rollout_kernel<R, true, false, false, true, TAB, COST, true>, device, st, a, P0, \
                     (TAB) ? (size_t)MuTab<R>::kBytes : (size_t)0)
#define B200MP_START(REAR0, GENERIC, AUX, TAB) \
    ((AUX) || g.cost ? B200MP_START2(REAR0, GENERIC, AUX, TAB, true) : B200MP_START2(REAR0, GENERIC, AUX, TAB, (AUX)))
    if (aux && tab_aux) return rear0 ? B200MP_START(true, true, true, (sizeof(R) == 8)) : B200MP_START(false, true, true, (sizeof(R) == 8));
    if (aux)   // logging mode (state_dot + outputs): one generic instantiation per steer layout
        return rear0 ? B200MP_START(true, true, true, false) : B200MP_START(false, true, true, false);
    if (generic && tabg) return rear0 ? B200MP_START(true, true, false, true) : B200MP_START(false, true, false, true);
    if (generic) return rear0 ? B200MP_START(true, true, false, false) : B200MP_START(false, true, false, false);
    // per-step controls on the front-steer fast path (sampling MPC): software-pipelined control preparation
    if (rear0 && !generic && !aux && g.hold == 1 && g.torque_ch == 1) {
        if (tab) return g.cost ? B200MP_START_H1(true, true) : B200MP_START_H1(true, false);
        return g.cost ? B200MP_START_H1(false, true) : B200MP_START_H1(false, false);
    }
    if (tab) return rear0 ? B200MP_START(true, false, false, true) : B200MP_START(false, false, false, true);
    return rear0 ? B200MP_START(true, false, false, false) : B200MP_START(false, false, false, false);
#undef B200MP_START
#undef B200MP_START2
#undef B200MP_START_H1
}

// One translation unit per precision (rollout_kernels_f64.cu / rollout_kernels_f32.cu define the macro): the two sets
// of ~40 kernel instantiations compile in parallel.
#if defined(B200MP_ROLLOUT_F64) && defined(B200MP_SLICE_PROFILE)
extern "C" int b200mp_debug_slice_prof(unsigned long long *out8, int reset)
{
    if (out8) cudaMemcpyFromSymbol(out8, g_slice_prof, sizeof(unsigned long long) * 8);
    if (reset) {
        unsigned long long z[8] = {0};
        cudaMemcpyToSymbol(g_slice_prof, z, sizeof(z));
    }
    return 0;
}
#endif
#if defined(B200MP_ROLLOUT_F64)
int launch_rollout_f64(int device, cudaStream_t st, const B200mpRolloutArgs &a) { return launch_rollout<double>(device, st, a); }
#endif
#if defined(B200MP_ROLLOUT_F32)
int launch_rollout_f32(int device, cudaStream_t st, const B200mpRolloutArgs &a) { return launch_rollout<float>(device, st, a); }
#endif

#if defined(B200MP_ROLLOUT_F64)
// ---------------------------------------------------------------------------------------------------
// Batched single RHS evaluation: VehicleModel.planar_model (vehicle_model.py:220-425), full return list.
__global__ void __launch_bounds__(128)
planar_model_kernel(int B, const double *__restrict__ state, const double *__restrict__ torque,
                    const double *__restrict__ mu, const double *__restrict__ delta, const double *__restrict__ axay,
                    const int *__restrict__ param_set, int n_sets, const DevParams<double> *__restrict__ table,
                    double *__restrict__ state_dot, double *__restrict__ misc, double *__restrict__ outputs)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B) return;
    const size_t Bs = (size_t)B;
    const DevParams<double> P = table[param_set ? min(max(param_set[r], 0), n_sets - 1) : 0];   // clamped: never out of bounds
    double y[10], D[4], dl[4], tau[4], Fz[4], k[10], out[18], axc, ayc;
    WheelCtrl<double> c;
    for (int i = 0; i < 10; ++i) y[i] = state[i * Bs + r];
    for (int i = 0; i < 4; ++i) {
        D[i] = mu ? mu[i * Bs + r] : P.Dc[i];
        dl[i] = delta[i * Bs + r];
        tau[i] = torque[i * Bs + r];
    }
    set_torque(c, P, tau);
    set_steer<double, false>(c, dl);
    normal_loads(P, axay[r], axay[Bs + r], Fz);
    double sy, cy;
    Math<double>::sincos(y[7], &sy, &cy);
    planar_rhs<double, false, true, false>(P, D, y, sy, cy, c, Fz, k, axc, ayc, out);
    if (state_dot)
        for (int i = 0; i < 10; ++i) state_dot[i * Bs + r] = k[i];
    if (misc) {
        misc[0 * Bs + r] = y[0] * cy - y[1] * sy;       // vx  :410
        misc[1 * Bs + r] = y[1] * sy + y[0] * cy;       // vy  :411 [sic], reproduced
        misc[2 * Bs + r] = axc * cy - ayc * sy;         // ax  :415
        misc[3 * Bs + r] = axc * sy + ayc * cy;         // ay  :416
        misc[4 * Bs + r] = axc;
        misc[5 * Bs + r] = ayc;
    }
    if (outputs)
        for (int i = 0; i < 18; ++i) outputs[i * Bs + r] = out[i];
}

int launch_planar_model_f64(int device, cudaStream_t st, int B, const double *state, const double *torque,
                            const double *mu, const double *delta, const double *axay, const int *param_set,
                            double *state_dot, double *misc, double *outputs)
{
    if (B < 0 || !state || !torque || !delta || !axay) {
        set_error("planar_model: state, torque, delta, axay must be non-NULL and B >= 0");
        return B200MP_E_ARG;
    }
    DeviceState &ds = dev_state(device);
    if (ds.n_sets < 1) {
        set_error("planar_model: no parameter table on device %d (call b200mp_set_params first)", device);
        return B200MP_E_PARAMS;
    }
    if (B == 0) return 0;
    planar_model_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, state, torque, mu, delta, axay, param_set, ds.n_sets, ds.table64,
                                                        state_dot, misc, outputs);
    B200MP_CUDA(cudaGetLastError());
    return 0;
}

#endif   // B200MP_ROLLOUT_F64

}  // namespace b200mp
