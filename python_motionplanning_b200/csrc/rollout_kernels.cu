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
#include "b200mp_internal.h"

namespace b200mp {

template <typename R> struct RolloutDev {
    int B, n_steps, step0, hold, store_stride, torque_ch;
    R dt;
    const R *state0, *delta, *torque, *mu;
    size_t ctrl_bstride;  // B, or 0 when one control sequence is broadcast to all rollouts
    const int *param_set;
    const DevParams<R> *table;
    R *traj, *aux, *state_end, *cost;
    const R *cost_in, *cost_ref;
    R w_u, u_ref;
};

// Launch shape.  65,536 rollouts (config 2) are 2,048 warps = 13.8 per SM: with <= 144 registers per
// thread 14 warps fit on an SM and the batch is exactly one wave (no tail); above that the second wave
// runs at 15 % occupancy.  Tunables are macros so tools/kbench can sweep them.
#ifndef B200MP_ROLLOUT_BLOCK
#define B200MP_ROLLOUT_BLOCK 64
#endif
#ifndef B200MP_ROLLOUT_MAXNREG
#define B200MP_ROLLOUT_MAXNREG 0
#endif
constexpr int kRolloutBlock = B200MP_ROLLOUT_BLOCK;
#if B200MP_ROLLOUT_MAXNREG > 0
#define B200MP_ROLLOUT_BOUNDS __maxnreg__(B200MP_ROLLOUT_MAXNREG)
#else
#define B200MP_ROLLOUT_BOUNDS __launch_bounds__(B200MP_ROLLOUT_BLOCK)
#endif

template <typename R, bool REAR0, bool GENERIC, bool AUX>
__global__ void B200MP_ROLLOUT_BOUNDS
rk4_rollout_kernel(const __grid_constant__ RolloutDev<R> a, const __grid_constant__ DevParams<R> P0)
{
    const int r = blockIdx.x * kRolloutBlock + threadIdx.x;
    if (r >= a.B) return;
    const size_t B = (size_t)a.B;

    R y[10], ax, ay;
#pragma unroll
    for (int c = 0; c < 10; ++c) y[c] = a.state0[c * B + r];
    ax = a.state0[10 * B + r];
    ay = a.state0[11 * B + r];

    // parameters: constant bank (P0) or a per-rollout gather
    DevParams<R> Pl;
    R Dl[4];
    if (GENERIC) {
        Pl = a.param_set ? a.table[a.param_set[r]] : P0;
#pragma unroll
        for (int i = 0; i < 4; ++i) Dl[i] = a.mu ? a.mu[i * B + r] : Pl.Dc[i];   // vehicle_model.py:232-235
    }
    const DevParams<R> &P = GENERIC ? Pl : P0;
    const R *D = GENERIC ? Dl : P0.Dc;

    const size_t cb = a.ctrl_bstride;
    const size_t rc = cb ? (size_t)r : 0;
    const size_t cB = cb ? B : 1;
    R J = (a.cost && a.cost_in) ? a.cost_in[r] : (R)0;

    R *tp = a.traj ? a.traj + r : nullptr;
    R *xp = (AUX && a.aux) ? a.aux + r : nullptr;
    int until_store = a.store_stride;

    WheelCtrl<R> c;
    int n = 0;
    while (n < a.n_steps) {
        const int seg = (a.step0 + n) / a.hold;
        int seg_end = (seg + 1) * a.hold - a.step0;
        if (seg_end > a.n_steps) seg_end = a.n_steps;
        {   // controls of this segment
            R dl[4];
            if (REAR0) {
                dl[0] = a.delta[(size_t)seg * cB + rc];
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) dl[i] = a.delta[((size_t)seg * 4 + i) * cB + rc];
            }
            set_steer<R, REAR0>(c, dl);
            if (a.torque_ch == 1) {
                c.tq[0] = c.tq[1] = c.tq[2] = c.tq[3] = a.torque[(size_t)seg * cB + rc];
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) c.tq[i] = a.torque[((size_t)seg * 4 + i) * cB + rc];
            }
        }
#pragma unroll 1
        for (; n < seg_end; ++n) {
            R sdot[AUX ? 10 : 1], outs[AUX ? 18 : 1];
            rk4_step<R, REAR0, AUX, !GENERIC>(P, D, c, a.dt, y, ax, ay, sdot, outs);
            if (a.cost) {
                const size_t g = (size_t)(a.step0 + n);
                const R ex = y[8] - a.cost_ref[2 * g], ey = y[9] - a.cost_ref[2 * g + 1], eu = y[0] - a.u_ref;
                J = J + (ex * ex + ey * ey + a.w_u * (eu * eu));
            }
            if (a.store_stride > 0 && --until_store == 0) {
                until_store = a.store_stride;
                if (tp) {
#pragma unroll
                    for (int cidx = 0; cidx < 10; ++cidx) tp[cidx * B] = y[cidx];
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
        }
    }
#pragma unroll
    for (int cidx = 0; cidx < 10; ++cidx) a.state_end[cidx * B + r] = y[cidx];
    a.state_end[10 * B + r] = ax;
    a.state_end[11 * B + r] = ay;
    if (a.cost) a.cost[r] = J;
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
    if (g.B == 0 || g.n_steps == 0) {
        if (g.B > 0 && g.state_end != g.state0)
            B200MP_CUDA(cudaMemcpyAsync(g.state_end, g.state0, sizeof(R) * 12 * (size_t)g.B, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
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
    a.param_set = g.param_set;
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
    const dim3 grid((unsigned)((g.B + kRolloutBlock - 1) / kRolloutBlock)), block(kRolloutBlock);
    if (aux) {
        // logging mode (state_dot + outputs): one generic instantiation per steer layout
        if (rear0)
            rk4_rollout_kernel<R, true, true, true><<<grid, block, 0, st>>>(a, P0);
        else
            rk4_rollout_kernel<R, false, true, true><<<grid, block, 0, st>>>(a, P0);
    } else if (rear0) {
        if (generic)
            rk4_rollout_kernel<R, true, true, false><<<grid, block, 0, st>>>(a, P0);
        else
            rk4_rollout_kernel<R, true, false, false><<<grid, block, 0, st>>>(a, P0);
    } else {
        if (generic)
            rk4_rollout_kernel<R, false, true, false><<<grid, block, 0, st>>>(a, P0);
        else
            rk4_rollout_kernel<R, false, false, false><<<grid, block, 0, st>>>(a, P0);
    }
    B200MP_CUDA(cudaGetLastError());
    return 0;
}

int launch_rollout_f64(int device, cudaStream_t st, const B200mpRolloutArgs &a) { return launch_rollout<double>(device, st, a); }
int launch_rollout_f32(int device, cudaStream_t st, const B200mpRolloutArgs &a) { return launch_rollout<float>(device, st, a); }

// ---------------------------------------------------------------------------------------------------
// Batched single RHS evaluation: VehicleModel.planar_model (vehicle_model.py:220-425), full return list.
__global__ void __launch_bounds__(128)
planar_model_kernel(int B, const double *__restrict__ state, const double *__restrict__ torque,
                    const double *__restrict__ mu, const double *__restrict__ delta, const double *__restrict__ axay,
                    const int *__restrict__ param_set, const DevParams<double> *__restrict__ table,
                    double *__restrict__ state_dot, double *__restrict__ misc, double *__restrict__ outputs)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B) return;
    const size_t Bs = (size_t)B;
    const DevParams<double> P = table[param_set ? param_set[r] : 0];
    double y[10], D[4], dl[4], Fz[4], k[10], out[18], axc, ayc;
    WheelCtrl<double> c;
    for (int i = 0; i < 10; ++i) y[i] = state[i * Bs + r];
    for (int i = 0; i < 4; ++i) {
        D[i] = mu ? mu[i * Bs + r] : P.Dc[i];
        dl[i] = delta[i * Bs + r];
        c.tq[i] = torque[i * Bs + r];
    }
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
    planar_model_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, state, torque, mu, delta, axay, param_set, ds.table64,
                                                        state_dot, misc, outputs);
    B200MP_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200mp
