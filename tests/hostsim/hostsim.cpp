// Host compilation of the *device* math headers (vehicle_rhs.cuh) -- a development check only.
// It lets the algebra of the CUDA rollout kernels be compared with the oracle on a machine without
// a GPU.  It is compiled and called only by tests/test_hostsim.py; the product never loads it.
// (On the host Math<> falls back to libm, so this checks the formulas, not the MUFU/Newton paths.)
#include <cstddef>
#include <cstring>

#include "../../python_motionplanning_b200/csrc/vehicle_rhs.cuh"

using namespace b200mp;

static int g_use_table = 0;          // hostsim_set_table(1): no-log runs take the tabulated friction path (FP64 and FP32)
static double g_table_err = 0.0, g_table_err32 = 0.0;

template <typename R, bool REAR0, bool AUX>
static void run(int B, int n_steps, double dt, int hold, const double *state0, const double *delta, const double *torque,
                int tch, const double *mu, const HostParams *params, const int *param_set, int store_stride,
                double *traj, double *aux, double *state_end)
{
    const int dch = REAR0 ? 1 : 4;
    for (int r = 0; r < B; ++r) {
        DevParams<R> P = derive_params<R>(params[param_set ? param_set[r] : 0]);
        R D[4], y[10], ax, ay;
        for (int i = 0; i < 4; ++i) D[i] = mu ? (R)mu[(size_t)i * B + r] : P.Dc[i];
        for (int c = 0; c < 10; ++c) y[c] = (R)state0[(size_t)c * B + r];
        ax = (R)state0[(size_t)10 * B + r];
        ay = (R)state0[(size_t)11 * B + r];
        WheelCtrl<R> c;
        // tabulated friction as the generic device path uses it: the table of the rollout's tyre normalised to D = 1 (in the
        // precision of R), the per-wheel D (parameter set or mu_max) scaling the normal load inside rk4_step
        alignas(16) static double table[kMuTableDoubles];
        alignas(16) static float table32[kMuTableFloats];
        static double key_B = -1.0, key_C = -1.0;
        MuTableView T;
        T.c = sizeof(R) == 8 ? (const void *)table : (const void *)table32;
        const bool tab = g_use_table && !AUX;
        if (tab) {
            const HostParams &hp = params[param_set ? param_set[r] : 0];
            if (hp.B[0] != key_B || hp.C[0] != key_C) {
                g_table_err = build_mu_table(hp.B[0], hp.C[0], 1.0, table);
                g_table_err32 = build_mu_table_f32(hp.B[0], hp.C[0], 1.0, table32);
                key_B = hp.B[0];
                key_C = hp.C[0];
            }
            T.B2 = hp.B[0] * hp.B[0];
        }
        for (int n = 0; n < n_steps; ++n) {
            const size_t seg = n / hold;
            if (n % hold == 0) {
                R dl[4] = {0, 0, 0, 0};
                for (int i = 0; i < dch; ++i) dl[i] = (R)delta[(seg * dch + i) * B + r];
                set_steer<R, REAR0>(c, dl);
                R tau[4];
                for (int i = 0; i < 4; ++i) tau[i] = (R)torque[(seg * tch + (tch == 1 ? 0 : i)) * B + r];
                set_torque(c, P, tau);
            }
            R sdot[10], outs[18];
            if (tab)
                rk4_step<R, REAR0, false, false, true, true>(P, D, c, (R)dt, y, ax, ay, sdot, outs, T);
            else
                rk4_step<R, REAR0, AUX, false, !AUX>(P, D, c, (R)dt, y, ax, ay, sdot, outs);   // !AUX: speculative form + checked fallback
            if (store_stride > 0 && (n + 1) % store_stride == 0) {
                const size_t o = (size_t)((n + 1) / store_stride - 1);
                if (traj) for (int k = 0; k < 10; ++k) traj[(o * 10 + k) * B + r] = (double)y[k];
                if (AUX && aux) {
                    for (int k = 0; k < 10; ++k) aux[(o * 28 + k) * B + r] = (double)sdot[k];
                    for (int k = 0; k < 18; ++k) aux[(o * 28 + 10 + k) * B + r] = (double)outs[k];
                }
            }
        }
        for (int k = 0; k < 10; ++k) state_end[(size_t)k * B + r] = (double)y[k];
        state_end[(size_t)10 * B + r] = (double)ax;
        state_end[(size_t)11 * B + r] = (double)ay;
    }
}

extern "C" double hostsim_set_table(int on)
{
    g_use_table = on;
    return g_table_err;
}

extern "C" double hostsim_table_err_f32(void) { return g_table_err32; }

extern "C" void hostsim_rollout(int use_f32, int B, int n_steps, double dt, int hold, const double *state0,
                                const double *delta, int dch, const double *torque, int tch, const double *mu,
                                const HostParams *params, const int *param_set, int store_stride, double *traj,
                                double *aux, double *state_end)
{
#define GO(R, REAR0, AUX) run<R, REAR0, AUX>(B, n_steps, dt, hold, state0, delta, torque, tch, mu, params, param_set, store_stride, traj, aux, state_end)
    if (use_f32) {
        if (dch == 1) GO(float, true, false); else GO(float, false, false);
    } else if (aux) {
        if (dch == 1) GO(double, true, true); else GO(double, false, true);
    } else {
        if (dch == 1) GO(double, true, false); else GO(double, false, false);
    }
}
