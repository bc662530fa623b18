// K6: conformal-lattice path generation on the device (sm_100a, FP64) -- SURVEY.md §8f N1.
//
// Replaces, for P goal states at once, the two producers that sit directly before the collision test:
//   PathOptimizer.sample_spiral (+ thetaf)   reference libs/motionplanner/path_optimizer.py:109-174
//   transform_paths                          reference libs/motionplanner/local_planner.py:424-470
// so that a lattice of thousands of cubic spirals goes from its optimisation parameters [p1, p2, sf] to
// global-frame path points without 4.8 MB of paths crossing PCIe.
//
// One thread = one spiral; its samples are produced in arc-length order because the reference's x, y
// are a cumulative sum (scipy cumulative_trapezoid without an initial value, numpy cumsum), and the
// partial sums are formed in that order here so that they differ from the reference's only through the
// ulp-level difference between CUDA's and numpy's cos/sin/pow (parity contract: 1e-12 relative).
// The reference's length quirk is kept: n_samples arc-length samples give n_samples headings but
// n_samples - 1 points, and the transform keeps the first n_samples - 1 headings, so heading j belongs to
// the sample BEFORE point j (collision_checker.py:87-89 depends on it).
// Results are staged through shared memory so that the [P][n_pts] row-major outputs the collision kernel
// reads are written with coalesced stores.
#include <math.h>

#include "b200mp_internal.h"

namespace b200mp {

constexpr int kLatBlock = 64;

__global__ void __launch_bounds__(kLatBlock)
lattice_kernel(int P, int n_samples, const double *__restrict__ k1, const double *__restrict__ k2,
               const double *__restrict__ sf, const double *__restrict__ ego_x, const double *__restrict__ ego_y,
               const double *__restrict__ ego_yaw, int ego_broadcast, double *__restrict__ px, double *__restrict__ py,
               double *__restrict__ pyaw, double *__restrict__ pcos, double *__restrict__ psin, double *__restrict__ end_xy)
{
    extern __shared__ double stage[];   // [kLatBlock][n_pts + 1]: one row per thread, padded against bank conflicts
    const int n_pts = n_samples - 1;
    const int row = n_pts + 1;
    const int p = blockIdx.x * kLatBlock + threadIdx.x;
    const bool live = p < P;
    double b = 0, c = 0, d = 0, step = 0, S = 0, ex = 0, ey = 0, eyaw = 0, ce = 1, se = 0;
    if (live) {
        const double p1 = k1[p], p2 = k2[p];
        S = sf[p];
        // path_optimizer.py:150-155 with p0 = p3 = 0, in the reference's operator order
        b = -((0.0 - 9.0 * p1) + 9.0 * p2 / 2.0 - 0.0) / S;
        c = ((0.0 - 45.0 * p1 / 2.0) + 18.0 * p2 - 0.0) / (S * S);
        d = -((0.0 - 27.0 * p1 / 2.0) + 27.0 * p2 / 2.0 - 0.0) / (S * S * S);
        step = S / (double)(n_samples - 1);                       // np.linspace(0, sf): arange * step, last = sf
        const int e = ego_broadcast ? 0 : p;
        if (ego_x) {
            ex = ego_x[e];
            ey = ego_y[e];
            eyaw = ego_yaw[e];
            sincos(eyaw, &se, &ce);
        }
    }
    const double b2 = b / 2, c3 = c / 3, d4 = d / 4;
    // four passes (x, y, heading, then cos / sin of the heading), each staged and written out coalesced
    for (int pass = 0; pass < 5; ++pass) {
        double *out = pass == 0 ? px : pass == 1 ? py : pass == 2 ? pyaw : pass == 3 ? pcos : psin;
        if (!out) continue;                                       // uniform across the grid
        if (live) {
            double s_prev = 0.0, t_prev = 0.0, c_prev = 1.0, n_prev = 0.0, X = 0.0, Y = 0.0;
            double *mine = stage + threadIdx.x * row;
            for (int j = 1; j < n_samples; ++j) {
                const double s = (j == n_samples - 1) ? S : (double)j * step;
                const double s2 = s * s;
                const double t = ((0.0 * s + b2 * s2) + c3 * (s2 * s)) + d4 * (s2 * s2);   // thetaf, :109-117
                double cn, sn;
                sincos(t, &sn, &cn);
                const double ds = s - s_prev;
                X = X + ds * (cn + c_prev) / 2.0;                 // cumulative trapezoid, :172-173
                Y = Y + ds * (sn + n_prev) / 2.0;
                double v;
                if (pass == 0) {
                    v = ex + X * ce - Y * se;                     // local_planner.py:462-465
                } else if (pass == 1) {
                    v = ey + X * se + Y * ce;
                } else {
                    const double yaw = t_prev + eyaw;             // heading of the PREVIOUS sample (:466, length quirk)
                    v = pass == 2 ? yaw : (pass == 3 ? cos(yaw) : sin(yaw));
                }
                mine[j - 1] = v;
                s_prev = s;
                t_prev = t;
                c_prev = cn;
                n_prev = sn;
            }
            if (pass == 1 && end_xy) {                            // path end points for select_best (x[-1], y[-1])
                end_xy[p] = ex + X * ce - Y * se;
                end_xy[(size_t)P + p] = mine[n_pts - 1];
            }
        }
        __syncthreads();
        const int p0 = blockIdx.x * kLatBlock;
        const int n_live = min(kLatBlock, P - p0);
        for (int i = threadIdx.x; i < n_live * n_pts; i += kLatBlock) {
            const int r = i / n_pts, jj = i - r * n_pts;
            out[(size_t)p0 * n_pts + i] = stage[r * row + jj];
        }
        __syncthreads();
    }
}

int launch_lattice_f64(int device, cudaStream_t st, int P, int n_samples, const double *k1, const double *k2,
                       const double *sf, const double *ego_x, const double *ego_y, const double *ego_yaw,
                       int ego_broadcast, double *px, double *py, double *pyaw, double *pcos, double *psin, double *end_xy)
{
    (void)device;
    if (P < 0 || n_samples < 2 || n_samples > 512) {
        set_error("sample_lattice: bad sizes P=%d n_samples=%d (2..512)", P, n_samples);
        return B200MP_E_ARG;
    }
    if (P == 0) return 0;
    if (!k1 || !k2 || !sf || !px || !py) {
        set_error("sample_lattice: kappa1, kappa2, sf, px and py must be non-NULL");
        return B200MP_E_ARG;
    }
    if ((ego_x || ego_y || ego_yaw) && !(ego_x && ego_y && ego_yaw)) {
        set_error("sample_lattice: ego_x, ego_y and ego_yaw go together");
        return B200MP_E_ARG;
    }
    const size_t smem = sizeof(double) * kLatBlock * (size_t)n_samples;
    if (smem > 48 * 1024)
        B200MP_CUDA(cudaFuncSetAttribute(lattice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lattice_kernel<<<(P + kLatBlock - 1) / kLatBlock, kLatBlock, smem, st>>>(P, n_samples, k1, k2, sf, ego_x, ego_y, ego_yaw,
                                                                           ego_broadcast, px, py, pyaw, pcos, psin, end_xy);
    B200MP_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200mp
