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

constexpr int kLatBlock = 32;   // 4,096 paths = 128 CTAs (one thread per spiral: the work is a serial chain of sincos)

__global__ void __launch_bounds__(kLatBlock)
lattice_kernel(int P, int n_samples, const double *__restrict__ k1, const double *__restrict__ k2,
               const double *__restrict__ sf, const double *__restrict__ ego_x, const double *__restrict__ ego_y,
               const double *__restrict__ ego_yaw, int ego_broadcast, double *__restrict__ px, double *__restrict__ py,
               double *__restrict__ pyaw, double *__restrict__ pcos, double *__restrict__ psin, double *__restrict__ end_xy,
               int single_pass)
{
    extern __shared__ double stage[];   // [outputs][kLatBlock][n_pts + 1]: one row per thread and output, padded against bank conflicts
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
    if (single_pass) {
        // ONE sweep over the samples fills a staging row per requested output (x, y, heading, cos, sin); the multi-pass form
        // below repeats the whole spiral -- 49 sincos -- for every output and is kept for sample counts whose staging rows
        // do not fit in shared memory
        double *outs[5] = {px, py, pyaw, pcos, psin};
        int slot[5], n_out = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) slot[k] = outs[k] ? n_out++ : -1;
        const size_t plane = (size_t)kLatBlock * row;
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
                const double yaw = t_prev + eyaw;                 // heading of the PREVIOUS sample (:466, length quirk)
                mine[slot[0] * plane + (j - 1)] = ex + X * ce - Y * se;   // local_planner.py:462-465 (px, py are mandatory)
                mine[slot[1] * plane + (j - 1)] = ey + X * se + Y * ce;
                if (slot[2] >= 0) mine[slot[2] * plane + (j - 1)] = yaw;
                if (slot[3] >= 0) mine[slot[3] * plane + (j - 1)] = cos(yaw);
                if (slot[4] >= 0) mine[slot[4] * plane + (j - 1)] = sin(yaw);
                s_prev = s;
                t_prev = t;
                c_prev = cn;
                n_prev = sn;
            }
            if (end_xy) {                                         // path end points for select_best (x[-1], y[-1])
                end_xy[p] = ex + X * ce - Y * se;
                end_xy[(size_t)P + p] = ey + X * se + Y * ce;
            }
        }
        __syncthreads();
        const int p0 = blockIdx.x * kLatBlock;
        const int n_live = min(kLatBlock, P - p0);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            if (slot[k] < 0) continue;
            for (int i = threadIdx.x; i < n_live * n_pts; i += kLatBlock) {
                const int r = i / n_pts, jj = i - r * n_pts;
                outs[k][(size_t)p0 * n_pts + i] = stage[slot[k] * plane + r * row + jj];
            }
        }
        return;
    }
    // one pass per output (x, y, heading, then cos / sin of the heading), each staged and written out coalesced
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
    const size_t plane = sizeof(double) * kLatBlock * (size_t)n_samples;
    const int n_out = 2 + (pyaw != nullptr) + (pcos != nullptr) + (psin != nullptr);
    const int single_pass = plane * n_out <= 160 * 1024;          // all staging rows at once (the planner's 50 samples: 64 KB)
    const size_t smem = single_pass ? plane * n_out : plane;
    if (smem > 48 * 1024)
        B200MP_CUDA(cudaFuncSetAttribute(lattice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lattice_kernel<<<(P + kLatBlock - 1) / kLatBlock, kLatBlock, smem, st>>>(P, n_samples, k1, k2, sf, ego_x, ego_y, ego_yaw,
                                                                           ego_broadcast, px, py, pyaw, pcos, psin, end_xy,
                                                                           single_pass);
    B200MP_CUDA(cudaGetLastError());
    return 0;
}


// ---------------------------------------------------------------------------------------------------
// K7: batched cubic-spiral optimisation -- the problem PathOptimizer.optimize_spiral poses
// (reference libs/motionplanner/path_optimizer.py:31-88) with the objective of :183-530,
//     J(p1, p2, sf) = fbe + 25 (fxf + fyf) + 30 ftf,
//     theta(u) = sf g(u),  g(u) = p1 G1(u) + p2 G2(u),  G1 = 4.5u^2 - 7.5u^3 + 3.375u^4,  G2 = -2.25u^2 + 6u^3 - 3.375u^4
//     x = sf/24 sum_i w_i cos theta(i/8),  y likewise with sin  (8-panel Simpson, w = 1 4 2 4 2 4 2 4 1)
//     fxf = (xf - x)^2, fyf = (yf - y)^2, ftf = (tf - theta(1))^2, fbe = sf (324 p1^2 + 324 p2^2 - 81 p1 p2)/840
// (the reference's fxf/fyf/ftf/fbe and *_grad are machine-generated expansions of exactly these), start
// [0, 0, |goal|], bounds p1, p2 in [-0.5, 0.5], sf >= |goal|.  The reference hands the problem to scipy's L-BFGS-B
// (ftol 2.2e-9, gtol 1e-5), which is not part of its sources; here one thread per goal state runs a projected
// Levenberg-Marquardt iteration (model Hessian 2 J^T W J + Hessian(fbe), active bounds frozen, step accepted when J
// decreases) to a projected-gradient tolerance of 1e-13: the same minimiser, resolved further than L-BFGS-B stops.
// The planner's acceptance test (|end - goal| <= 0.1 on the SAMPLED spiral, local_planner.py:317-323) is evaluated
// in the same thread with the sampler's trapezoid.
struct SpiralEval {
    double f, r[3], J[3][3], g[3];
};

// One goal state is worked on by a group of kSpiralLanes = 8 consecutive lanes: the eight non-trivial Simpson nodes
// u = 1/8 .. 1 (node 0 contributes cos 0 = 1 and nothing else) are evaluated one per lane -- the sincos is the expensive
// part -- and the eight weighted sums are combined with an xor butterfly inside the group.  The butterfly adds the same pairs
// in every lane (a + b and b + a round alike), so all eight lanes hold bit-identical sums and run the (replicated)
// Levenberg-Marquardt algebra in lockstep.  One thread per goal left 4,096 goals = 128 warps on 64 SMs with a serial chain
// of ~250 sincos each (ncu: 2.9 % of the warp slots active, 0.45 ms); eight lanes per goal cut the chain.
constexpr int kSpiralLanes = 8;

__device__ __forceinline__ double group_sum(double v, unsigned mask)
{
#pragma unroll
    for (int w = 1; w < kSpiralLanes; w <<= 1) v += __shfl_xor_sync(mask, v, w);
    return v;
}

__device__ __forceinline__ double spiral_objective(double p1, double p2, double sf, double xf, double yf, double tf,
                                                   SpiralEval *e, int sub, unsigned mask)
{
    const double wgt = (sub == kSpiralLanes - 1) ? 1.0 : ((sub & 1) ? 2.0 : 4.0);   // nodes 1..8: 4 2 4 2 4 2 4 1
    const double u = (sub + 1) * 0.125, u2 = u * u;
    const double G1 = u2 * (4.5 + u * (-7.5 + 3.375 * u)), G2 = u2 * (-2.25 + u * (6.0 - 3.375 * u));
    const double g = p1 * G1 + p2 * G2;
    double sn, cs;
    sincos(sf * g, &sn, &cs);
    const double sc = 1.0 + group_sum(wgt * cs, mask), ss = group_sum(wgt * sn, mask);
    double s1 = 0, s2 = 0, sg = 0, c1 = 0, c2 = 0, cg = 0;
    if (e) {
        s1 = group_sum(wgt * sn * G1, mask);
        s2 = group_sum(wgt * sn * G2, mask);
        sg = group_sum(wgt * sn * g, mask);
        c1 = group_sum(wgt * cs * G1, mask);
        c2 = group_sum(wgt * cs * G2, mask);
        cg = group_sum(wgt * cs * g, mask);
    }
    const double g_end = p1 * 0.375 + p2 * 0.375;       // G1(1) = G2(1) = 0.375
    const double k = sf / 24.0;
    const double ex = xf - k * sc, ey = yf - k * ss, et = tf - sf * g_end;
    const double q = (324.0 * p1 * p1 + 324.0 * p2 * p2 - 81.0 * p1 * p2) / 840.0;
    const double f = sf * q + 25.0 * (ex * ex + ey * ey) + 30.0 * et * et;
    if (e) {
        e->f = f;
        e->r[0] = ex; e->r[1] = ey; e->r[2] = et;
        e->J[0][0] = sf * k * s1;  e->J[0][1] = sf * k * s2;  e->J[0][2] = -(sc / 24.0 - k * sg);
        e->J[1][0] = -sf * k * c1; e->J[1][1] = -sf * k * c2; e->J[1][2] = -(ss / 24.0 + k * cg);
        e->J[2][0] = -sf * 0.375;  e->J[2][1] = -sf * 0.375;  e->J[2][2] = -g_end;
        const double wt[3] = {25.0, 25.0, 30.0};
#pragma unroll
        for (int j = 0; j < 3; ++j)
            e->g[j] = 2.0 * (wt[0] * ex * e->J[0][j] + wt[1] * ey * e->J[1][j] + wt[2] * et * e->J[2][j]);
        e->g[0] += sf * (648.0 * p1 - 81.0 * p2) / 840.0;
        e->g[1] += sf * (648.0 * p2 - 81.0 * p1) / 840.0;
        e->g[2] += q;
    }
    return f;
}

// solves A x = b for the free variables (others get x = 0); returns false when a pivot vanishes
__device__ __forceinline__ bool solve3_masked(double A[3][3], const double b[3], const bool fr[3], double x[3])
{
    double M[3][4];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) M[i][j] = (fr[i] && fr[j]) ? A[i][j] : (i == j ? 1.0 : 0.0);
        M[i][3] = fr[i] ? b[i] : 0.0;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int piv = c;
#pragma unroll
        for (int r = c + 1; r < 3; ++r)
            if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
        if (!(fabs(M[piv][c]) > 0.0)) return false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double t = M[c][j];
            M[c][j] = M[piv][j];
            M[piv][j] = t;
        }
#pragma unroll
        for (int r = c + 1; r < 3; ++r) {
            const double m = M[r][c] / M[c][c];
#pragma unroll
            for (int j = c; j < 4; ++j) M[r][j] -= m * M[c][j];
        }
    }
    x[2] = M[2][3] / M[2][2];
    x[1] = (M[1][3] - M[1][2] * x[2]) / M[1][1];
    x[0] = (M[0][3] - M[0][1] * x[1] - M[0][2] * x[2]) / M[0][0];
    return true;
}

__global__ void __launch_bounds__(64)
spiral_opt_kernel(int P, int n_samples, const double *__restrict__ gxf, const double *__restrict__ gyf,
                  const double *__restrict__ gtf, double *__restrict__ p_out, double *__restrict__ f_out,
                  int *__restrict__ it_out, unsigned char *__restrict__ valid_out)
{
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = gt / kSpiralLanes, sub = gt % kSpiralLanes;      // goal state, lane within its group
    if (r >= P) return;                                            // whole groups leave together
    const unsigned mask = 0xFFu << ((threadIdx.x & 31) & ~(kSpiralLanes - 1));
    const double xf = gxf[r], yf = gyf[r], tf = gtf[r];
    const double sf0 = sqrt(xf * xf + yf * yf);               // the straight-line distance bounds sf from below (:58, :74)
    const double lo[3] = {-0.5, -0.5, sf0}, hi[3] = {0.5, 0.5, INFINITY};
    double p[3] = {0.0, 0.0, sf0};
    SpiralEval e;
    double f = spiral_objective(p[0], p[1], p[2], xf, yf, tf, &e, sub, mask);
    double lam = 1.0e-3;
    int it = 0;
    for (; it < 100; ++it) {
        bool fr[3];
        double pgmax = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            fr[j] = !((p[j] <= lo[j] && e.g[j] > 0.0) || (p[j] >= hi[j] && e.g[j] < 0.0));
            if (fr[j]) pgmax = fmax(pgmax, fabs(e.g[j]));
        }
        if (!(pgmax > 1.0e-13 * fmax(1.0, fabs(f)))) break;
        double H[3][3];
        const double wt[3] = {25.0, 25.0, 30.0};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                H[i][j] = 2.0 * (wt[0] * e.J[0][i] * e.J[0][j] + wt[1] * e.J[1][i] * e.J[1][j] + wt[2] * e.J[2][i] * e.J[2][j]);
        const double h01 = -81.0 * p[2] / 840.0, h02 = (648.0 * p[0] - 81.0 * p[1]) / 840.0, h12 = (648.0 * p[1] - 81.0 * p[0]) / 840.0;
        H[0][0] += 648.0 * p[2] / 840.0;
        H[1][1] += 648.0 * p[2] / 840.0;
        H[0][1] += h01; H[1][0] += h01;
        H[0][2] += h02; H[2][0] += h02;
        H[1][2] += h12; H[2][1] += h12;
        bool improved = false;
        for (int tries = 0; tries < 30; ++tries) {
            double A[3][3], nb[3], step[3], q[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int j = 0; j < 3; ++j) A[i][j] = H[i][j];
                A[i][i] += lam * fmax(H[i][i], 1.0e-12);
                nb[i] = -e.g[i];
            }
            if (!solve3_masked(A, nb, fr, step)) {
                lam *= 10.0;
                continue;
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) q[j] = fmin(fmax(p[j] + step[j], lo[j]), hi[j]);
            const double fq = spiral_objective(q[0], q[1], q[2], xf, yf, tf, nullptr, sub, mask);
            if (fq < f) {
                p[0] = q[0]; p[1] = q[1]; p[2] = q[2];
                lam = fmax(lam * 0.2, 1.0e-12);
                improved = true;
                break;
            }
            lam *= 10.0;
        }
        if (!improved) break;
        f = spiral_objective(p[0], p[1], p[2], xf, yf, tf, &e, sub, mask);
    }
    if (sub == 0) {
        p_out[r] = p[0];
        p_out[(size_t)P + r] = p[1];
        p_out[2 * (size_t)P + r] = p[2];
        if (f_out) f_out[r] = f;
        if (it_out) it_out[r] = it;
    }
    if (valid_out) {
        // the planner accepts a spiral when its SAMPLED end state is within 0.1 of the goal (local_planner.py:317-323): the
        // sampler's trapezoid, X = sum_j (s_j - s_{j-1}) (cos_j + cos_{j-1}) / 2, regrouped by sample so that the lanes of the
        // group take every eighth sample: X = sum_j cos_j (s_{j+1} - s_{j-1}) / 2 with s_{-1} = s_0, s_n = s_{n-1}
        const double S = p[2];
        const double b2 = -((0.0 - 9.0 * p[0]) + 9.0 * p[1] / 2.0) / S / 2, c3 = ((0.0 - 45.0 * p[0] / 2.0) + 18.0 * p[1]) / (S * S) / 3,
                     d4 = -((0.0 - 27.0 * p[0] / 2.0) + 27.0 * p[1] / 2.0) / (S * S * S) / 4;
        const double step = S / (double)(n_samples - 1);
        const int last = n_samples - 1;
        auto arc = [&](int j) { return j <= 0 ? 0.0 : (j >= last ? S : (double)j * step); };
        double X = 0, Y = 0;
        for (int j = sub; j <= last; j += kSpiralLanes) {
            const double sj = arc(j), s2 = sj * sj;
            const double t = (b2 * s2 + c3 * (s2 * sj)) + d4 * (s2 * s2);
            double cn, sn;
            sincos(t, &sn, &cn);
            const double wj = (arc(j + 1) - arc(j - 1)) / 2.0;
            X += wj * cn;
            Y += wj * sn;
        }
        X = group_sum(X, mask);
        Y = group_sum(Y, mask);
        const double S2 = S * S;
        const double t_end = (b2 * S2 + c3 * (S2 * S)) + d4 * (S2 * S2);
        const double dx = X - xf, dy = Y - yf, dt = t_end - tf;
        if (sub == 0) valid_out[r] = !(sqrt(dx * dx + dy * dy + dt * dt) > 0.1) ? 1 : 0;
    }
}

int launch_spiral_opt_f64(int device, cudaStream_t st, int P, int n_samples, const double *xf, const double *yf,
                          const double *tf, double *p_out, double *f_out, int *it_out, unsigned char *valid_out)
{
    (void)device;
    if (P < 0 || n_samples < 2) {
        set_error("optimize_spirals: bad sizes P=%d n_samples=%d", P, n_samples);
        return B200MP_E_ARG;
    }
    if (P == 0) return 0;
    if (!xf || !yf || !tf || !p_out) {
        set_error("optimize_spirals: xf, yf, tf and p_out must be non-NULL");
        return B200MP_E_ARG;
    }
    const long long threads = (long long)P * kSpiralLanes;   // eight lanes per goal state
    spiral_opt_kernel<<<(unsigned)((threads + 63) / 64), 64, 0, st>>>(P, n_samples, xf, yf, tf, p_out, f_out, it_out, valid_out);
    B200MP_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200mp
