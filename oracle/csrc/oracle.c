/* Plain-C CPU oracle for the b200 motion-planning hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A scalar restatement, in the reference's operator order, of
 *   VehicleModel.planar_model            /root/reference/libs/vehicle_model/vehicle_model.py:220-425
 *   VehicleModel.planar_model_RK4        /root/reference/libs/vehicle_model/vehicle_model.py:427-445
 *   CollisionChecker.collision_check     /root/reference/libs/motionplanner/collision_checker.py:32-117
 *   CollisionChecker.select_best_path_index  .../collision_checker.py:134-203
 * used (a) where the NumPy oracle is too slow for full-size parity and (b) as the multi-threaded
 * CPU baseline that bench.py times beside the GPU.  It is never linked into, loaded by or called
 * from the product library.
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off -fno-fast-math (oracle/build.py).
 * -ffp-contract=off matters: scipy's cdist and numpy's scalar arithmetic round every product.
 * Differences from the literal reference are at the ulp level only: glibc atan vs numpy's SIMD
 * arctan (1 ulp in 0.06 % of inputs) and x*x vs the scalar reference's pow(x, 2) (1 ulp in 0.08 %),
 * SURVEY.md Appendix B.  Collision booleans and selected indices are bit-exact.
 *
 * Pinned against tests/golden (literal reference outputs) by tests/test_oracle_pinned.py.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define G_ACC 9.81 /* vehicle_model.py:230 */

typedef struct {
    double m, a, b, Izz, Jw, hg, T, wL, wR, rw;
    double B[4], C[4], D[4]; /* D is used only when no mu_max is supplied */
} oracle_params;

/* One RHS evaluation.  y = [U V wz wFL wFR wRL wRR yaw x y].
 * misc = [vx vy ax ay axc ayc]; outputs = [fx(4) fy(4) Fz(4) s(4) fxtFL fytFL]. */
void oracle_planar_model(const double *y, const double *tq, const double *mu, const double *delta,
                         const oracle_params *p, double ax_prev, double ay_prev,
                         double *state_dot, double *misc, double *outputs)
{
    const double U = y[0], V = y[1], wz = y[2], yaw = y[7];
    const double g = G_ACC;
    /* :245-253 */
    const double fFz0 = p->b / (p->a + p->b) * p->m * g / 2;
    const double fRz0 = p->a / (p->a + p->b) * p->m * g / 2;
    const double DfzxL = p->m * p->hg * p->wR / ((p->a + p->b) * (p->wL + p->wR));
    const double DfzxR = p->m * p->hg * p->wL / ((p->a + p->b) * (p->wL + p->wR));
    const double DfzyF = p->m * p->hg * p->b / ((p->a + p->b) * (p->wL + p->wR));
    const double DfzyR = p->m * p->hg * p->a / ((p->a + p->b) * (p->wL + p->wR));
    double Fz[4];
    Fz[0] = fFz0 - DfzxL * ax_prev - DfzyF * ay_prev; /* :255-258 */
    Fz[1] = fFz0 - DfzxR * ax_prev + DfzyF * ay_prev;
    Fz[2] = fRz0 + DfzxL * ax_prev - DfzyR * ay_prev;
    Fz[3] = fRz0 + DfzxR * ax_prev + DfzyR * ay_prev;

    double vxc[4], vyc[4]; /* :261-271 */
    vxc[0] = U - p->T * wz / 2; vxc[1] = U + p->T * wz / 2;
    vxc[2] = U - p->T * wz / 2; vxc[3] = U + p->T * wz / 2;
    vyc[0] = V + p->a * wz;     vyc[1] = V + p->a * wz;
    vyc[2] = V - p->b * wz;     vyc[3] = V - p->b * wz;

    double fx[4], fy[4], fxt[4], fyt[4], sl[4];
    for (int i = 0; i < 4; ++i) {
        const double cd = cos(delta[i]), sd = sin(delta[i]);
        const double D = mu ? mu[i] : p->D[i];                 /* :232-235 */
        const double vx = vxc[i] * cd + vyc[i] * sd;           /* :274-281 */
        const double vy = -vxc[i] * sd + vyc[i] * cd;
        const double sx = p->rw * y[3 + i] / vx - 1;           /* :284-287 */
        const double sy = -vy / fabs(vx);                      /* :290-293 */
        const double s = sqrt(sx * sx + sy * sy);              /* :296-299 */
        const double muc = D * sin(p->C[i] * atan(p->B[i] * s)); /* :303-306 */
        double mux, muy;
        if (s != 0) {                                          /* :309-348 */
            mux = sx * muc / s;
            muy = sy * muc / s;
        } else {
            mux = D * sin(p->C[i] * atan(p->B[i] * sx));
            muy = D * sin(p->C[i] * atan(p->B[i] * sy));
        }
        fxt[i] = mux * Fz[i];                                  /* :351-360 */
        fyt[i] = muy * Fz[i];
        fx[i] = fxt[i] * cd - fyt[i] * sd;                     /* :363-373 */
        fy[i] = fxt[i] * sd + fyt[i] * cd;
        sl[i] = s;
    }
    const double cy = cos(yaw), sy_ = sin(yaw);
    const double U_dot = 1 / p->m * (fx[0] + fx[1] + fx[2] + fx[3]) + V * wz;   /* :376-385 */
    const double V_dot = 1 / p->m * (fy[0] + fy[1] + fy[2] + fy[3]) - U * wz;
    const double wz_dot = 1 / p->Izz * (p->a * (fy[0] + fy[1]) - p->b * (fy[2] + fy[3])
                                        + p->T / 2 * (fx[1] - fx[0] + fx[3] - fx[2]));
    state_dot[0] = U_dot;
    state_dot[1] = V_dot;
    state_dot[2] = wz_dot;
    state_dot[3] = (tq[0] - p->rw * fxt[0]) / p->Jw;
    state_dot[4] = (tq[1] - p->rw * fxt[1]) / p->Jw;
    state_dot[5] = (tq[2] - p->rw * fx[2]) / p->Jw;   /* chassis-frame force, reference quirk */
    state_dot[6] = (tq[3] - p->rw * fx[3]) / p->Jw;
    state_dot[7] = wz;
    state_dot[8] = U * cy - V * sy_;
    state_dot[9] = U * sy_ + V * cy;
    const double axc = U_dot - V * wz, ayc = V_dot + U * wz;   /* :413-414 */
    misc[0] = U * cy - V * sy_;
    misc[1] = V * sy_ + U * cy;                                /* [sic] :411 */
    misc[2] = axc * cy - ayc * sy_;
    misc[3] = axc * sy_ + ayc * cy;
    misc[4] = axc;
    misc[5] = ayc;
    if (outputs) {
        for (int i = 0; i < 4; ++i) {
            outputs[i] = fx[i]; outputs[4 + i] = fy[i]; outputs[8 + i] = Fz[i]; outputs[12 + i] = sl[i];
        }
        outputs[16] = fxt[0];
        outputs[17] = fyt[0];
    }
}

/* One RK4 step (:427-445).  y is updated in place; axay holds ax_prev, ay_prev in and axc, ayc out.
 * state_dot / outputs (RK4-weighted means) may be NULL. */
void oracle_rk4_step(double *y, const double *tq, const double *mu, const double *delta,
                     const oracle_params *p, double h, double *axay, double *state_dot, double *outputs)
{
    double K[4][10], M[4][6], O[4][18], ys[10];
    const double ax = axay[0], ay = axay[1];
    oracle_planar_model(y, tq, mu, delta, p, ax, ay, K[0], M[0], outputs ? O[0] : NULL);
    for (int c = 0; c < 10; ++c) ys[c] = y[c] + h / 2 * K[0][c];
    oracle_planar_model(ys, tq, mu, delta, p, ax, ay, K[1], M[1], outputs ? O[1] : NULL);
    for (int c = 0; c < 10; ++c) ys[c] = y[c] + h / 2 * K[1][c];
    oracle_planar_model(ys, tq, mu, delta, p, ax, ay, K[2], M[2], outputs ? O[2] : NULL);
    for (int c = 0; c < 10; ++c) ys[c] = y[c] + h * K[2][c];
    oracle_planar_model(ys, tq, mu, delta, p, ax, ay, K[3], M[3], outputs ? O[3] : NULL);
    const double h6 = 1.0 / 6 * h;
    for (int c = 0; c < 10; ++c) {
        const double sum = K[0][c] + 2 * K[1][c] + 2 * K[2][c] + K[3][c];
        y[c] = y[c] + h6 * sum;
        if (state_dot) state_dot[c] = sum / 6;
    }
    if (outputs)
        for (int c = 0; c < 18; ++c) outputs[c] = (O[0][c] + 2 * O[1][c] + 2 * O[2][c] + O[3][c]) / 6;
    axay[0] = (M[0][4] + 2 * M[1][4] + 2 * M[2][4] + M[3][4]) / 6;
    axay[1] = (M[0][5] + 2 * M[1][5] + 2 * M[2][5] + M[3][5]) / 6;
}

/* Batched open-loop rollout: the loop Car.drive runs per vehicle (drive.py:141-143).
 * Layouts are structure-of-arrays with the rollout index fastest:
 *   state0    [12][B]   (10 states, ax_prev, ay_prev)
 *   delta     [n_seg][dch][B]  dch = 1 (front steer, FL = FR, rear 0) or 4
 *   torque    [n_seg][tch][B]  tch = 1 (equal on 4 wheels) or 4
 *   mu        [4][B] or NULL (-> params D)
 *   param_set [B] or NULL (-> set 0)
 *   traj      [n_steps/store_stride][10][B] or NULL;  aux [n_out][28][B] (state_dot 10 + outputs 18) or NULL
 *   state_end [12][B]
 *   cost      [B] or NULL: J = sum_t (x_t-xr_t)^2 + (y_t-yr_t)^2 + w_u (U_t - u_ref)^2, cost_ref = [n_steps][2]
 * ctrl_bstride 0 broadcasts one control sequence to every rollout.  Returns steps executed. */
long oracle_rollout(int B, int n_steps, double dt, int hold, const double *state0,
                    const double *delta, int dch, const double *torque, int tch, int ctrl_bstride,
                    const double *mu, const oracle_params *params, const int *param_set,
                    int store_stride, double *traj, double *aux, double *state_end,
                    double *cost, const double *cost_ref, double w_u, double u_ref, int nthreads)
{
    if (nthreads <= 0) nthreads = 1;
    const size_t cB = ctrl_bstride ? (size_t)B : 1;
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int r = 0; r < B; ++r) {
        double y[10], axay[2], tq[4], dl[4], m4[4], sd[10], out[18];
        for (int c = 0; c < 10; ++c) y[c] = state0[(size_t)c * B + r];
        axay[0] = state0[(size_t)10 * B + r];
        axay[1] = state0[(size_t)11 * B + r];
        const oracle_params *p = &params[param_set ? param_set[r] : 0];
        if (mu) for (int i = 0; i < 4; ++i) m4[i] = mu[(size_t)i * B + r];
        const size_t rc = ctrl_bstride ? (size_t)r : 0;
        double J = 0.0;
        for (int n = 0; n < n_steps; ++n) {
            const size_t seg = (size_t)(n / hold);
            if (dch == 1) {
                dl[0] = dl[1] = delta[seg * cB + rc]; dl[2] = dl[3] = 0.0;
            } else
                for (int i = 0; i < 4; ++i) dl[i] = delta[(seg * 4 + i) * cB + rc];
            if (tch == 1) {
                tq[0] = tq[1] = tq[2] = tq[3] = torque[seg * cB + rc];
            } else
                for (int i = 0; i < 4; ++i) tq[i] = torque[(seg * 4 + i) * cB + rc];
            oracle_rk4_step(y, tq, mu ? m4 : NULL, dl, p, dt, axay, aux ? sd : NULL, aux ? out : NULL);
            if (cost) {
                const double ex = y[8] - cost_ref[2 * n], ey = y[9] - cost_ref[2 * n + 1], eu = y[0] - u_ref;
                J = J + (ex * ex + ey * ey + w_u * (eu * eu));
            }
            if (store_stride > 0 && (n + 1) % store_stride == 0) {
                const size_t o = (size_t)((n + 1) / store_stride - 1);
                if (traj) for (int c = 0; c < 10; ++c) traj[(o * 10 + c) * B + r] = y[c];
                if (aux) {
                    for (int c = 0; c < 10; ++c) aux[(o * 28 + c) * B + r] = sd[c];
                    for (int c = 0; c < 18; ++c) aux[(o * 28 + 10 + c) * B + r] = out[c];
                }
            }
        }
        for (int c = 0; c < 10; ++c) state_end[(size_t)c * B + r] = y[c];
        state_end[(size_t)10 * B + r] = axay[0];
        state_end[(size_t)11 * B + r] = axay[1];
        if (cost) cost[r] = J;
    }
    return (long)B * n_steps;
}

/* Circle-offset collision test for P paths (collision_checker.py:66-113).
 * px, py, pcos, psin are [P][n_pts] (cos/sin of yaw_j computed by the caller with numpy, as the
 * reference does); obstacles [M][2].  free_out[p] = 1 when collision-free.  With early_exit the
 * loop order and the break are the reference's (point-major, obstacle-minor); min_clear (may be NULL)
 * needs early_exit = 0.  Returns the number of circle-vs-point tests actually executed. */
long long oracle_collision_check(int P, int n_pts, int n_circ, const double *off, const double *rad,
                                 const double *px, const double *py, const double *pcos, const double *psin,
                                 int M, const double *obs, int early_exit, unsigned char *free_out,
                                 double *min_clear, int nthreads)
{
    long long tests = 0;
    if (nthreads <= 0) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads) reduction(+ : tests)
    for (int p = 0; p < P; ++p) {
        int is_free = 1;
        double clr = INFINITY;
        for (int j = 0; j < n_pts && (is_free || !early_exit); ++j) {
            const size_t q = (size_t)p * n_pts + j;
            double cx[8], cy[8];
            for (int k = 0; k < n_circ; ++k) { /* :87-89, two roundings each */
                cx[k] = px[q] + off[k] * pcos[q];
                cy[k] = py[q] + off[k] * psin[q];
            }
            for (int o = 0; o < M; ++o) {
                const double ox = obs[2 * o], oy = obs[2 * o + 1];
                int hit = 0;
                for (int k = 0; k < n_circ; ++k) {
                    const double dx = ox - cx[k], dy = oy - cy[k];
                    const double d = sqrt(dx * dx + dy * dy) - rad[k]; /* cdist then subtract, :101-105 */
                    if (d < 0) hit = 1;                                 /* :106-107 */
                    if (d < clr) clr = d;
                }
                tests += n_circ;
                if (hit) {
                    is_free = 0;
                    if (early_exit) break;
                }
            }
        }
        free_out[p] = (unsigned char)is_free;
        if (min_clear) min_clear[p] = clr;
    }
    return tests;
}

static inline double norm2(double v0, double v1, int mode)
{
    switch (mode) {
    case 1: return sqrt(fma(v1, v1, v0 * v0));
    case 2: return sqrt(fma(v0, v0, v1 * v1));
    default: return sqrt(v0 * v0 + v1 * v1);
    }
}

void oracle_norm2(int n, const double *v0, const double *v1, int mode, double *out)
{
    for (int i = 0; i < n; ++i) out[i] = norm2(v0[i], v1[i], mode);
}

/* select_best_path_index on path end points (collision_checker.py:162-203).
 * norm_mode selects the closed form np.linalg.norm([v0, v1]) follows on the host (see
 * oracle/collision_numpy.py).  Returns the index, or -1 for the reference's None.
 * scores_out (may be NULL) receives every path's score (inf for colliding paths). */
int oracle_select_best(int P, const double *ex, const double *ey, const unsigned char *free_in,
                       double gx, double gy, double weight, int norm_mode, double *scores_out, int nthreads)
{
    if (nthreads <= 0) nthreads = 1;
    double *scores = scores_out ? scores_out : (double *)malloc(sizeof(double) * (size_t)(P > 0 ? P : 1));
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int i = 0; i < P; ++i) {
        double score;
        if (free_in[i]) {
            score = norm2(ex[i] - gx, ey[i] - gy, norm_mode);           /* :175 */
            for (int j = 0; j < P; ++j)                                 /* :181-190, ascending j */
                if (j != i && !free_in[j])
                    score += weight * norm2(ex[i] - ex[j], ey[i] - ey[j], norm_mode);
        } else
            score = INFINITY;                                           /* :196 */
        scores[i] = score;
    }
    int best = -1;
    double best_score = INFINITY;
    for (int i = 0; i < P; ++i)
        if (scores[i] < best_score) { /* strict <, first minimum wins, :199-201 */
            best_score = scores[i];
            best = i;
        }
    if (!scores_out) free(scores);
    return best;
}

/* ------------------------------------------------------------------------------------------------
 * Closed-loop tracking: the controllers around the RK4 step as Car.drive runs
 * them (drive.py:126-151) -- StanleyController.get_lookahead_index / stanley_control
 * (stanley_controller.py:56-129), LongitudinalController.long_control (:138-159), the first-order
 * steering filter (drive.py:137-138) -- restated in the reference's operator order.
 * gains = [k, k_soft, max_steer, kp, ki, kd, lookahead, deadband, alpha]; norm_mode as oracle_norm2. */
static double py_mod(double a, double b) /* numpy float remainder for b > 0 */
{
    double m = fmod(a, b);
    if (m != 0.0) {
        if (m < 0.0) m += b;
    } else
        m = copysign(0.0, b);
    return m;
}

static double np_sign(double v) { return v > 0 ? 1.0 : (v < 0 ? -1.0 : (v == 0 ? 0.0 : v)); }

/* stanley_controller.py:56-76 */
int oracle_lookahead_index(const double *wp, int W, double x, double y, double lookahead, int mode, double *min_dist_out)
{
    int min_idx = 0;
    double min_dist = INFINITY;
    for (int i = 0; i < W; ++i) {
        const double d = norm2(wp[2 * i] - x, wp[2 * i + 1] - y, mode);
        if (d < min_dist) {
            min_dist = d;
            min_idx = i;
        }
    }
    double total = min_dist;
    int la = min_idx;
    for (int i = min_idx + 1; i < W; ++i) {
        if (total >= lookahead) break;
        total += norm2(wp[2 * i] - wp[2 * (i - 1)], wp[2 * i + 1] - wp[2 * (i - 1) + 1], mode);
        la = i;
    }
    if (min_dist_out) *min_dist_out = min_dist;
    return la;
}

/* stanley_controller.py:78-129; returns the limited steering angle */
double oracle_stanley_control(const double *wp, int W, double x, double y, double yaw, double v, const double *gains,
                              int mode, int *target_idx, double *crosstrack)
{
    const double k = gains[0], ksoft = gains[1], max_steer = gains[2], L = gains[6], deadband = gains[7];
    const int ce = oracle_lookahead_index(wp, W, x, y, L, mode, NULL);
    const double cv0 = wp[2 * ce] - x - L * cos(yaw);
    const double cv1 = wp[2 * ce + 1] - y - L * sin(yaw);
    double cte = norm2(cv0, cv1, mode);
    if (cte < deadband) cte = 0;
    const double ch = atan2(cv1, cv0);
    double che = ch - yaw;
    che = py_mod(che + M_PI, 2 * M_PI) - M_PI;
    const double sgn = np_sign(che);
    double th;
    if (ce < W - 1)
        th = atan2(wp[2 * (ce + 1) + 1] - wp[2 * ce + 1], wp[2 * (ce + 1)] - wp[2 * ce]);
    else
        th = atan2(wp[1] - wp[2 * (W - 1) + 1], wp[0] - wp[2 * (W - 1)]);
    double he = th - yaw;
    he = py_mod(he + M_PI, 2 * M_PI) - M_PI;
    const double steer = he + atan(k * sgn * cte / (v + ksoft));
    double lim = steer < -max_steer ? -max_steer : steer;   /* np.clip = minimum(maximum(x, lo), hi) */
    lim = lim > max_steer ? max_steer : lim;
    if (target_idx) *target_idx = ce;
    if (crosstrack) *crosstrack = cte;
    return lim;
}

/* Closed loop for V vehicles.  Layouts (vehicle index fastest):
 *   state0 [12][V]; ctrl0 [3][V] = steering-filter state x_del, integral of the speed error, previous speed
 *   wp [n_sets][Wmax][2], wp_count [n_sets]; vehicle r tracks set r / vehicles_per_set
 *   traj [n_out][10][V], log [n_out][45][V] (the DataLog row of drive.py:145-151), target_idx [n_ctrl][V] (each may be NULL)
 *   state_end [12][V], ctrl_end [3][V]
 * step0 (a multiple of ctrl_every) only offsets the logged time column. */
long oracle_track_loop(int V, int n_steps, int step0, double dt, int ctrl_every, const double *state0, const double *ctrl0,
                       const double *wp, const int *wp_count, int Wmax, int vehicles_per_set, const oracle_params *p,
                       double target_vel, const double *gains, int norm_mode, int store_stride, double *traj,
                       double *log, int *target_idx, double *state_end, double *ctrl_end, int nthreads)
{
    if (nthreads <= 0) nthreads = 1;
    const double kp = gains[3], ki = gains[4], kd = gains[5], alpha = gains[8];
    const double mu1[4] = {1.0, 1.0, 1.0, 1.0};   /* drive.py:142 */
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int r = 0; r < V; ++r) {
        double y[10], axay[2], sd[10], out[18];
        for (int c = 0; c < 10; ++c) y[c] = state0[(size_t)c * V + r];
        axay[0] = state0[(size_t)10 * V + r];
        axay[1] = state0[(size_t)11 * V + r];
        double x_del = ctrl0[r], e_int = ctrl0[(size_t)V + r], prev_v = ctrl0[(size_t)2 * V + r];
        const int set = r / vehicles_per_set;
        const double *w = wp + (size_t)set * Wmax * 2;
        const int W = wp_count[set];
        double delta = 0, tau = 0, cte = 0;
        for (int n = 0; n < n_steps; ++n) {
            if (n % ctrl_every == 0) {
                const double v = y[0];
                int ce = 0;
                const double raw = oracle_stanley_control(w, W, y[8], y[9], y[7], v, gains, norm_mode, &ce, &cte);
                /* long_control, :138-159 */
                const double vel_error = target_vel - v;
                e_int = e_int + vel_error * dt;
                const double pp = kp * vel_error, ii = ki * e_int, dd = kd * (v - prev_v) / dt;
                tau = pp + ii + dd;
                if (v <= 0.01) tau = fabs(tau);
                prev_v = v;
                x_del = (1 - alpha) * x_del + alpha * raw;   /* drive.py:137-138 */
                delta = x_del;
                if (target_idx) target_idx[(size_t)(n / ctrl_every) * V + r] = ce;
            }
            const double dl[4] = {delta, delta, 0, 0}, tq[4] = {tau, tau, tau, tau};
            oracle_rk4_step(y, tq, mu1, dl, p, dt, axay, sd, out);
            if (store_stride > 0 && (n + 1) % store_stride == 0) {
                const size_t o = (size_t)((n + 1) / store_stride - 1);
                if (traj) for (int c = 0; c < 10; ++c) traj[(o * 10 + c) * V + r] = y[c];
                if (log) {
                    double *row = log + o * 45 * V + r;
                    row[0] = (double)(step0 + n) * dt;
                    for (int c = 0; c < 10; ++c) row[(size_t)(1 + c) * V] = y[c];
                    for (int c = 0; c < 10; ++c) row[(size_t)(11 + c) * V] = sd[c];
                    row[(size_t)21 * V] = delta;
                    for (int c = 0; c < 4; ++c) row[(size_t)(22 + c) * V] = tau;
                    for (int c = 0; c < 18; ++c) row[(size_t)(26 + c) * V] = out[c];
                    row[(size_t)44 * V] = cte;
                }
            }
        }
        for (int c = 0; c < 10; ++c) state_end[(size_t)c * V + r] = y[c];
        state_end[(size_t)10 * V + r] = axay[0];
        state_end[(size_t)11 * V + r] = axay[1];
        ctrl_end[r] = x_del;
        ctrl_end[(size_t)V + r] = e_int;
        ctrl_end[(size_t)2 * V + r] = prev_v;
    }
    return (long)V * n_steps;
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
