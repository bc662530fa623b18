// K2 / K3: batched circle-offset collision test and best-path selection (sm_100a, FP64, bit-exact).
//
// K2 replaces Pool.starmap(CollisionChecker.collision_check, ...) (reference local_planner.py:369-372,
// collision_checker.py:32-117): every path point carries n_circ circles; a path is free unless some
// obstacle point lies strictly inside some circle.  Bit-exactness rules:
//   * circle centre = fl(x + fl(off * cos(yaw)))            two roundings, no FMA (collision_checker.py:88-89)
//   * scipy cdist distance = sqrt(fl(fl(dx*dx) + fl(dy*dy)))  no FMA
//   * collision iff d - r < 0  <=>  sqrt_rn(q) < r  <=>  q < T(r), T(r) = the smallest double whose
//     correctly rounded square root is >= r (sqrt_rn is monotone).  T is found on the host, which
//     removes the square root from the inner loop without changing a single boolean.
// All products and sums below use the __d*_rn intrinsics, which nvcc never contracts into FMAs.
//
// Mapping: one thread per path point (all its circles in registers), obstacle points staged through
// shared memory in tiles and read as broadcast 16-byte loads; the per-path verdict is a byte that
// any thread of the path clears.  A path that is already known to collide skips the remaining tiles,
// which is the batched form of the reference's early exit (:109-113) and cannot change the result.
//
// K3 follows select_best_path_index (collision_checker.py:134-203) on the path end points: thread i
// accumulates its score over colliding j in ascending order with separately rounded multiply/add, the
// 2-vector norm in the closed form the host BLAS uses; then a lowest-index argmin.
#include <math.h>

#include "b200mp_internal.h"

namespace b200mp {

constexpr int kMaxCircles = 8;
constexpr int kObsTile = 1024;
constexpr int kColBlock = 128;

struct CircleSpec {
    double off[kMaxCircles];
    double rad[kMaxCircles];
    double thr[kMaxCircles];   // q < thr  <=>  sqrt_rn(q) < rad
};

// smallest double T with sqrt_rn(T) >= r  (host; sqrt is correctly rounded)
static double sqrt_threshold(double r)
{
    if (!(r > 0.0)) return 0.0;   // d >= 0 is never < r <= 0
    double q = r * r;
    while (sqrt(q) >= r) q = nextafter(q, 0.0);
    while (sqrt(q) < r) q = nextafter(q, INFINITY);
    return q;
}

template <int NC, bool CLEAR>
__global__ void __launch_bounds__(kColBlock)
collision_kernel(int n_items, int n_pts, const __grid_constant__ CircleSpec cs, const double *__restrict__ px,
                 const double *__restrict__ py, const double *__restrict__ pcos, const double *__restrict__ psin,
                 const double *__restrict__ pyaw, int yaw_stride, int M, const double2 *__restrict__ obs,
                 unsigned char *free_out, double *__restrict__ clear_pts)
{
    __shared__ double2 tile[kObsTile];
    const int t = blockIdx.x * kColBlock + threadIdx.x;
    bool active = t < n_items;
    const int p = active ? t / n_pts : 0;
    double cx[NC], cy[NC];
    if (active) {
        const double x = px[t], y = py[t];
        double c, s;
        if (pcos) {
            c = pcos[t];
            s = psin[t];
        } else {
            sincos(pyaw[(size_t)p * yaw_stride + (t - p * n_pts)], &s, &c);
        }
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            cx[k] = __dadd_rn(x, __dmul_rn(cs.off[k], c));
            cy[k] = __dadd_rn(y, __dmul_rn(cs.off[k], s));
        }
    }
    double clr = INFINITY;
    for (int m0 = 0; m0 < M; m0 += kObsTile) {
        const int m1 = min(kObsTile, M - m0);
        __syncthreads();
        for (int i = threadIdx.x; i < m1; i += kColBlock) tile[i] = obs[m0 + i];
        if (!CLEAR && active && ((volatile unsigned char *)free_out)[p] == 0) active = false;   // path already known to collide
        if (!__syncthreads_or(active)) break;
        if (!active) continue;
        bool hit = false;
#pragma unroll 4
        for (int o = 0; o < m1; ++o) {
            const double2 ob = tile[o];
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                const double dx = __dsub_rn(ob.x, cx[k]);
                const double dy = __dsub_rn(ob.y, cy[k]);
                const double q = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                if (CLEAR) {
                    const double d = __dsub_rn(__dsqrt_rn(q), cs.rad[k]);   // collision_checker.py:101-105
                    hit |= d < 0.0;
                    clr = fmin(clr, d);
                } else {
                    hit |= q < cs.thr[k];
                }
            }
        }
        if (hit) free_out[p] = 0;
    }
    if (CLEAR && t < n_items) clear_pts[t] = clr;
}

__global__ void clearance_reduce_kernel(int P, int n_pts, const double *__restrict__ clear_pts, double *__restrict__ out)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    double m = INFINITY;
    for (int j = 0; j < n_pts; ++j) m = fmin(m, clear_pts[(size_t)p * n_pts + j]);
    out[p] = m;
}

template <int NC>
static int launch_collision_nc(int device, cudaStream_t st, int P, int n_pts, const CircleSpec &cs, const double *px,
                               const double *py, const double *pcos, const double *psin, const double *pyaw,
                               int yaw_stride, int M, const double *obs, unsigned char *free_out, double *min_clear)
{
    const long long items = (long long)P * n_pts;
    const int grid = (int)((items + kColBlock - 1) / kColBlock);
    if (min_clear) {
        void *scratch = nullptr;
        int rc = ensure_scratch(device, sizeof(double) * (size_t)items, &scratch);
        if (rc) return rc;
        collision_kernel<NC, true><<<grid, kColBlock, 0, st>>>((int)items, n_pts, cs, px, py, pcos, psin, pyaw, yaw_stride,
                                                              M, (const double2 *)obs, free_out, (double *)scratch);
        B200MP_CUDA(cudaGetLastError());
        clearance_reduce_kernel<<<(P + 127) / 128, 128, 0, st>>>(P, n_pts, (const double *)scratch, min_clear);
    } else {
        collision_kernel<NC, false><<<grid, kColBlock, 0, st>>>((int)items, n_pts, cs, px, py, pcos, psin, pyaw, yaw_stride,
                                                               M, (const double2 *)obs, free_out, nullptr);
    }
    B200MP_CUDA(cudaGetLastError());
    return 0;
}

int launch_collision_f64(int device, cudaStream_t st, int P, int n_pts, int n_circ, const double *off,
                         const double *rad, const double *px, const double *py, const double *pcos,
                         const double *psin, const double *pyaw, int yaw_stride, int M, const double *obs,
                         unsigned char *free_out, double *min_clear)
{
    if (P < 0 || n_pts < 0 || M < 0 || n_circ < 1 || n_circ > kMaxCircles) {
        set_error("collision_check: bad sizes P=%d n_pts=%d n_circ=%d (1..%d) M=%d", P, n_pts, n_circ, kMaxCircles, M);
        return B200MP_E_ARG;
    }
    if (P == 0) return 0;
    if (!off || !rad || !free_out) {
        set_error("collision_check: off, rad and free_out must be non-NULL");
        return B200MP_E_ARG;
    }
    // every path starts collision-free (empty path or empty obstacle list -> True, as the reference)
    B200MP_CUDA(cudaMemsetAsync(free_out, 1, (size_t)P, st));
    if (n_pts == 0 || M == 0) {
        if (min_clear) {
            // +inf clearance: no test was made
            clearance_reduce_kernel<<<(P + 127) / 128, 128, 0, st>>>(P, 0, nullptr, min_clear);
            B200MP_CUDA(cudaGetLastError());
        }
        return 0;
    }
    if (!px || !py || !obs || ((!pcos || !psin) && !pyaw)) {
        set_error("collision_check: px, py, obs and (pcos, psin) or pyaw must be non-NULL");
        return B200MP_E_ARG;
    }
    if ((long long)P * n_pts > 0x7fffffffLL) {
        set_error("collision_check: P*n_pts exceeds 2^31-1");
        return B200MP_E_ARG;
    }
    if ((((size_t)obs) & 15) != 0) {
        set_error("collision_check: obs must be 16-byte aligned");
        return B200MP_E_ARG;
    }
    CircleSpec cs{};
    for (int k = 0; k < n_circ; ++k) {
        cs.off[k] = off[k];
        cs.rad[k] = rad[k];
        cs.thr[k] = sqrt_threshold(rad[k]);
    }
    if (!pcos || !psin) pcos = psin = nullptr;
    switch (n_circ) {
#define B200MP_NC(N) \
    case N: return launch_collision_nc<N>(device, st, P, n_pts, cs, px, py, pcos, psin, pyaw, yaw_stride, M, obs, free_out, min_clear);
        B200MP_NC(1) B200MP_NC(2) B200MP_NC(3) B200MP_NC(4) B200MP_NC(5) B200MP_NC(6) B200MP_NC(7) B200MP_NC(8)
#undef B200MP_NC
    }
    return B200MP_E_ARG;
}

// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double norm2_host_form(double v0, double v1, int mode)
{
    double q;
    if (mode == B200MP_NORM2_FMA_V1)
        q = __fma_rn(v1, v1, __dmul_rn(v0, v0));
    else if (mode == B200MP_NORM2_FMA_V0)
        q = __fma_rn(v0, v0, __dmul_rn(v1, v1));
    else
        q = __dadd_rn(__dmul_rn(v0, v0), __dmul_rn(v1, v1));
    return __dsqrt_rn(q);
}

__global__ void __launch_bounds__(128)
select_score_kernel(int P, const double *__restrict__ ex, const double *__restrict__ ey,
                    const unsigned char *__restrict__ free_in, double gx, double gy, double weight, int mode,
                    double *__restrict__ scores)
{
    __shared__ double sx[128], sy[128];
    __shared__ unsigned char sf[128];
    const int i = blockIdx.x * 128 + threadIdx.x;
    const bool valid = i < P;
    const double xi = valid ? ex[i] : 0.0, yi = valid ? ey[i] : 0.0;
    const bool fi = valid && free_in[i] != 0;
    double score = norm2_host_form(__dsub_rn(xi, gx), __dsub_rn(yi, gy), mode);   // collision_checker.py:175
    for (int j0 = 0; j0 < P; j0 += 128) {
        __syncthreads();
        const int j = j0 + threadIdx.x;
        if (j < P) {
            sx[threadIdx.x] = ex[j];
            sy[threadIdx.x] = ey[j];
            sf[threadIdx.x] = free_in[j];
        }
        __syncthreads();
        const int jn = min(128, P - j0);
        for (int jj = 0; jj < jn; ++jj) {               // ascending j, sequential adds (:181-190)
            if (sf[jj] == 0) {                          // block-uniform branch; j == i is never colliding when i is free
                const double n = norm2_host_form(__dsub_rn(xi, sx[jj]), __dsub_rn(yi, sy[jj]), mode);
                score = __dadd_rn(score, __dmul_rn(weight, n));
            }
        }
    }
    if (valid) scores[i] = fi ? score : INFINITY;       // :196
}

int launch_select_best_f64(int device, cudaStream_t st, int P, const double *ex, const double *ey,
                           const unsigned char *free_in, double gx, double gy, double weight, int norm_mode,
                           double *scores_out, int *best_out)
{
    if (P < 0 || !best_out || (P > 0 && (!ex || !ey || !free_in)) || norm_mode < 0 || norm_mode > 2) {
        set_error("select_best: bad arguments (P=%d, norm_mode=%d)", P, norm_mode);
        return B200MP_E_ARG;
    }
    // scratch: [P] scores (when the caller does not want them) followed by the argmin partials
    void *scratch = nullptr;
    const size_t score_bytes = (sizeof(double) * (size_t)P + 255) & ~(size_t)255;
    int rc = ensure_scratch(device, score_bytes + argmin_scratch_bytes(P), &scratch);
    if (rc) return rc;
    double *scores = scores_out ? scores_out : (double *)scratch;
    if (P > 0) {
        select_score_kernel<<<(P + 127) / 128, 128, 0, st>>>(P, ex, ey, free_in, gx, gy, weight, norm_mode, scores);
        B200MP_CUDA(cudaGetLastError());
    }
    return argmin_launch(st, P, scores, 0, (char *)scratch + score_bytes, nullptr, nullptr, best_out);
}

}  // namespace b200mp
