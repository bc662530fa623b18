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
// K3 follows select_best_path_index (collision_checker.py:134-203) on the path end points: one warp per path i
// accumulates its score over colliding j in ascending order with separately rounded multiply/add (the norms of 32 j
// at a time in parallel, folded in lane order), the 2-vector norm in the closed form the host BLAS uses; then a
// lowest-index argmin.
#include <math.h>
#include <stdlib.h>

#include "b200mp_internal.h"

namespace b200mp {

#ifndef B200MP_CULL_UNROLL
#define B200MP_CULL_UNROLL 16
#endif
constexpr int kCullUnroll = B200MP_CULL_UNROLL;
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

// ---------------------------------------------------------------------------------------------------
// Exact verdicts from yaws (no host trigonometry on the bulk of the data).  The reference evaluates cos / sin of
// every path yaw with numpy (collision_checker.py:88-89); an obstacle point is inside a circle iff
// sqrt_rn(q64) < r with q64 computed from centres that carry numpy's roundings.  The kernels below evaluate the yaws
// with the CUDA library sincos instead and PROVE each verdict against any host trigonometry within kTrigTau of the
// exact value (numpy measures <= 0.52 ulp here, CUDA documents <= 2 ulp; tau = 2^-44 is ~500 ulp):
//   |c_host - c_dev| <= tau  =>  host and device centres differ by at most
//       ec = 2 |off| (tau + 2^-50) + 2^-51 (|cx| + |cy|)            (two products, two sums, both roundings each)
//   and with d = true distance, sqrt_rn(q64) = d (1 +- 2^-51) on either side, so
//       q64_dev <  LO64 = ((r (1 - 2^-50) - ec) / (1 + 2^-51))^2 (rounded down)  =>  the host's test collides
//       q64_dev >= HI64 = ((r (1 + 2^-50) + ec) / (1 - 2^-51))^2 (rounded up)    =>  the host's test is free.
// The FP32 screen's band (1e-5 m wide) is widened by ec (1e-13 m); a pair that lands in [LO64, HI64) -- an obstacle
// point within ~1e-13 m of a circle -- is NOT decided on the device: the path point's index is appended to a list
// the caller resolves with host-evaluated cos / sin (collision_resolve_kernel; a handful of points per batch at most).
constexpr double kTrigTau = 5.6843418860808015e-14;   // 2^-44

struct YawFix {
    int *list;      // list[0] = number of appended items (may exceed capacity), list[1..capacity] = item = p * n_pts + j
    int capacity;   // list == nullptr: legacy device-trig mode (pairs inside the band are decided with the device's values)
};

__device__ __forceinline__ void yaw_fix_append(const YawFix &yf, int item)
{
    const int i = atomicAdd(yf.list, 1);
    if (i < yf.capacity) yf.list[1 + i] = item;
}

// per-thread centre uncertainty and the certain-hit / certain-free thresholds on q64
template <int NC>
__device__ __forceinline__ double yaw_bounds(const CircleSpec &cs, const double *cx, const double *cy, double *lo64, double *hi64)
{
    double ec = 0.0;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const double e = 2.0 * fabs(cs.off[k]) * (kTrigTau + 8.881784197001252e-16) + 4.440892098500626e-16 * (fabs(cx[k]) + fabs(cy[k]));
        ec = fmax(ec, e);
        if (!(e == e)) ec = INFINITY;    // fmax drops NaN: a NaN centre never hits in either arithmetic, inf keeps it undecidable
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const double r = cs.rad[k];
        const double a = (r * (1.0 - 8.881784197001252e-16) - ec) / (1.0 + 4.440892098500626e-16);
        const double b = (r * (1.0 + 8.881784197001252e-16) + ec) / (1.0 - 4.440892098500626e-16);
        const bool ok = r > 1.0e-100 && r < 1.0e150 && ec < INFINITY;
        lo64[k] = (ok && a > 0.0) ? a * a * (1.0 - 2.0e-15) : 0.0;
        hi64[k] = ok ? b * b * (1.0 + 2.0e-15) : INFINITY;
        if (!(r > 0.0)) lo64[k] = hi64[k] = 0.0;   // r <= 0 or NaN: d >= 0 is never < r
    }
    return ec;
}

template <int NC, bool CLEAR, bool YAWFIX = false>
__global__ void __launch_bounds__(kColBlock)
collision_kernel(int n_items, int n_pts, const __grid_constant__ CircleSpec cs, const double *__restrict__ px,
                 const double *__restrict__ py, const double *__restrict__ pcos, const double *__restrict__ psin,
                 const double *__restrict__ pyaw, int yaw_stride, int M, const double2 *__restrict__ obs,
                 unsigned char *free_out, double *__restrict__ clear_pts, YawFix yf = YawFix{nullptr, 0})
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
    // CLEAR: min over obstacle points of fl(sqrt_rn(q) - r) equals fl(sqrt_rn(min q) - r) bit for bit (both roundings
    // are monotone), so the running minimum is kept on q and ONE square root per circle is taken at the end
    double qmin[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) qmin[k] = INFINITY;
    // YAWFIX: verdicts proven against host trigonometry (see yaw_bounds); a pair inside the band is left to the caller
    double lo64[YAWFIX ? NC : 1], hi64[YAWFIX ? NC : 1];
    bool appended = false;
    if (YAWFIX && active) yaw_bounds<NC>(cs, cx, cy, lo64, hi64);
    for (int m0 = 0; m0 < M; m0 += kObsTile) {
        const int m1 = min(kObsTile, M - m0);
        __syncthreads();
        for (int i = threadIdx.x; i < m1; i += kColBlock) tile[i] = obs[m0 + i];
        if (!CLEAR && active && ((volatile unsigned char *)free_out)[p] == 0) active = false;   // path already known to collide
        if (!__syncthreads_or(active)) break;
        if (!active) continue;
        bool hit = false, maybe = false;
#pragma unroll 4
        for (int o = 0; o < m1; ++o) {
            const double2 ob = tile[o];
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                const double dx = __dsub_rn(ob.x, cx[k]);
                const double dy = __dsub_rn(ob.y, cy[k]);
                const double q = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                if (CLEAR) {
                    qmin[k] = fmin(qmin[k], q);                             // NaN is skipped, as fmin(clr, d) did
                } else if (YAWFIX) {
                    hit |= q < lo64[k];
                    maybe |= q < hi64[k];
                } else {
                    hit |= q < cs.thr[k];
                }
            }
        }
        if (CLEAR) {
#pragma unroll
            for (int k = 0; k < NC; ++k) hit |= qmin[k] < cs.thr[k];        // sqrt_rn(q) - r < 0  <=>  q < thr
        }
        if (hit) {
            free_out[p] = 0;
        } else if (YAWFIX && maybe && !appended) {
            yaw_fix_append(yf, t);
            appended = true;
        }
    }
    if (CLEAR && t < n_items) {
        double clr = INFINITY;
#pragma unroll
        for (int k = 0; k < NC; ++k) clr = fmin(clr, __dsub_rn(__dsqrt_rn(qmin[k]), cs.rad[k]));   // collision_checker.py:101-105
        clear_pts[t] = clr;
    }
}

// ---------------------------------------------------------------------------------------------------
// K2 fast path: FP32 filter + exact FP64 recheck.  The verdict of a path is a boolean, so the FP64
// arithmetic above is needed only for tests whose outcome single precision cannot decide.  Every
// (circle, obstacle point) pair is first evaluated in FP32 on coordinates shifted to a common origin
// (obs[0]); with
//   A   = max |o - origin| over the obstacle tile (as stored in FP32), Ac = max |c - origin| of the thread,
//   E0  = 2 u (A + Ac) (1 + 1e-6) + 1e-18,  u = 2^-24     (conversion errors of both points, both axes)
//   d^  = sqrt(q32) (1 +- 4u)                              (FP32 subtract / multiply / FMA roundings)
// the FP64 distance the exact test uses satisfies |d - d^| <= E0, hence
//   q32 <  LO = ((r (1 - 2^-40) - E0) / (1 + 4u))^2  (rounded down)  =>  q64 <  T(r): certain collision
//   q32 >= HI = ((r (1 + 2^-40) + E0) / (1 - 4u))^2  (rounded up)    =>  q64 >= T(r): certainly free
// and only a thread whose per-circle minimum lands in [LO, HI) repeats its obstacle tile with the exact
// FP64 sequence (a band a few 1e-5 m wide around each circle: ~1e-3 of the threads per tile on
// BASELINE config 3).  Non-finite or huge coordinates make E0 non-finite/large, which sends the thread
// to the exact path, so the result is bit-identical to collision_kernel for every input.
// One CTA = 128 path points x one obstacle tile (grid.y = tiles): 5x more CTAs than tiles-in-a-loop,
// which evens out the per-SM load once early-exiting paths drop out.
constexpr int kFTile = 2048;

#ifndef B200MP_COLLISION_PACKED
#define B200MP_COLLISION_PACKED 1
#endif
__device__ __forceinline__ unsigned long long pack2(float a, float b)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float lo2(unsigned long long v)
{
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a;
}
__device__ __forceinline__ float hi2(unsigned long long v)
{
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return b;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// exact FP64 sequence over obstacle points [m_begin, m_end): bit 0 = some q < lo (a certain hit), bit 1 = some q < hi
// (lo = hi = thr when the centres carry the caller's own cos / sin; the yaw_bounds pair otherwise)
template <int NC>
__device__ __noinline__ int exact_tile_bits(const double *lo, const double *hi, const double *cx, const double *cy,
                                            const double2 *__restrict__ obs, int m_begin, int m_end)
{
    bool hit = false, maybe = false;
    for (int o = m_begin; o < m_end; ++o) {
        const double2 ob = obs[o];
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const double dx = __dsub_rn(ob.x, cx[k]);
            const double dy = __dsub_rn(ob.y, cy[k]);
            const double q = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            hit |= q < lo[k];
            maybe |= q < hi[k];
        }
    }
    return (hit ? 1 : 0) | (maybe ? 2 : 0);
}

template <int NC>
__global__ void __launch_bounds__(kColBlock)
collision_filter_kernel(int n_items, int n_pts, const __grid_constant__ CircleSpec cs, const double *__restrict__ px,
                        const double *__restrict__ py, const double *__restrict__ pcos, const double *__restrict__ psin,
                        const double *__restrict__ pyaw, int yaw_stride, int M, const double2 *__restrict__ obs,
                        unsigned char *free_out, YawFix yf)
{
    __shared__ __align__(16) float2 tile[kFTile];
    __shared__ int s_amax;
    const int t = blockIdx.x * kColBlock + threadIdx.x;
    bool active = t < n_items;
    const int p = active ? t / n_pts : 0;
    const int m0 = blockIdx.y * kFTile;
    const int m1 = min(kFTile, M - m0);
    const int m1pad = (m1 + 7) & ~7;
    const double2 org = obs[0];
    if (threadIdx.x == 0) s_amax = 0;
    if (active && ((volatile unsigned char *)free_out)[p] == 0) active = false;   // path already known to collide
    if (!__syncthreads_or(active)) return;

    // obstacle tile -> FP32 relative to the origin; padding points are far away from everything
    float amax = 0.0f;
    for (int i = threadIdx.x; i < m1pad; i += kColBlock) {
        float2 v = make_float2(INFINITY, INFINITY);   // padding: q32 = +inf, never below a threshold
        if (i < m1) {
            const double2 o = obs[m0 + i];
            v.x = (float)(o.x - org.x);
            v.y = (float)(o.y - org.y);
            amax = fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y)));   // a NaN point never hits in either precision
        }
#if B200MP_COLLISION_PACKED
        float *tf = reinterpret_cast<float *>(tile);
        tf[(i >> 1) * 4 + (i & 1)] = v.x;       // pair layout (xa, xb, ya, yb)
        tf[(i >> 1) * 4 + 2 + (i & 1)] = v.y;
#else
        tile[i] = v;
#endif
    }
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, w));
    if ((threadIdx.x & 31) == 0) atomicMax(&s_amax, __float_as_int(amax));   // non-negative floats order as ints
    __syncthreads();
    if (!active) return;

    double cx[NC], cy[NC];
    float fx[NC], fy[NC], lo[NC], hi[NC];
    {
        const double x = px[t], y = py[t];
        double c, s;
        if (pcos) {
            c = pcos[t];
            s = psin[t];
        } else {
            sincos(pyaw[(size_t)p * yaw_stride + (t - p * n_pts)], &s, &c);
        }
        double ac = 0.0;
        bool bad = false;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            cx[k] = __dadd_rn(x, __dmul_rn(cs.off[k], c));
            cy[k] = __dadd_rn(y, __dmul_rn(cs.off[k], s));
            const double rx = cx[k] - org.x, ry = cy[k] - org.y;
            fx[k] = (float)rx;
            fy[k] = (float)ry;
            const double m = fmax(fabs(rx), fabs(ry));
            bad |= !(fabs(rx) <= 1.0e30) | !(fabs(ry) <= 1.0e30);   // NaN / Inf / beyond FP32 range (fmax drops NaNs)
            ac = fmax(ac, m);
        }
        const double u = 5.9604644775390625e-08;   // 2^-24
        double e0 = 2.0 * u * ((double)__int_as_float(s_amax) + ac) * 1.000001 + 1.0e-18;
        if (yf.list) {   // centres from the device's sincos: the band also covers any host trigonometry (yaw_bounds)
            double l64[NC], h64[NC];
            e0 += yaw_bounds<NC>(cs, cx, cy, l64, h64);
        }
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const double r = cs.rad[k];
            // the screen is used only where it is sharp (error far below the radius) and nothing can overflow;
            // otherwise hi = NaN makes every outcome "undecided" and the tile is evaluated exactly
            const bool sharp = !bad && e0 < 0.25 * r && r < 1.0e15;
            const double a = (r * (1.0 - 9.094947017729282e-13) - e0) / (1.0 + 4.0 * u);
            const double b = (r * (1.0 + 9.094947017729282e-13) + e0) / (1.0 - 4.0 * u);
            lo[k] = sharp ? __double2float_rd(a * a * (1.0 - 1.0e-15)) : 0.0f;
            hi[k] = sharp ? __double2float_ru(b * b * (1.0 + 1.0e-15)) : __int_as_float(0x7fc00000);
            if (!(r > 0.0)) hi[k] = 0.0f;   // r <= 0 or NaN: d >= 0 is never < r, nothing to decide
        }
    }
    float mn[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) mn[k] = INFINITY;
#if B200MP_COLLISION_PACKED
    // two obstacle points per instruction (Blackwell packed FP32: add/mul/fma.f32x2); the tile stores
    // obstacle pairs as (xa, xb, ya, yb)
    unsigned long long ncx[NC], ncy[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        ncx[k] = pack2(-fx[k], -fx[k]);
        ncy[k] = pack2(-fy[k], -fy[k]);
    }
    const float4 *tile4 = reinterpret_cast<const float4 *>(tile);
#pragma unroll 2
    for (int o = 0; o < m1pad / 2; ++o) {
        const float4 ob = tile4[o];
        const unsigned long long X = pack2(ob.x, ob.y), Y = pack2(ob.z, ob.w);
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const unsigned long long dx = add2(X, ncx[k]), dy = add2(Y, ncy[k]);
            const unsigned long long q = fma2(dy, dy, mul2(dx, dx));
            mn[k] = fminf(mn[k], fminf(lo2(q), hi2(q)));
        }
    }
#else
#pragma unroll 4
    for (int o = 0; o < m1pad; ++o) {
        const float2 ob = tile[o];
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const float dx = ob.x - fx[k];
            const float dy = ob.y - fy[k];
            mn[k] = fminf(mn[k], fmaf(dy, dy, dx * dx));
        }
    }
#endif
    bool hit = false, undecided = false;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        hit |= mn[k] < lo[k];
        undecided |= !(mn[k] >= hi[k]);
    }
    if (!hit && undecided) {
        int bits;
        if (yf.list) {
            double l64[NC], h64[NC];
            yaw_bounds<NC>(cs, cx, cy, l64, h64);
            bits = exact_tile_bits<NC>(l64, h64, cx, cy, obs, m0, m0 + m1);
        } else {
            bits = exact_tile_bits<NC>(cs.thr, cs.thr, cx, cy, obs, m0, m0 + m1);
        }
        hit = bits & 1;
        if (bits == 2) yaw_fix_append(yf, t);   // within ~1e-13 m of a circle: the caller decides with host cos / sin
    }
    if (hit) free_out[p] = 0;
}

// ---------------------------------------------------------------------------------------------------
// Exact minimum clearance with an FP32 screen (CLEAR mode of K2 in the AUTO / SCREEN arithmetic).  The result is
// min over obstacle points of fl(sqrt_rn(q64) - r) = fl(sqrt_rn(min q64) - r) per circle, so only the exact minimum of
// q64 is needed.  Pass A sweeps all obstacle tiles in packed FP32 and keeps min q32 per circle; with the error model of
// the screen above, |d64 - sqrt(q32)(1 +- 4u)| <= E0 for every point, the exact minimiser o* satisfies
//     sqrt(q32(o*)) <= (d64(o*) + E0)/(1 - 4u) <= (d64(o_A) + E0)/(1 - 4u) <= (sqrt(min q32)(1 + 4u) + 2 E0)/(1 - 4u) =: sqrt(T),
// so pass B repeats the FP32 sweep and evaluates q64 exactly only for the points with q32 <= T (the minimiser, its ties
// and a few neighbours within ~1e-5 m): two FP32 sweeps instead of one FP64 sweep.  Where the screen is not sharp
// (non-finite or huge coordinates) T = +inf and every point is evaluated exactly, as in collision_kernel<NC, true>;
// NaN points are skipped by fmin in both, so the doubles returned are identical for every input.
template <int NC>
__global__ void __launch_bounds__(kColBlock)
clearance_screen_kernel(int n_items, int n_pts, const __grid_constant__ CircleSpec cs, const double *__restrict__ px,
                        const double *__restrict__ py, const double *__restrict__ pcos, const double *__restrict__ psin,
                        const double *__restrict__ pyaw, int yaw_stride, int M, const double2 *__restrict__ obs,
                        unsigned char *free_out, double *__restrict__ clear_pts)
{
    __shared__ __align__(16) float2 tile[kFTile];
    __shared__ int s_amax;
    const int t = blockIdx.x * kColBlock + threadIdx.x;
    const bool active = t < n_items;
    const int p = active ? t / n_pts : 0;
    const double2 org = obs[0];
    if (threadIdx.x == 0) s_amax = 0;

    double cx[NC], cy[NC], ac = 0.0;
    unsigned long long ncx[NC], ncy[NC];
    bool bad = false;
    {
        double x = 0.0, y = 0.0, c = 1.0, s = 0.0;
        if (active) {
            x = px[t];
            y = py[t];
            if (pcos) {
                c = pcos[t];
                s = psin[t];
            } else {
                sincos(pyaw[(size_t)p * yaw_stride + (t - p * n_pts)], &s, &c);
            }
        }
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            cx[k] = __dadd_rn(x, __dmul_rn(cs.off[k], c));
            cy[k] = __dadd_rn(y, __dmul_rn(cs.off[k], s));
            const double rx = cx[k] - org.x, ry = cy[k] - org.y;
            const float fx = (float)rx, fy = (float)ry;
            ncx[k] = pack2(-fx, -fx);
            ncy[k] = pack2(-fy, -fy);
            bad |= !(fabs(rx) <= 1.0e30) | !(fabs(ry) <= 1.0e30);   // NaN / Inf / beyond FP32 range
            ac = fmax(ac, fmax(fabs(rx), fabs(ry)));
        }
    }
    const float4 *tile4 = reinterpret_cast<const float4 *>(tile);
    float amax = 0.0f;
    // stages obstacle tile m0 as FP32 pairs (xa, xb, ya, yb) relative to the origin; padding points are at infinity
    auto stage = [&](int m0, int m1, int m1pad) {
        __syncthreads();
        float *tf = reinterpret_cast<float *>(tile);
        for (int i = threadIdx.x; i < m1pad; i += kColBlock) {
            float2 v = make_float2(INFINITY, INFINITY);
            if (i < m1) {
                const double2 o = obs[m0 + i];
                v.x = (float)(o.x - org.x);
                v.y = (float)(o.y - org.y);
                amax = fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y)));
            }
            tf[(i >> 1) * 4 + (i & 1)] = v.x;
            tf[(i >> 1) * 4 + 2 + (i & 1)] = v.y;
        }
        __syncthreads();
    };

    // ---- pass A: min q32 per circle over all obstacle points
    float mn[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) mn[k] = INFINITY;
    for (int m0 = 0; m0 < M; m0 += kFTile) {
        const int m1 = min(kFTile, M - m0), m1pad = (m1 + 7) & ~7;
        stage(m0, m1, m1pad);
        if (!active) continue;
#pragma unroll 2
        for (int o = 0; o < m1pad / 2; ++o) {
            const float4 ob = tile4[o];
            const unsigned long long X = pack2(ob.x, ob.y), Y = pack2(ob.z, ob.w);
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                const unsigned long long dx = add2(X, ncx[k]), dy = add2(Y, ncy[k]);
                const unsigned long long q = fma2(dy, dy, mul2(dx, dx));
                mn[k] = fminf(mn[k], fminf(lo2(q), hi2(q)));
            }
        }
    }
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, w));
    if ((threadIdx.x & 31) == 0) atomicMax(&s_amax, __float_as_int(amax));   // non-negative floats order as ints
    __syncthreads();

    // ---- candidate thresholds
    float T[NC];
    {
        const double u = 5.9604644775390625e-08;   // 2^-24
        const double e0 = 2.0 * u * ((double)__int_as_float(s_amax) + ac) * 1.000001 + 1.0e-18;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const double b = (sqrt((double)mn[k]) * (1.0 + 4.0 * u) + 2.0 * e0) / (1.0 - 4.0 * u);
            const bool sharp = !bad && e0 < 1.0e30 && mn[k] < INFINITY;
            T[k] = sharp ? __double2float_ru(b * b * 1.000001) : INFINITY;
        }
    }

    // ---- pass B: exact q64 for the candidates only
    double qmin[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) qmin[k] = INFINITY;
    for (int m0 = 0; m0 < M; m0 += kFTile) {
        const int m1 = min(kFTile, M - m0), m1pad = (m1 + 7) & ~7;
        stage(m0, m1, m1pad);
        if (!active) continue;
#pragma unroll 2
        for (int o = 0; o < m1pad / 2; ++o) {
            const float4 ob = tile4[o];
            const unsigned long long X = pack2(ob.x, ob.y), Y = pack2(ob.z, ob.w);
            bool cand = false;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                const unsigned long long dx = add2(X, ncx[k]), dy = add2(Y, ncy[k]);
                const unsigned long long q = fma2(dy, dy, mul2(dx, dx));
                cand |= !(fminf(lo2(q), hi2(q)) > T[k]);   // NaN counts as a candidate (it drops out in fmin below)
            }
            if (cand) {
                for (int i = 2 * o; i < 2 * o + 2 && i < m1; ++i) {
                    const double2 e = obs[m0 + i];
#pragma unroll
                    for (int k = 0; k < NC; ++k) {
                        const double dx = __dsub_rn(e.x, cx[k]);
                        const double dy = __dsub_rn(e.y, cy[k]);
                        qmin[k] = fmin(qmin[k], __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                    }
                }
            }
        }
    }
    if (!active) return;
    bool hit = false;
    double clr = INFINITY;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        hit |= qmin[k] < cs.thr[k];                                              // sqrt_rn(q) - r < 0  <=>  q < thr
        clr = fmin(clr, __dsub_rn(__dsqrt_rn(qmin[k]), cs.rad[k]));             // collision_checker.py:101-105
    }
    if (hit) free_out[p] = 0;
    clear_pts[t] = clr;
}

// ---------------------------------------------------------------------------------------------------
// K2 broad phase: the obstacle set of a planner is a handful of outlines (the reference samples boxes every
// 0.2 m, env.py:93-127), i.e. consecutive obstacle points are spatially coherent.  A prepare kernel shifts the
// points to the common origin in FP32 once per launch (instead of once per CTA and tile) and records the bounding
// box of every 32 consecutive points.  In the main kernel one CTA walks ALL obstacle chunks: each lane of a warp
// compares one chunk box with the warp's box (the union over its 32 path points of the circle centres grown by the
// "cannot be a collision beyond this" radius sqrt(HI) of the FP32 screen), a ballot collects the overlapping chunks,
// and only those are screened point by point (packed FP32, as above).  A culled pair has |o - c|_inf > sqrt(HI) in
// the FP32 coordinates, which is exactly a pair the screen declares free, so the flags stay bit-identical to the
// all-FP64 kernel; threads whose screen is not sharp (hi = NaN) make their warp's box infinite (nothing is culled)
// and an undecided thread repeats, in exact FP64, the chunks that overlap ITS box.  On BASELINE config 3 about 2 %
// of the chunks survive.
struct ObsPrep {
    float amax;                    // max |o - origin| over all obstacle points (FP32 coordinates)
    int pad;
    unsigned long long screened;   // (warp, chunk) pairs that survived the broad phase
    unsigned long long rechecked;  // (thread, chunk) pairs repeated in FP64
    double2 org;                   // the common origin of the FP32 coordinates: the caller's obs[0]
};

// Spatial ordering of the obstacle points.  The broad phase culls 32-point chunks by their bounding boxes, so it is only as
// good as consecutive points are close together: a planner's list is a concatenation of outlines (env.py:93-127), and every
// chunk that straddles two outlines has a box spanning both (a third of the chunks of BASELINE config 3 -- they were ~90 %
// of the chunks that survived the cull).  The verdicts and the minimum clearance are functions of the SET of points, so the
// points are first put into Morton order of a 64 x 64 grid over their bounding box (one CTA: bounds, histogram, scan and
// scatter in shared memory; the order inside a cell is whatever the atomics give and does not matter).  NaN / Inf points go
// to the last cell.  Lists beyond kSortMaxPoints keep the caller's order.
constexpr int kSortMaxPoints = 1 << 17;
constexpr int kSortGrid = 64;
constexpr int kSortCells = kSortGrid * kSortGrid;

__device__ __forceinline__ unsigned sort_cell(float x, float y, float x0, float y0, float sx, float sy)
{
    if (!(fabsf(x) < INFINITY) || !(fabsf(y) < INFINITY)) return kSortCells - 1;
    const int cx = min(kSortGrid - 1, max(0, (int)((x - x0) * sx))), cy = min(kSortGrid - 1, max(0, (int)((y - y0) * sy)));
    unsigned m = 0;
#pragma unroll
    for (int b = 0; b < 6; ++b) m |= (((unsigned)cx >> b) & 1u) << (2 * b) | (((unsigned)cy >> b) & 1u) << (2 * b + 1);
    return m;
}

__global__ void __launch_bounds__(1024)
obstacle_sort_kernel(int M, const double2 *__restrict__ obs, double2 *__restrict__ sorted)
{
    __shared__ int cell_pos[kSortCells];
    __shared__ float red[4][32];
    __shared__ int warp_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double2 org = obs[0];
    // bounds of the finite points (FP32, relative to the origin)
    float x0 = INFINITY, x1 = -INFINITY, y0 = INFINITY, y1 = -INFINITY;
    for (int i = tid; i < M; i += 1024) {
        const double2 o = obs[i];
        const float x = (float)(o.x - org.x), y = (float)(o.y - org.y);
        if (fabsf(x) < INFINITY && fabsf(y) < INFINITY) {
            x0 = fminf(x0, x);
            x1 = fmaxf(x1, x);
            y0 = fminf(y0, y);
            y1 = fmaxf(y1, y);
        }
    }
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) {
        x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, w));
        x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, w));
        y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, w));
        y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, w));
    }
    if (lane == 0) {
        red[0][wid] = x0;
        red[1][wid] = x1;
        red[2][wid] = y0;
        red[3][wid] = y1;
    }
    for (int c = tid; c < kSortCells; c += 1024) cell_pos[c] = 0;
    __syncthreads();
    x0 = red[0][lane];
    x1 = red[1][lane];
    y0 = red[2][lane];
    y1 = red[3][lane];
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) {
        x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, w));
        x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, w));
        y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, w));
        y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, w));
    }
    const float sx = (x1 > x0) ? (float)kSortGrid / (x1 - x0) : 0.0f, sy = (y1 > y0) ? (float)kSortGrid / (y1 - y0) : 0.0f;
    // histogram
    for (int i = tid; i < M; i += 1024) {
        const double2 o = obs[i];
        atomicAdd(&cell_pos[sort_cell((float)(o.x - org.x), (float)(o.y - org.y), x0, y0, sx, sy)], 1);
    }
    __syncthreads();
    // exclusive scan of the 4,096 counts: four consecutive cells per thread, warp scan, scan of the warp totals
    int c4[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c4[k] = cell_pos[tid * 4 + k];
        sum += c4[k];
    }
    int incl = sum;
#pragma unroll
    for (int w = 1; w < 32; w <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, w);
        if (lane >= w) incl += v;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int t = warp_tot[lane], ti = t;
#pragma unroll
        for (int w = 1; w < 32; w <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, ti, w);
            if (lane >= w) ti += v;
        }
        warp_tot[lane] = ti - t;   // exclusive
    }
    __syncthreads();
    int base = warp_tot[wid] + incl - sum;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        cell_pos[tid * 4 + k] = base;   // becomes the cell's write cursor
        base += c4[k];
    }
    __syncthreads();
    // scatter
    for (int i = tid; i < M; i += 1024) {
        const double2 o = obs[i];
        const int pos = atomicAdd(&cell_pos[sort_cell((float)(o.x - org.x), (float)(o.y - org.y), x0, y0, sx, sy)], 1);
        sorted[pos] = o;
    }
}

// pts / boxes / amax from the points in the order given; org_src points at the origin (the caller's obs[0])
__global__ void __launch_bounds__(128)
obstacle_prepare_kernel(int M, const double2 *__restrict__ obs, const double2 *__restrict__ org_src, float2 *__restrict__ pts,
                        float4 *__restrict__ boxes, ObsPrep *__restrict__ prep)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;        // one warp = one chunk of 32 points
    if ((i >> 5) >= (M + 31) / 32) return;                      // whole warps past the last chunk
    const double2 org = org_src[0];
    if (i == 0) prep->org = org;
    float2 v = make_float2(INFINITY, INFINITY);                // padding: never below a threshold
    float ax = 0.0f;
    float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
    if (i < M) {
        const double2 o = obs[i];
        v.x = (float)(o.x - org.x);
        v.y = (float)(o.y - org.y);
        ax = fmaxf(fabsf(v.x), fabsf(v.y));
        // a point with a NaN coordinate never hits in either precision: it stays out of the box (fminf/fmaxf drop NaN);
        // an Inf coordinate makes the box infinite, so its chunk is always examined
        xmin = xmax = v.x;
        ymin = ymax = v.y;
        if (v.x != v.x || v.y != v.y) {
            xmin = ymin = INFINITY;
            xmax = ymax = -INFINITY;
        }
    }
    pts[i] = v;                                                 // pts has room for the padded count
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) {
        xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, w));
        xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, w));
        ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, w));
        ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, w));
        ax = fmaxf(ax, __shfl_xor_sync(0xffffffffu, ax, w));
    }
    if ((threadIdx.x & 31) == 0) {
        boxes[i >> 5] = make_float4(xmin, ymin, xmax, ymax);
        atomicMax((int *)&prep->amax, __float_as_int(ax));      // non-negative floats order as ints
    }
}

template <int NC>
__global__ void __launch_bounds__(kColBlock)
collision_cull_kernel(int n_items, int n_pts, const __grid_constant__ CircleSpec cs, const double *__restrict__ px,
                      const double *__restrict__ py, const double *__restrict__ pcos, const double *__restrict__ psin,
                      const double *__restrict__ pyaw, int yaw_stride, int M, const double2 *__restrict__ obs,
                      const float2 *__restrict__ pts, const float4 *__restrict__ boxes, ObsPrep *prep,
                      unsigned char *free_out, YawFix yf)
{
    const int t = blockIdx.x * kColBlock + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool active = t < n_items;
    const int p = active ? t / n_pts : 0;
    if (active && ((volatile unsigned char *)free_out)[p] == 0) active = false;   // path already known to collide
    if (!__any_sync(0xffffffffu, active)) return;
    const int n_chunks = (M + 31) >> 5;
    // grid.y splits the chunk list (32 chunks = one ballot group per step) so that few paths x many obstacles still
    // fill the GPU; a path's verdict is the AND over the slices (free_out is only ever cleared)
    const int groups = (n_chunks + 31) >> 5;
    const int g_per = (groups + gridDim.y - 1) / gridDim.y;
    const int c_begin = blockIdx.y * g_per * 32, c_end = min(n_chunks, c_begin + g_per * 32);
    if (c_begin >= c_end) return;
    const double2 org = prep->org;   // (obs may be the spatially sorted copy: the origin is the caller's obs[0] either way)

    double cx[NC], cy[NC];
    float fx[NC], fy[NC], lo[NC], hi[NC];
    float bx0 = INFINITY, by0 = INFINITY, bx1 = -INFINITY, by1 = -INFINITY;   // this thread's reach box
    if (active) {
        const double x = px[t], y = py[t];
        double c, s;
        if (pcos) {
            c = pcos[t];
            s = psin[t];
        } else {
            sincos(pyaw[(size_t)p * yaw_stride + (t - p * n_pts)], &s, &c);
        }
        double ac = 0.0;
        bool bad = false;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            cx[k] = __dadd_rn(x, __dmul_rn(cs.off[k], c));
            cy[k] = __dadd_rn(y, __dmul_rn(cs.off[k], s));
            const double rx = cx[k] - org.x, ry = cy[k] - org.y;
            fx[k] = (float)rx;
            fy[k] = (float)ry;
            bad |= !(fabs(rx) <= 1.0e30) | !(fabs(ry) <= 1.0e30);
            ac = fmax(ac, fmax(fabs(rx), fabs(ry)));
        }
        const double u = 5.9604644775390625e-08;   // 2^-24
        double e0 = 2.0 * u * ((double)*(volatile float *)&prep->amax + ac) * 1.000001 + 1.0e-18;
        if (yf.list) {   // centres from the device's sincos: the band also covers any host trigonometry (yaw_bounds)
            double l64[NC], h64[NC];
            e0 += yaw_bounds<NC>(cs, cx, cy, l64, h64);
        }
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const double r = cs.rad[k];
            const bool sharp = !bad && e0 < 0.25 * r && r < 1.0e15;
            const double a = (r * (1.0 - 9.094947017729282e-13) - e0) / (1.0 + 4.0 * u);
            const double b = (r * (1.0 + 9.094947017729282e-13) + e0) / (1.0 - 4.0 * u);
            lo[k] = sharp ? __double2float_rd(a * a * (1.0 - 1.0e-15)) : 0.0f;
            hi[k] = sharp ? __double2float_ru(b * b * (1.0 + 1.0e-15)) : __int_as_float(0x7fc00000);
            if (!(r > 0.0)) hi[k] = 0.0f;
            // reach of circle k: beyond it q32 >= hi certainly (slack for the FP32 roundings of the box test itself)
            const float reach = (hi[k] == hi[k]) ? __fmul_ru(__fsqrt_ru(hi[k]), 1.00001f) : INFINITY;
            if (hi[k] != 0.0f) {
                bx0 = fminf(bx0, __fsub_rd(fx[k], reach));
                bx1 = fmaxf(bx1, __fadd_ru(fx[k], reach));
                by0 = fminf(by0, __fsub_rd(fy[k], reach));
                by1 = fmaxf(by1, __fadd_ru(fy[k], reach));
                if (reach == INFINITY) {   // not sharp: every chunk must be examined (and rechecked exactly)
                    bx0 = by0 = -INFINITY;
                    bx1 = by1 = INFINITY;
                }
            }
        }
    }
    // the warp's box
    float wx0 = bx0, wy0 = by0, wx1 = bx1, wy1 = by1;
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) {
        wx0 = fminf(wx0, __shfl_xor_sync(0xffffffffu, wx0, w));
        wy0 = fminf(wy0, __shfl_xor_sync(0xffffffffu, wy0, w));
        wx1 = fmaxf(wx1, __shfl_xor_sync(0xffffffffu, wx1, w));
        wy1 = fmaxf(wy1, __shfl_xor_sync(0xffffffffu, wy1, w));
    }
    float mn[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) mn[k] = INFINITY;
    const float4 *pts4 = reinterpret_cast<const float4 *>(pts);
    for (int g = c_begin; g < c_end; g += 32) {
        const int c = g + lane;
        bool over = false;
        if (c < c_end) {
            const float4 b = boxes[c];
            over = !(b.x > wx1 || b.z < wx0 || b.y > wy1 || b.w < wy0);   // an empty (all-NaN) box overlaps nothing
        }
        unsigned mask = __ballot_sync(0xffffffffu, over);
        if (mask && active && ((volatile unsigned char *)free_out)[p] == 0) active = false;
        if (!__any_sync(0xffffffffu, active)) return;
        if (mask && lane == 0) atomicAdd(&prep->screened, (unsigned long long)__popc(mask));
        while (mask) {
            const int ch = g + __ffs(mask) - 1;
            mask &= mask - 1;
            if (!active) continue;
#pragma unroll kCullUnroll
            for (int o = 0; o < 16; ++o) {
                const float4 ob = pts4[ch * 16 + o];                        // two points: (xa, ya, xb, yb)
                const unsigned long long X = pack2(ob.x, ob.z), Y = pack2(ob.y, ob.w);
#pragma unroll
                for (int k = 0; k < NC; ++k) {
                    const unsigned long long dx = add2(X, pack2(-fx[k], -fx[k])), dy = add2(Y, pack2(-fy[k], -fy[k]));
                    const unsigned long long q = fma2(dy, dy, mul2(dx, dx));
                    mn[k] = fminf(mn[k], fminf(lo2(q), hi2(q)));
                }
            }
        }
    }
    if (!active) return;
    bool hit = false, undecided = false;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        hit |= mn[k] < lo[k];
        undecided |= !(mn[k] >= hi[k]);
    }
    if (!hit && undecided) {
        // exact FP64 sequence over the chunks that overlap this thread's own box (the others are certainly free)
        double l64[NC], h64[NC];
        if (yf.list) {
            yaw_bounds<NC>(cs, cx, cy, l64, h64);
        } else {
#pragma unroll
            for (int k = 0; k < NC; ++k) l64[k] = h64[k] = cs.thr[k];
        }
        int bits = 0;
        for (int c = c_begin; c < c_end && !(bits & 1); ++c) {
            const float4 b = boxes[c];
            if (b.x > bx1 || b.z < bx0 || b.y > by1 || b.w < by0) continue;
            bits |= exact_tile_bits<NC>(l64, h64, cx, cy, obs, c * 32, min(M, c * 32 + 32));
            atomicAdd(&prep->rechecked, 1ULL);
        }
        hit = bits & 1;
        if (bits == 2) yaw_fix_append(yf, t);   // within ~1e-13 m of a circle: the caller decides with host cos / sin
    }
    if (hit) free_out[p] = 0;
}

// Resolves the path points the yaw-mode kernels left undecided: one CTA per listed item, exact FP64 sequence against
// every obstacle point with the cos / sin the caller evaluated on the host for exactly these points.
template <int NC>
__global__ void __launch_bounds__(128)
collision_resolve_kernel(int n_list, const int *__restrict__ items, const double *__restrict__ cos_sin, int n_pts,
                         const __grid_constant__ CircleSpec cs, const double *__restrict__ px, const double *__restrict__ py,
                         int M, const double2 *__restrict__ obs, unsigned char *free_out)
{
    const int i = blockIdx.x;
    if (i >= n_list) return;
    const int t = items[i];
    const int p = t / n_pts;
    if (((volatile unsigned char *)free_out)[p] == 0) return;
    const double x = px[t], y = py[t], c = cos_sin[i], s = cos_sin[n_list + i];
    double cx[NC], cy[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        cx[k] = __dadd_rn(x, __dmul_rn(cs.off[k], c));
        cy[k] = __dadd_rn(y, __dmul_rn(cs.off[k], s));
    }
    bool hit = false;
    for (int o = threadIdx.x; o < M; o += blockDim.x) {
        const double2 ob = obs[o];
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const double dx = __dsub_rn(ob.x, cx[k]);
            const double dy = __dsub_rn(ob.y, cy[k]);
            hit |= __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) < cs.thr[k];
        }
    }
    if (hit) free_out[p] = 0;
}

// ---------------------------------------------------------------------------------------------------
// Exact minimum clearance with the broad phase: the two FP32 sweeps of clearance_screen_kernel (pass A: min q32 per
// circle; pass B: exact q64 for the points with q32 <= T) visit only the 32-point obstacle chunks whose bounding box can
// still matter.  For a thread with circle centres inside [tx0, tx1] x [ty0, ty1] (FP32, relative to the origin) and a
// chunk box [bx0, bx1] x [by0, by1], every point p of the chunk has |p.x - c.x| >= gx = max(0, bx0 - tx1, tx0 - bx1)
// for every circle (likewise y), and because FP32 subtraction, multiplication and FMA are monotone, the q32 the sweep
// would compute is >= lb = fma_rd(gy, gy, mul_rd(gx, gx)) with gx, gy rounded DOWN.  Pass A skips a chunk when
// lb > max_k(current min q32_k) -- it cannot lower any minimum -- and pass B when lb > max_k T_k -- it holds no
// candidate.  Both tests are strict and NaN-safe (an undecidable comparison evaluates the chunk), non-sharp threads have
// T = +inf and visit everything, so the doubles returned are those of clearance_screen_kernel / collision_kernel<NC, true>
// for every input.  Obstacle outlines are spatially coherent, so after the first few chunks the running minimum prunes
// nearly all of the rest: config 3 evaluates a few per cent of the 313 chunks per path point.
template <int NC>
__global__ void __launch_bounds__(kColBlock)
clearance_cull_kernel(int n_items, int n_pts, const __grid_constant__ CircleSpec cs, const double *__restrict__ px,
                      const double *__restrict__ py, const double *__restrict__ pcos, const double *__restrict__ psin,
                      const double *__restrict__ pyaw, int yaw_stride, int M, const double2 *__restrict__ obs,
                      const float2 *__restrict__ pts, const float4 *__restrict__ boxes, const ObsPrep *__restrict__ prep,
                      unsigned char *free_out, double *__restrict__ clear_pts)
{
    const int t = blockIdx.x * kColBlock + threadIdx.x;
    if (t >= n_items) return;
    const int p = t / n_pts;
    const double2 org = prep->org;   // (obs may be the spatially sorted copy: the origin is the caller's obs[0] either way)
    const int n_chunks = (M + 31) >> 5;
    double cx[NC], cy[NC], ac = 0.0;
    unsigned long long ncx[NC], ncy[NC];
    float tx0 = INFINITY, tx1 = -INFINITY, ty0 = INFINITY, ty1 = -INFINITY;
    bool bad = false;
    {
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
            const double rx = cx[k] - org.x, ry = cy[k] - org.y;
            const float fx = (float)rx, fy = (float)ry;
            ncx[k] = pack2(-fx, -fx);
            ncy[k] = pack2(-fy, -fy);
            tx0 = fminf(tx0, fx);
            tx1 = fmaxf(tx1, fx);
            ty0 = fminf(ty0, fy);
            ty1 = fmaxf(ty1, fy);
            bad |= !(fabs(rx) <= 1.0e30) | !(fabs(ry) <= 1.0e30);   // NaN / Inf / beyond FP32 range
            ac = fmax(ac, fmax(fabs(rx), fabs(ry)));
        }
        if (bad) tx0 = ty0 = tx1 = ty1 = __int_as_float(0x7fc00000);   // every bound test undecidable -> every chunk visited
    }
    const float4 *pts4 = reinterpret_cast<const float4 *>(pts);
    // lower bound of q32 over the points of chunk c, for every circle of this thread (rounded down)
    auto chunk_lb = [&](int c) {
        const float4 b = boxes[c];
        const float gx = fmaxf(0.0f, fmaxf(__fsub_rd(b.x, tx1), __fsub_rd(tx0, b.z)));
        const float gy = fmaxf(0.0f, fmaxf(__fsub_rd(b.y, ty1), __fsub_rd(ty0, b.w)));
        return __fmaf_rd(gy, gy, __fmul_rd(gx, gx));
    };

    // ---- pass A: min q32 per circle
    float mn[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) mn[k] = INFINITY;
    auto eval_chunk = [&](int c) {
#pragma unroll 4
        for (int o = 0; o < 16; ++o) {
            const float4 ob = pts4[c * 16 + o];                        // two points: (xa, ya, xb, yb)
            const unsigned long long X = pack2(ob.x, ob.z), Y = pack2(ob.y, ob.w);
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                const unsigned long long dx = add2(X, ncx[k]), dy = add2(Y, ncy[k]);
                const unsigned long long q = fma2(dy, dy, mul2(dx, dx));
                mn[k] = fminf(mn[k], fminf(lo2(q), hi2(q)));
            }
        }
    };
    // Seed: the chunk whose box is nearest to this thread goes first.  The lanes of a warp evaluate their (different) seed
    // chunks in the same instructions, and every lane then enters the sweep with a minimum close to its final one -- without
    // it a lane keeps evaluating chunks until it has met a near one, and a warp executes the UNION of its lanes' chunks
    // (measured: 0.79 ms with the plain sweep).
    int seed = 0;
    {
        float best = INFINITY;
        for (int c = 0; c < n_chunks; ++c) {
            const float lb = chunk_lb(c);
            if (lb < best) {
                best = lb;
                seed = c;
            }
        }
    }
    eval_chunk(seed);
    for (int c = 0; c < n_chunks; ++c) {
        float mx = mn[0];
#pragma unroll
        for (int k = 1; k < NC; ++k) mx = fmaxf(mx, mn[k]);
        if ((chunk_lb(c) > mx && !bad) || c == seed) continue;
        eval_chunk(c);
    }

    // ---- candidate thresholds (as in clearance_screen_kernel)
    float T[NC], maxT = 0.0f;
    {
        const double u = 5.9604644775390625e-08;   // 2^-24
        const double e0 = 2.0 * u * ((double)prep->amax + ac) * 1.000001 + 1.0e-18;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const double b = (sqrt((double)mn[k]) * (1.0 + 4.0 * u) + 2.0 * e0) / (1.0 - 4.0 * u);
            const bool sharp = !bad && e0 < 1.0e30 && mn[k] < INFINITY;
            T[k] = sharp ? __double2float_ru(b * b * 1.000001) : INFINITY;
            maxT = fmaxf(maxT, T[k]);
        }
    }

    // ---- pass B: exact q64 for the candidates only
    double qmin[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) qmin[k] = INFINITY;
    for (int c = 0; c < n_chunks; ++c) {
        if (chunk_lb(c) > maxT && !bad) continue;
        for (int o = 0; o < 16; ++o) {
            const float4 ob = pts4[c * 16 + o];
            const unsigned long long X = pack2(ob.x, ob.z), Y = pack2(ob.y, ob.w);
            bool cand = false;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                const unsigned long long dx = add2(X, ncx[k]), dy = add2(Y, ncy[k]);
                const unsigned long long q = fma2(dy, dy, mul2(dx, dx));
                cand |= !(fminf(lo2(q), hi2(q)) > T[k]);   // NaN counts as a candidate (it drops out in fmin below)
            }
            if (cand) {
                for (int i = c * 32 + 2 * o; i < c * 32 + 2 * o + 2 && i < M; ++i) {
                    const double2 e = obs[i];
#pragma unroll
                    for (int k = 0; k < NC; ++k) {
                        const double dx = __dsub_rn(e.x, cx[k]);
                        const double dy = __dsub_rn(e.y, cy[k]);
                        qmin[k] = fmin(qmin[k], __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                    }
                }
            }
        }
    }
    bool hit = false;
    double clr = INFINITY;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        hit |= qmin[k] < cs.thr[k];                                              // sqrt_rn(q) - r < 0  <=>  q < thr
        clr = fmin(clr, __dsub_rn(__dsqrt_rn(qmin[k]), cs.rad[k]));             // collision_checker.py:101-105
    }
    if (hit) free_out[p] = 0;
    clear_pts[t] = clr;
}

__global__ void clearance_reduce_kernel(int P, int n_pts, const double *__restrict__ clear_pts, double *__restrict__ out)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    double m = INFINITY;
    for (int j = 0; j < n_pts; ++j) m = fmin(m, clear_pts[(size_t)p * n_pts + j]);
    out[p] = m;
}

// Broad-phase inputs in `area` (pts | boxes | prep | sorted FP64 points): returns the FP64 point array the cull kernels
// must index (the spatially sorted copy, or the caller's array for very long lists).
static size_t broad_phase_bytes(int M)
{
    const int n_chunks = (M + 31) / 32;
    return sizeof(float2) * 32 * (size_t)n_chunks + sizeof(float4) * (size_t)n_chunks + sizeof(ObsPrep) + sizeof(double2) * (size_t)M + 64;
}
static int broad_phase_prepare(cudaStream_t st, int M, const double *obs, void *area, float2 **pts, float4 **boxes, ObsPrep **prep,
                               const double2 **obs_eff)
{
    const int n_chunks = (M + 31) / 32;
    const size_t pts_bytes = sizeof(float2) * 32 * (size_t)n_chunks, box_bytes = sizeof(float4) * (size_t)n_chunks;
    *pts = (float2 *)area;
    *boxes = (float4 *)((char *)area + pts_bytes);
    *prep = (ObsPrep *)((char *)area + pts_bytes + box_bytes);
    double2 *sorted = (double2 *)((char *)area + ((pts_bytes + box_bytes + sizeof(ObsPrep) + 15) & ~(size_t)15));
    B200MP_CUDA(cudaMemsetAsync(*prep, 0, sizeof(ObsPrep), st));
    *obs_eff = (const double2 *)obs;
    bool sort = M <= kSortMaxPoints;
#ifdef B200MP_DEV_TUNABLES
    if (getenv("B200MP_NO_OBS_SORT")) sort = false;   // development A/B (tools/collision_quick.py, tools/cbench.cu)
#endif
    if (sort) {
        obstacle_sort_kernel<<<1, 1024, 0, st>>>(M, (const double2 *)obs, sorted);
        B200MP_CUDA(cudaGetLastError());
        *obs_eff = sorted;
    }
    obstacle_prepare_kernel<<<(n_chunks * 32 + 127) / 128, 128, 0, st>>>(M, *obs_eff, (const double2 *)obs, *pts, *boxes, *prep);
    B200MP_CUDA(cudaGetLastError());
    return 0;
}

template <int NC>
static int launch_collision_nc(int device, cudaStream_t st, int P, int n_pts, const CircleSpec &cs, const double *px,
                               const double *py, const double *pcos, const double *psin, const double *pyaw,
                               int yaw_stride, int M, const double *obs, unsigned char *free_out, double *min_clear,
                               YawFix yf, int mode)
{
    const long long items = (long long)P * n_pts;
    const int grid = (int)((items + kColBlock - 1) / kColBlock);
    if (min_clear) {
        void *scratch = nullptr;
        int rc = ensure_scratch(device, st, sizeof(double) * (size_t)items, &scratch);
        if (rc) return rc;
        if (mode == B200MP_COLLISION_FP64_ONLY) {
            collision_kernel<NC, true><<<grid, kColBlock, 0, st>>>((int)items, n_pts, cs, px, py, pcos, psin, pyaw, yaw_stride,
                                                                  M, (const double2 *)obs, free_out, (double *)scratch);
        } else if (mode == B200MP_COLLISION_SCREEN_ONLY || items * (long long)M < (1LL << 22)) {
            clearance_screen_kernel<NC><<<grid, kColBlock, 0, st>>>((int)items, n_pts, cs, px, py, pcos, psin, pyaw, yaw_stride,
                                                                   M, (const double2 *)obs, free_out, (double *)scratch);
        } else {
            // broad phase: its inputs sit behind the per-point clearances in the same scratch area
            const size_t clr_bytes = (sizeof(double) * (size_t)items + 255) & ~(size_t)255;
            rc = ensure_scratch(device, st, clr_bytes + broad_phase_bytes(M), &scratch);
            if (rc) return rc;
            float2 *pts;
            float4 *boxes;
            ObsPrep *prep;
            const double2 *obs_eff;
            rc = broad_phase_prepare(st, M, obs, (char *)scratch + clr_bytes, &pts, &boxes, &prep, &obs_eff);
            if (rc) return rc;
            clearance_cull_kernel<NC><<<grid, kColBlock, 0, st>>>((int)items, n_pts, cs, px, py, pcos, psin, pyaw, yaw_stride, M,
                                                                 obs_eff, pts, boxes, prep, free_out, (double *)scratch);
        }
        B200MP_CUDA(cudaGetLastError());
        clearance_reduce_kernel<<<(P + 127) / 128, 128, 0, st>>>(P, n_pts, (const double *)scratch, min_clear);
    } else if (mode == B200MP_COLLISION_FP64_ONLY) {
        if (yf.list)
            collision_kernel<NC, false, true><<<grid, kColBlock, 0, st>>>((int)items, n_pts, cs, px, py, pcos, psin, pyaw,
                                                                         yaw_stride, M, (const double2 *)obs, free_out, nullptr, yf);
        else
            collision_kernel<NC, false><<<grid, kColBlock, 0, st>>>((int)items, n_pts, cs, px, py, pcos, psin, pyaw, yaw_stride,
                                                                   M, (const double2 *)obs, free_out, nullptr);
    } else if (mode == B200MP_COLLISION_SCREEN_ONLY || items * (long long)M < (1LL << 22)) {
        // (also for small problems such as the planner's own 7 paths x 106 points: one launch, no prepare pass)
        const dim3 g2(grid, (M + kFTile - 1) / kFTile);
        collision_filter_kernel<NC><<<g2, kColBlock, 0, st>>>((int)items, n_pts, cs, px, py, pcos, psin, pyaw, yaw_stride, M,
                                                             (const double2 *)obs, free_out, yf);
    } else {
        const int n_chunks = (M + 31) / 32;
        void *scratch = nullptr;
        int rc = ensure_scratch(device, st, broad_phase_bytes(M), &scratch);
        if (rc) return rc;
        float2 *pts;
        float4 *boxes;
        ObsPrep *prep;
        const double2 *obs_eff;
        rc = broad_phase_prepare(st, M, obs, scratch, &pts, &boxes, &prep, &obs_eff);
        if (rc) return rc;
        // grid.y splits the chunk list in ballot groups of 32 chunks, one group per CTA: a path's work is spread over many
        // small CTAs, which evens out the load (free paths near obstacles are the expensive ones) and lets the early exit of
        // a colliding path reach the other slices sooner.  Measured on 4,096 paths x 10,000 points (tools/cbench.cu): 0.52 ms
        // with one CTA per 128 path points, 0.32 / 0.25 / 0.20 ms with 3 / 5 / 10 slices (0.18 ms with the inner loop fully
        // unrolled); the per-thread set-up repeated in every slice is ~25 % of a slice's work.  Capped at 2^18 CTAs.
        const int groups = (n_chunks + 31) / 32;
        int gy = groups;
        if ((long long)gy * grid > (1LL << 18)) gy = (int)((1LL << 18) / grid);
#ifdef B200MP_DEV_TUNABLES
        if (const char *ev = getenv("B200MP_CULL_GY")) gy = atoi(ev);   // development A/B (tools/cbench.cu)
#endif
        gy = gy < 1 ? 1 : (gy > groups ? groups : gy);
        collision_cull_kernel<NC><<<dim3(grid, gy), kColBlock, 0, st>>>((int)items, n_pts, cs, px, py, pcos, psin, pyaw, yaw_stride,
                                                                     M, obs_eff, pts, boxes, prep, free_out, yf);
        // the statistics of THIS launch sit at this offset until the next user of the scratch area overwrites them
        DeviceState &ds = dev_state(device);
        ds.cull_stats_ptr = prep;
        ds.cull_stats_stream = st;
    }
    B200MP_CUDA(cudaGetLastError());
    return 0;
}

// last launch's broad-phase statistics (synchronises the stream); zeros unless the broad-phase kernel was the last
// user of this stream's scratch area
int collision_stats(int device, cudaStream_t st, int M, unsigned long long out[2])
{
    (void)M;
    DeviceState &ds = dev_state(device);
    out[0] = out[1] = 0;
    if (!ds.cull_stats_ptr || ds.cull_stats_stream != st) return 0;
    ObsPrep h;
    B200MP_CUDA(cudaMemcpyAsync(&h, ds.cull_stats_ptr, sizeof(h), cudaMemcpyDeviceToHost, st));
    B200MP_CUDA(cudaStreamSynchronize(st));
    out[0] = h.screened;
    out[1] = h.rechecked;
    return 0;
}

static int collision_args_ok(int P, int n_pts, int n_circ, int M, const double *off, const double *rad,
                             const unsigned char *free_out)
{
    if (P < 0 || n_pts < 0 || M < 0 || n_circ < 1 || n_circ > kMaxCircles) {
        set_error("collision_check: bad sizes P=%d n_pts=%d n_circ=%d (1..%d) M=%d", P, n_pts, n_circ, kMaxCircles, M);
        return B200MP_E_ARG;
    }
    if (P > 0 && (!off || !rad || !free_out)) {
        set_error("collision_check: off, rad and free_out must be non-NULL");
        return B200MP_E_ARG;
    }
    if ((long long)P * n_pts > 0x7fffffffLL) {
        set_error("collision_check: P*n_pts exceeds 2^31-1");
        return B200MP_E_ARG;
    }
    return 0;
}

static CircleSpec make_circle_spec(int n_circ, const double *off, const double *rad)
{
    CircleSpec cs{};
    for (int k = 0; k < n_circ; ++k) {
        cs.off[k] = off[k];
        cs.rad[k] = rad[k];
        cs.thr[k] = sqrt_threshold(rad[k]);
    }
    return cs;
}

int launch_collision_f64(int device, cudaStream_t st, int P, int n_pts, int n_circ, const double *off,
                         const double *rad, const double *px, const double *py, const double *pcos,
                         const double *psin, const double *pyaw, int yaw_stride, int M, const double *obs,
                         unsigned char *free_out, double *min_clear, int *undecided, int undecided_capacity, int mode)
{
    int rc = collision_args_ok(P, n_pts, n_circ, M, off, rad, free_out);
    if (rc) return rc;
    if (mode < 0) mode = collision_mode();
    if (undecided) B200MP_CUDA(cudaMemsetAsync(undecided, 0, sizeof(int), st));
    if (P == 0) return 0;
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
    if ((((size_t)obs) & 15) != 0) {
        set_error("collision_check: obs must be 16-byte aligned");
        return B200MP_E_ARG;
    }
    if (undecided && (min_clear || (pcos && psin) || undecided_capacity < 0)) {
        set_error("collision_check: the undecided list goes with pyaw input and flags only");
        return B200MP_E_ARG;
    }
    const CircleSpec cs = make_circle_spec(n_circ, off, rad);
    if (!pcos || !psin) pcos = psin = nullptr;
    const YawFix yf{undecided, undecided_capacity};
    switch (n_circ) {
#define B200MP_NC(N) \
    case N: return launch_collision_nc<N>(device, st, P, n_pts, cs, px, py, pcos, psin, pyaw, yaw_stride, M, obs, free_out, min_clear, yf, mode);
        B200MP_NC(1) B200MP_NC(2) B200MP_NC(3) B200MP_NC(4) B200MP_NC(5) B200MP_NC(6) B200MP_NC(7) B200MP_NC(8)
#undef B200MP_NC
    }
    return B200MP_E_ARG;
}

int launch_collision_resolve_f64(int device, cudaStream_t st, int n_list, const int *items, const double *cos_sin, int P,
                                 int n_pts, int n_circ, const double *off, const double *rad, const double *px,
                                 const double *py, int M, const double *obs, unsigned char *free_out)
{
    (void)device;
    int rc = collision_args_ok(P, n_pts, n_circ, M, off, rad, free_out);
    if (rc) return rc;
    if (n_list < 0 || (n_list > 0 && (!items || !cos_sin || !px || !py))) {
        set_error("collision_resolve: bad list (n_list=%d)", n_list);
        return B200MP_E_ARG;
    }
    if (n_list == 0 || M == 0 || P == 0) return 0;
    if (!obs || (((size_t)obs) & 15) != 0) {
        set_error("collision_resolve: obs must be non-NULL and 16-byte aligned");
        return B200MP_E_ARG;
    }
    const CircleSpec cs = make_circle_spec(n_circ, off, rad);
    switch (n_circ) {
#define B200MP_NC(N) \
    case N: collision_resolve_kernel<N><<<n_list, 128, 0, st>>>(n_list, items, cos_sin, n_pts, cs, px, py, M, (const double2 *)obs, free_out); break;
        B200MP_NC(1) B200MP_NC(2) B200MP_NC(3) B200MP_NC(4) B200MP_NC(5) B200MP_NC(6) B200MP_NC(7) B200MP_NC(8)
#undef B200MP_NC
    }
    B200MP_CUDA(cudaGetLastError());
    return 0;
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

// One WARP per candidate path i.  The reference adds weight * norm(end_i - end_j) for the colliding j in ascending j
// (collision_checker.py:181-190); floating-point addition is not associative, so the order is kept: the 32 lanes
// evaluate the 32 norms of a chunk of j in parallel (the expensive part: a correctly rounded square root each), and the
// chunk is then folded into the running score one lane after the other through shuffles.  A collision-free j
// contributes +0.0, which leaves any score unchanged bit for bit, so no compaction is needed; chunks without a
// colliding j are skipped.  P = 4,096: 0.36 ms with one thread per i -> 0.05 ms.
__global__ void __launch_bounds__(128)
select_score_kernel(int P, const double *__restrict__ ex, const double *__restrict__ ey,
                    const unsigned char *__restrict__ free_in, double gx, double gy, double weight, int mode,
                    double *__restrict__ scores)
{
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= P) return;
    const double xi = ex[i], yi = ey[i];
    // free_in: 1 = collision-free candidate, 0 = colliding (penalises the others), any other value = excluded: neither a
    // candidate nor a penalty term -- a path the planner dropped before the selection (local_planner.py:317-323)
    const bool fi = free_in[i] == 1;
    double score = norm2_host_form(__dsub_rn(xi, gx), __dsub_rn(yi, gy), mode);   // collision_checker.py:175
    if (fi) {                                                                       // a colliding i scores +inf (:196)
        for (int j0 = 0; j0 < P; j0 += 32) {
            const int j = j0 + lane;
            const bool coll = j < P && free_in[j] == 0;
            if (!__any_sync(0xffffffffu, coll)) continue;
            double v = 0.0;
            if (coll) v = __dmul_rn(weight, norm2_host_form(__dsub_rn(xi, ex[j]), __dsub_rn(yi, ey[j]), mode));
#pragma unroll
            for (int l = 0; l < 32; ++l) score = __dadd_rn(score, __shfl_sync(0xffffffffu, v, l));   // ascending j
        }
    }
    if (lane == 0) scores[i] = fi ? score : INFINITY;
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
    int rc = ensure_scratch(device, st, score_bytes + argmin_scratch_bytes(P), &scratch);
    if (rc) return rc;
    double *scores = scores_out ? scores_out : (double *)scratch;
    if (P > 0) {
        select_score_kernel<<<(P + 3) / 4, 128, 0, st>>>(P, ex, ey, free_in, gx, gy, weight, norm_mode, scores);
        B200MP_CUDA(cudaGetLastError());
    }
    return argmin_launch(st, P, scores, 0, (char *)scratch + score_bytes, nullptr, nullptr, best_out);
}

}  // namespace b200mp
