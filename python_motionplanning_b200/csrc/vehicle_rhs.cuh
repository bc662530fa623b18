// 7-DoF planar vehicle model: right-hand side and classic RK4 step, one rollout per thread.
//
// Computes what the reference's VehicleModel.planar_model (libs/vehicle_model/vehicle_model.py:220-425)
// and VehicleModel.planar_model_RK4 (:427-445) compute, re-derived for a register-resident thread
// program rather than transcribed:
//   * parameter-only sub-expressions (static loads, load-transfer gains, 1/m, 1/Izz, 1/Jw, T/2) are
//     folded on the host into DevParams, which reaches the kernel as a by-value argument (constant
//     bank operands, no registers);
//   * sin/cos of the steer angles are hoisted out of the four stages (controls are frozen across a
//     step, :429-436) and out of every step of a zero-order-hold segment (drive.py:128);
//   * normal loads depend only on ax_prev, ay_prev and are formed once per step (:255-258);
//   * each wheel needs one reciprocal (1/vx serves both slips, :284-293) and one reciprocal square
//     root (sqrt(q) = q*rsqrt(q) and mu/s = mu*rsqrt(q), :296-348);
//   * x and y never feed back into the right-hand side, so stage states carry 8 components;
//   * RK4 is kept in running-sum form (acc += w_s * K_s) so only y, acc, ys and one K are live.
// Quirks of the reference that are reproduced on purpose: mu_max replaces Pacejka D;
// rear wheel-spin uses the chassis-frame force; lateral slip divides by |vx|; zero combined slip gives
// zero friction; yaw is never wrapped; ax_prev/ay_prev are the RK4-averaged axc/ayc of the previous step.
#pragma once

#include "b200mp_math.cuh"

namespace b200mp {

constexpr double kGravity = 9.81;  // vehicle_model.py:230

// Device-ready parameter set (derived on the host in double, reference operator order).
template <typename R> struct DevParams {
    R inv_m, inv_Izz, inv_Jw, a, b, halfT, rw;
    R rw_inv_Jw;   // rw / Jw: the wheel-spin equation is evaluated as tq/Jw - (rw/Jw) f, one FMA per wheel-stage
    R Fz0F, Fz0R, DfzxL, DfzxR, DfzyF, DfzyR;
    R Bc[4], Cc[4], Dc[4];
};

// Plain-double mirror of B200mpVehicleParams (kept layout-identical; see b200mp.h)
struct HostParams {
    double m, a, b, Izz, Jw, hg, T, wL, wR, rw;
    double B[4], C[4], D[4];
};

template <typename R> inline DevParams<R> derive_params(const HostParams &p)
{
    const double g = kGravity;
    DevParams<R> d;
    d.inv_m = (R)(1 / p.m);
    d.inv_Izz = (R)(1 / p.Izz);
    d.inv_Jw = (R)(1 / p.Jw);
    d.a = (R)p.a;
    d.b = (R)p.b;
    d.halfT = (R)(p.T / 2);
    d.rw = (R)p.rw;
    d.rw_inv_Jw = (R)(p.rw / p.Jw);
    d.Fz0F = (R)(p.b / (p.a + p.b) * p.m * g / 2);                      // :245-248
    d.Fz0R = (R)(p.a / (p.a + p.b) * p.m * g / 2);
    d.DfzxL = (R)(p.m * p.hg * p.wR / ((p.a + p.b) * (p.wL + p.wR)));   // :250-253
    d.DfzxR = (R)(p.m * p.hg * p.wL / ((p.a + p.b) * (p.wL + p.wR)));
    d.DfzyF = (R)(p.m * p.hg * p.b / ((p.a + p.b) * (p.wL + p.wR)));
    d.DfzyR = (R)(p.m * p.hg * p.a / ((p.a + p.b) * (p.wL + p.wR)));
    for (int i = 0; i < 4; ++i) {
        d.Bc[i] = (R)p.B[i];
        d.Cc[i] = (R)p.C[i];
        d.Dc[i] = (R)p.D[i];
    }
    return d;
}

// ---- tabulated combined-slip friction (fast path) -------------------------------------------------
// Per wheel and stage the model needs  mu/s = D sin(C atan(B s)) / s  with s = sqrt(sx^2 + sy^2)
// (vehicle_model.py:296-348): a square root, a reciprocal, an atan and a sin, ~45 FP64 instructions even
// with the custom routines of b200mp_math.cuh -- 60 % of the kernel.  As a function of
//     x = 1 + (B s)^2 = 1 + B^2 (sx^2 + sy^2)
// the quantity  G(x) = D B sin(C atan(sqrt(x-1))) / sqrt(x-1)  (so that mu_x = sx G, mu_y = sy G) is analytic
// on x > 0 (sin(C atan t)/t is even in t; the nearest singularity is x = 0), needs neither the square
// root nor the reciprocal, and is 0/0-free at zero slip (G(1) = D B C).  For ONE tyre (B, C, D) it is
// tabulated on the host when the parameter set is uploaded: 64 intervals per binade of x (bounded by the
// six leading mantissa bits, so the interval index and midpoint come from the bit pattern of x), a
// degree-6 polynomial in t = x - midpoint per interval from a Chebyshev interpolant computed in long double:
//     G ~ c0 + t (c1 + t (c2 + t (c3 + t [c4 + t c5 + t^2 c6])))   with the bracket evaluated in FP32 (2 FFMA) and the
// rest in FP64 (4 DFMA).  The bracket's term is < 2^-28 of G, so its FP32 rounding stays below 2^-52 G; it runs on the
// otherwise idle FP32 pipe.  (Measured: on this kernel the cost of a step is ~ 2 x FP64 instructions + 1 x all others,
// so the variant that adds the FP32 part last -- shorter dependency chain, 3 more FP32 multiplies -- is slower.)  One row is 4 doubles + 3 floats (+ 4 bytes of padding) = 48 bytes = three
// 128-bit shared-memory loads; a 48-byte stride maps 8 consecutive rows onto the 8 distinct 16-byte bank
// groups.  The relative error, measured on the host against the long-double function on the exact device
// scheme, is <= 4e-16 (rounding level).  The table covers x < 2^14 (B s < 128, i.e. slip beyond 6 for any
// realistic B; 896 rows = 42 KB); anything outside, NaN included, clears the speculative step's `ok` flag
// and the step is repeated on the closed-form path.
// Cost per wheel-stage from q = sx^2 + sy^2 on: 7 FP64 instructions + 2 conversions instead of 49.
// B200MP_MU_DEG5: 128 intervals per binade and a degree-5 polynomial entirely in FP64 (5 DFMA, six doubles per row, no
// FP32 tail: two F2F conversions and two FFMA fewer per wheel-stage for one more DFMA); the table is twice as large
// (96 KB with MASKIDX), so the rollout kernels then run 128-thread CTAs, two per SM.
#ifndef B200MP_MU_DEG5
#define B200MP_MU_DEG5 0
#endif
constexpr bool kMuDeg5 = B200MP_MU_DEG5 != 0;
constexpr int kMuBits = kMuDeg5 ? 7 : 6;                  // leading mantissa bits used for the interval index
constexpr int kMuPerBinade = 1 << kMuBits;                // 64 (128)
constexpr int kMuBinades = 14;
constexpr int kMuIntervals = kMuPerBinade * kMuBinades;   // 896 rows = 42 KB (1,792 rows = 84 KB)
constexpr int kMuCoefD = kMuDeg5 ? 6 : 4;                 // degrees 0..3 (0..5) in FP64
constexpr int kMuCoefF = kMuDeg5 ? 0 : 3;                 // degrees 4..6 in FP32
constexpr int kMuCoef = kMuCoefD + kMuCoefF;              // degree 6 (5)
constexpr int kMuStride = 6;                              // row stride in doubles (48 bytes)
// Row addressing.  MASKIDX: the row of x is selected by the raw bit field [low 4 exponent bits | kMuBits leading mantissa
// bits] of x (one LOP3), its byte offset is one IMAD.HI of that field, and the table holds 16 binades' worth of rows in
// the order of that field (row = ((e + 15) & 15) * 64 + m for x in binade e; the two unused binades are zero rows) --
// 48 KB instead of 42 KB, but no subtract / shift / clamp-select per wheel-stage: whatever the bits of x are, the
// address is inside the table, and "x is in the tabulated range" is a single unsigned compare of the high word.
#ifndef B200MP_MU_MASKIDX
#define B200MP_MU_MASKIDX 0
#endif
constexpr bool kMuMaskIdx = B200MP_MU_MASKIDX != 0;
constexpr int kMuRows = kMuMaskIdx ? (16 << kMuBits) : kMuIntervals;
constexpr int kMuTableDoubles = kMuRows * kMuStride;
constexpr int kMuKeyShift = 20 - kMuBits;
constexpr unsigned kMuKeyMask = (unsigned)((16 << kMuBits) - 1) << kMuKeyShift;
constexpr unsigned kMuKeyToBytes = (unsigned)(kMuStride * 8) << (32 - kMuKeyShift);   // umulhi(key, .) = row * 48
static_assert((unsigned long long)(kMuStride * 8) << (32 - kMuKeyShift) < (1ull << 32), "byte-offset multiplier must fit 32 bits");
// memory row of interval k = e * kMuPerBinade + m
inline constexpr int mu_row_of(int k)
{
    return kMuMaskIdx ? (((((k >> kMuBits) + 15) & 15) << kMuBits) | (k & (kMuPerBinade - 1))) : k;
}
#ifndef B200MP_MU_CACHE_ROWS
#define B200MP_MU_CACHE_ROWS 1
#endif
constexpr bool kMuCacheRows = B200MP_MU_CACHE_ROWS != 0;

// One table row as it sits in memory (16-byte aligned, loaded as three 128-bit words).
#if B200MP_MU_DEG5
struct alignas(16) MuRow {
    double c3, c2, c1, c0, c5, c4;
};
#else
struct alignas(16) MuRow {
    double c3, c2, c1, c0;
    float c6, c5, c4, pad;
};
#endif
static_assert(sizeof(MuRow) == kMuStride * sizeof(double), "MuRow must be 48 bytes");

struct MuTableView {
    const void *c;     // the table (MuRow[kMuIntervals] for FP64, MuRowF[kMuIntervalsF] for FP32); shared memory in the kernels
    double B2;         // B^2
};

// Evaluation of one table row exactly as the device does it (also used by the host audit and hostsim).
B200MP_HD double mu_row_eval(const MuRow &r, double t)
{
#if B200MP_MU_DEG5
    return fma(fma(fma(fma(fma(r.c5, t, r.c4), t, r.c3), t, r.c2), t, r.c1), t, r.c0);
#else
    const float tf = (float)t;
    const float tail = fmaf(fmaf(r.c6, tf, r.c5), tf, r.c4);
    return fma(fma(fma(fma((double)tail, t, r.c3), t, r.c2), t, r.c1), t, r.c0);
#endif
}

// 128-bit loads of one row (the compiler will not merge scalar shared-memory loads on its own)
B200MP_HD unsigned mu_key_to_bytes(int key)   // key = the masked high-word bits of x (MASKIDX)
{
#if defined(__CUDA_ARCH__)
    return __umulhi((unsigned)key, kMuKeyToBytes);
#else
    return (unsigned)(((unsigned long long)(unsigned)key * kMuKeyToBytes) >> 32);
#endif
}

// k: row index, or (MASKIDX) the key returned by MuTab<double>::locate
B200MP_HD MuRow mu_row_load(const double *table, int k)
{
    MuRow r;
#if defined(__CUDA_ARCH__)
    const double2 *p = kMuMaskIdx ? reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(table) + mu_key_to_bytes(k))
                                  : reinterpret_cast<const double2 *>(table + k * kMuStride);
    const double2 a = p[0], b = p[1];
    r.c3 = a.x; r.c2 = a.y; r.c1 = b.x; r.c0 = b.y;
#if B200MP_MU_DEG5
    const double2 c = p[2];
    r.c5 = c.x; r.c4 = c.y;
#else
    const float4 f = *reinterpret_cast<const float4 *>(p + 2);
    r.c6 = f.x; r.c5 = f.y; r.c4 = f.z; r.pad = 0.0f;
#endif
#else
    r = kMuMaskIdx ? *reinterpret_cast<const MuRow *>(reinterpret_cast<const char *>(table) + mu_key_to_bytes(k))
                   : *reinterpret_cast<const MuRow *>(table + k * kMuStride);
#endif
    return r;
}

B200MP_HD double mu_table_eval(const double *row, double t)
{
    return mu_row_eval(*reinterpret_cast<const MuRow *>(row), t);
}

// The table row of each wheel, kept in registers across the four stages of an RK4 step: the slip of a wheel moves
// by less than an interval (1/64 of a binade of x) between most stages, so stages 2-4 reload a row only for the lanes
// whose interval changed -- shared-memory traffic drops, and so do the bank conflicts of a warp whose rollouts
// sit in many different intervals.
template <typename Row> struct MuRowCacheT {
    int k[4];
    Row r[4];
    double rinv[4];   // B200MP_CARRY_RCP: 1/vx of each wheel at the previous stage of the step
};
typedef MuRowCacheT<MuRow> MuRowCache;

// Host: monomial coefficients (in x - mid) of the degree n-1 Chebyshev interpolant of G on [lo, hi], in long double.
template <typename F> inline void mu_fit_interval(F G, long double lo, long double hi, int n, long double *mono, long double *mid_out)
{
    typedef long double L;
    const L pi = 3.14159265358979323846264338327950288L;
    const L mid = 0.5L * (lo + hi), half = 0.5L * (hi - lo);
    L f[16], c[16], T0[16] = {0}, T1[16] = {0};
    for (int i = 0; i < n; ++i) f[i] = G(mid + half * cosl(pi * (2 * i + 1) / (2 * n)));
    for (int j = 0; j < n; ++j) {
        L acc = 0;
        for (int i = 0; i < n; ++i) acc += f[i] * cosl(pi * j * (2 * i + 1) / (2 * n));
        c[j] = acc * 2 / n;
    }
    c[0] /= 2;
    // Chebyshev series in tau = (x - mid)/half -> monomials in tau -> monomials in (x - mid)
    T0[0] = 1;
    T1[1] = 1;
    for (int j = 0; j < n; ++j) mono[j] = c[0] * T0[j] + c[1] * T1[j];
    for (int d = 2; d < n; ++d) {
        L T2[16];
        for (int j = 0; j < n; ++j) T2[j] = (j > 0 ? 2 * T1[j - 1] : 0) - T0[j];
        for (int j = 0; j < n; ++j) {
            mono[j] += c[d] * T2[j];
            T0[j] = T1[j];
            T1[j] = T2[j];
        }
    }
    L scale = 1;
    for (int j = 0; j < n; ++j) {
        mono[j] /= scale;
        scale *= half;
    }
    *mid_out = mid;
}

struct MuFunction {   // G(x) = D B sin(C atan(sqrt(x-1))) / sqrt(x-1) in long double
    double B, C, D;
    long double operator()(long double x) const
    {
        typedef long double L;
        const L u = x - 1.0L;
        if (u <= 0.0L) return (L)D * (L)B * (L)C;
        const L r = sqrtl(u);
        return (L)D * (L)B * sinl((L)C * atanl(r)) / r;
    }
};

// Host: fills table[kMuTableDoubles] for one tyre; returns the measured max relative error of the device
// evaluation scheme against the long-double reference.
inline double build_mu_table(double B, double C, double D, double *table)
{
    typedef long double L;
    const MuFunction G{B, C, D};
    double worst = 0.0;
    for (int i = 0; i < kMuTableDoubles; ++i) table[i] = 0.0;
    for (int k = 0; k < kMuIntervals; ++k) {
        const int e = k / kMuPerBinade, m = k % kMuPerBinade;
        const L lo = ldexpl(1.0L + (L)m / kMuPerBinade, e), hi = ldexpl(1.0L + (L)(m + 1) / kMuPerBinade, e);
        L mono[16], mid;
        mu_fit_interval(G, lo, hi, kMuCoef, mono, &mid);
        double *row = table + mu_row_of(k) * kMuStride;
        MuRow *mr = reinterpret_cast<MuRow *>(row);
        mr->c3 = (double)mono[3];
        mr->c2 = (double)mono[2];
        mr->c1 = (double)mono[1];
        mr->c0 = (double)mono[0];
#if B200MP_MU_DEG5
        mr->c5 = (double)mono[5];
        mr->c4 = (double)mono[4];
#else
        mr->c6 = (float)mono[6];
        mr->c5 = (float)mono[5];
        mr->c4 = (float)mono[4];
        mr->pad = 0.0f;
#endif
        for (int i = 0; i <= 32; ++i) {                   // audit: the device scheme vs the long-double function
            const double x = (double)(lo + (hi - lo) * i / 32.0L * 0.999999L);
            const double g = mu_table_eval(row, x - (double)mid);
            const L ref = G((L)x);
            const double err = (double)fabsl(((L)g - ref) / ref);
            if (err > worst) worst = err;
        }
    }
    return worst;
}

// ---- FP32 twin of the table (K1f): the same function of x = 1 + (B s)^2 evaluated in single precision, 32 intervals
// per binade (five leading mantissa bits of the float), one cubic per interval = 4 floats = ONE 128-bit load; 448 rows
// = 7 KB.  Cubic truncation is ~1e-8 of G, below the FP32 rounding of the evaluation itself (audited <= 4e-7 against
// the long-double function, i.e. a few FP32 ulps, like the closed-form FP32 routines it replaces).
constexpr int kMuBitsF = 5;
constexpr int kMuPerBinadeF = 1 << kMuBitsF;
constexpr int kMuIntervalsF = kMuPerBinadeF * kMuBinades;   // 448
// Mask-indexed rows for the FP32 table (see B200MP_MU_MASKIDX): row = [low 4 exponent bits | 5 leading mantissa bits] of the
// float x, byte offset = that field shifted into place -- one LOP3 and one shift instead of subtract / shift / compare /
// select / scale; 512 rows = 8 KB.  K1f is issue-bound (every instruction costs a slot), so unlike K1 it gains from it.
#ifndef B200MP_MU_MASKIDX_F32
#define B200MP_MU_MASKIDX_F32 1
#endif
constexpr bool kMuMaskIdxF = B200MP_MU_MASKIDX_F32 != 0;
constexpr int kMuRowsF = kMuMaskIdxF ? (16 << kMuBitsF) : kMuIntervalsF;
constexpr int kMuTableFloats = kMuRowsF * 4;
constexpr int kMuKeyShiftF = 23 - kMuBitsF;
constexpr unsigned kMuKeyMaskF = (unsigned)((16 << kMuBitsF) - 1) << kMuKeyShiftF;
inline constexpr int mu_row_of_f(int k)
{
    return kMuMaskIdxF ? (((((k >> kMuBitsF) + 15) & 15) << kMuBitsF) | (k & (kMuPerBinadeF - 1))) : k;
}
struct alignas(16) MuRowF {
    float c3, c2, c1, c0;
};
B200MP_HD float mu_row_eval(const MuRowF &r, float t) { return fmaf(fmaf(fmaf(r.c3, t, r.c2), t, r.c1), t, r.c0); }
// k: row index, or (MASKIDX_F32) the key returned by MuTab<float>::locate (row << kMuKeyShiftF; a row is 16 bytes)
B200MP_HD MuRowF mu_row_load(const float *table, int k)
{
    MuRowF r;
    if (kMuMaskIdxF) k = (int)((unsigned)k >> kMuKeyShiftF);
#if defined(__CUDA_ARCH__)
    const float4 v = reinterpret_cast<const float4 *>(table)[k];
    r.c3 = v.x; r.c2 = v.y; r.c1 = v.z; r.c0 = v.w;
#else
    r = reinterpret_cast<const MuRowF *>(table)[k];
#endif
    return r;
}
inline double build_mu_table_f32(double B, double C, double D, float *table)
{
    typedef long double L;
    const MuFunction G{B, C, D};
    double worst = 0.0;
    for (int i = 0; i < kMuTableFloats; ++i) table[i] = 0.0f;
    for (int k = 0; k < kMuIntervalsF; ++k) {
        const int e = k / kMuPerBinadeF, m = k % kMuPerBinadeF;
        const L lo = ldexpl(1.0L + (L)m / kMuPerBinadeF, e), hi = ldexpl(1.0L + (L)(m + 1) / kMuPerBinadeF, e);
        L mono[16], mid;
        mu_fit_interval(G, lo, hi, 4, mono, &mid);
        MuRowF *mr = reinterpret_cast<MuRowF *>(table) + mu_row_of_f(k);
        mr->c3 = (float)mono[3];
        mr->c2 = (float)mono[2];
        mr->c1 = (float)mono[1];
        mr->c0 = (float)mono[0];
        for (int i = 0; i <= 32; ++i) {
            const float x = (float)(lo + (hi - lo) * i / 32.0L * 0.999999L);
            if ((L)x < lo || (L)x >= hi) continue;
            const float g = mu_row_eval(*mr, x - (float)mid);
            const L ref = G((L)x);
            const double err = (double)fabsl(((L)g - ref) / ref);
            if (err > worst) worst = err;
        }
    }
    return worst;
}

// Per-precision view of the table machinery used by wheel_forces<.., TAB = true>.
template <typename R> struct MuTab;
template <> struct MuTab<double> {
    typedef MuRow Row;
    static constexpr int kIntervals = kMuIntervals;
    static constexpr int kBytes = kMuTableDoubles * 8;
    // row key (a valid row whatever x is), t = x - midpoint and "x is inside the table" from the bit pattern of
    // x = 1 + B^2 q >= 1 (or NaN)
    static B200MP_HD int locate(double q, double B2, double *t, bool *inside)
    {
        const double x = fma(q, B2, 1.0);
        const int hi = Math<double>::hi_word(x);
        const int keep = (int)(0xFFFFFFFFu << (20 - kMuBits));
        *t = x - Math<double>::from_words((hi & keep) | (1 << (19 - kMuBits)), 0);
        if (kMuMaskIdx) {
            *inside = (unsigned)hi < 0x3FF00000u + ((unsigned)kMuBinades << 20);   // x < 2^14; NaN and negative fail
            return (int)((unsigned)hi & kMuKeyMask);
        }
        const int kraw = (hi - 0x3FF00000) >> (20 - kMuBits);                      // exponent + leading mantissa bits
        *inside = (unsigned)kraw < (unsigned)kIntervals;
        return *inside ? kraw : 0;
    }
    static B200MP_HD Row load(const void *table, int k) { return mu_row_load(static_cast<const double *>(table), k); }
};
template <> struct MuTab<float> {
    typedef MuRowF Row;
    static constexpr int kIntervals = kMuIntervalsF;
    static constexpr int kBytes = kMuTableFloats * 4;
    static B200MP_HD int locate(float q, float B2, float *t, bool *inside)
    {
        const float x = fmaf(q, B2, 1.0f);
        const int b = Math<float>::bits(x);
        const int keep = (int)(0xFFFFFFFFu << (23 - kMuBitsF));
        *t = x - Math<float>::from_bits((b & keep) | (1 << (22 - kMuBitsF)));
        if (kMuMaskIdxF) {
            *inside = (unsigned)b < 0x3F800000u + ((unsigned)kMuBinades << 23);   // x < 2^14; NaN and negative fail
            return (int)((unsigned)b & kMuKeyMaskF);
        }
        const int kraw = (b - 0x3F800000) >> (23 - kMuBitsF);
        *inside = (unsigned)kraw < (unsigned)kIntervals;
        return *inside ? kraw : 0;
    }
    static B200MP_HD Row load(const void *table, int k) { return mu_row_load(static_cast<const float *>(table), k); }
};

// Controls of one zero-order-hold segment, with the steer trigonometry already evaluated.
template <typename R> struct WheelCtrl {
    R cd[4], sd[4];
    R tq[4];   // wheel torque / Jw (set_torque)
};

template <typename R>
B200MP_HD void set_torque(WheelCtrl<R> &c, const DevParams<R> &P, const R tau[4])
{
#pragma unroll
    for (int i = 0; i < 4; ++i) c.tq[i] = tau[i] * P.inv_Jw;
}

template <typename R, bool REAR0>
B200MP_HD void set_steer(WheelCtrl<R> &c, const R delta[4])
{
    Math<R>::sincos(delta[0], &c.sd[0], &c.cd[0]);
    if (REAR0) {
        // front-steer layout: FL = FR, rear wheels straight (drive.py:143)
        c.sd[1] = c.sd[0];
        c.cd[1] = c.cd[0];
        c.sd[2] = c.sd[3] = (R)0;
        c.cd[2] = c.cd[3] = (R)1;
    } else {
        for (int i = 1; i < 4; ++i) Math<R>::sincos(delta[i], &c.sd[i], &c.cd[i]);
    }
}

template <typename R>
B200MP_HD void normal_loads(const DevParams<R> &P, R ax_prev, R ay_prev, R Fz[4])
{
    Fz[0] = P.Fz0F - P.DfzxL * ax_prev - P.DfzyF * ay_prev;   // :255-258
    Fz[1] = P.Fz0F - P.DfzxR * ax_prev + P.DfzyF * ay_prev;
    Fz[2] = P.Fz0R + P.DfzxL * ax_prev - P.DfzyR * ay_prev;
    Fz[3] = P.Fz0R + P.DfzxR * ax_prev + P.DfzyR * ay_prev;
}

// Tyre of wheel I: slips -> combined-slip Pacejka friction -> forces in the chassis frame.
// TY1: all four tyres share one (B, C) pair -- read entry 0 so the kernel carries 2 constants, not 8.
// Longitudinal speed of wheel I in its own frame (:274-281); the same expression as in wheel_forces, so the two merge.
template <typename R, int I, bool REAR0>
B200MP_HD R wheel_vx(R vxc, R vyc, R cd, R sd)
{
    return (REAR0 && I >= 2) ? vxc : vxc * cd + vyc * sd;
}

// The four reciprocals 1/vx of a stage from ONE hardware reciprocal: 1/(v0 v1 v2 v3) and nine multiplies (a two-level
// product tree) instead of four MUFU seeds with their zero-low-word moves and four Newton corrections; the FP64
// instruction count is the same (12), 3 MUFU + 6 MOV fewer per stage.  Each result carries three more roundings
// (~2 ulp).  A zero or non-finite vx makes all four NaN/Inf -- as it makes the whole state after the reference's step.
#ifndef B200MP_BATCH_RCP
#define B200MP_BATCH_RCP 0
#endif
#ifndef B200MP_HEADING_FRAME
#define B200MP_HEADING_FRAME 0
#endif
// B200MP_CARRY_RCP (FP64 tabulated step): stages 2-4 refine the previous stage's 1/vx of the same wheel instead of taking a
// new hardware seed -- vx moves by ~1e-5 relative between stages, so r (1 + e + e^2 + e^3) with e = 1 - vx r is exact to e^4
// (|e| <= 2^-13 checked, else the step is repeated on the checked path).  12 of the 16 MUFU.RCP64H of a step and their
// zero-low-word moves go, the reciprocal's part of the stage's dependency chain shrinks from seed + 3 DFMA to 4 DFMA; costs
// eight live registers across the stages.  Measured: see profiles/r02_k1_launch_shape.md.
#ifndef B200MP_CARRY_RCP
#define B200MP_CARRY_RCP 0
#endif
#ifndef B200MP_WHEEL_ORDER
#define B200MP_WHEEL_ORDER 0
#endif
#ifndef B200MP_HEADING_FRAME_F32
#define B200MP_HEADING_FRAME_F32 1   /* K1f is issue-bound: there the 12 instructions per step are time */
#endif
template <typename R>
B200MP_HD void rcp4(const R v[4], R r[4])
{
    typedef Math<R> M;
    const R p01 = v[0] * v[1], p23 = v[2] * v[3];
    const R inv = M::rcp(p01 * p23);
    const R r01 = inv * p23, r23 = inv * p01;
    r[0] = r01 * v[1];
    r[1] = r01 * v[0];
    r[2] = r23 * v[3];
    r[3] = r23 * v[2];
}

template <typename R, int I, bool REAR0, bool TY1, bool TAB, bool FIRST = true, bool EXT_R = false>
B200MP_HD void wheel_forces(const DevParams<R> &P, R D, R vxc, R vyc, R w, R cd, R sd, R Fz,
                            R &fx, R &fy, R &fxt, R &fyt, R &s, const MuTableView &T, bool &ok, MuRowCacheT<typename MuTab<R>::Row> *RC = nullptr,
                            R r_ext = (R)0)
{
    constexpr int J = TY1 ? 0 : I;
    typedef Math<R> M;
    R vx, vy;
    if (REAR0 && I >= 2) {
        vx = vxc;
        vy = vyc;
    } else {
        vx = vxc * cd + vyc * sd;        // :274-281
        vy = vyc * cd - vxc * sd;
    }
    R r;
    constexpr bool kCarry = B200MP_CARRY_RCP != 0 && TAB && sizeof(R) == 8 && !EXT_R;
    if (kCarry && RC && !FIRST) {
        const R y0 = (R)RC->rinv[I];
        const R e = fma(-vx, y0, (R)1);
        const R t1 = fma(e, e, e);
        const R t2 = fma(t1, e, e);
        r = fma(y0, t2, y0);
        ok &= M::abs(e) <= (R)0.0001220703125;
    } else {
        r = EXT_R ? r_ext : M::rcp(vx);
    }
    if (kCarry && RC) RC->rinv[I] = (double)r;
    // (rw*w - vx)/vx instead of rw*w/vx - 1 (:284-287): same value, no cancellation after the divide,
    // and exactly 0 when the rounded product equals vx (the reference's zero-slip equilibrium)
    const R sx = (M::mul_rn(P.rw, w) - vx) * r;
    const R sy = -vy * M::abs(r);        // :290-293
    const R q = sx * sx + sy * sy;       // :296-299
    R gF;
    if (TAB) {
        // tabulated G(x) = D B sin(C atan(B s)) / (B s), x = 1 + (B s)^2: no sqrt, no reciprocal, no atan, no sin
        typedef MuTab<R> MT;
        R t;
        bool inside;
        const int k = MT::locate(q, (R)T.B2, &t, &inside);   // always a valid row; NaN / Inf / beyond the table: the step is repeated exactly
        ok &= inside;
        R g;
        if (RC) {
            if (FIRST || k != RC->k[I]) {                 // stage 1: always; later stages: only lanes that changed interval
                RC->r[I] = MT::load(T.c, k);
                RC->k[I] = k;
            }
            g = mu_row_eval(RC->r[I], t);
        } else {
            g = mu_row_eval(MT::load(T.c, k), t);
        }
        s = (R)0;                        // the combined slip itself is not formed on this path (logging uses the other)
        gF = g * Fz;
    } else {
        (void)J;
        const bool slipping = q != (R)0;
        const R rs = M::rsqrt(q);
        s = slipping ? q * rs : (R)0;    // sqrt(q); rsqrt(0) = inf must not leak into s
        const R mu = D * M::sin(P.Cc[J] * M::atan(P.Bc[J] * s));   // :303-306
        const R g = slipping ? mu * rs : (R)0;                     // :309-348 (zero slip -> zero friction)
        gF = g * Fz;                     // :351-360  mu_x*Fz = sx*(mu/s)*Fz
    }
    fxt = sx * gF;
    fyt = sy * gF;
    if (REAR0 && I >= 2) {
        fx = fxt;
        fy = fyt;
    } else {
        fx = fxt * cd - fyt * sd;        // :363-373
        fy = fxt * sd + fyt * cd;
    }
}

// y8 = [U V wz wFL wFR wRL wRR yaw]; sy, cy = sin / cos of the heading y8[7] (supplied by the caller,
// which gets the four stage headings of a step from one sincos plus small-angle rotations).
// k[10] receives the derivative of the full 10-state (k[7] = wz, k[8] = x_dot, k[9] = y_dot).
// out (AUX only) = the reference's 18 "outputs".
template <typename R, bool REAR0, bool AUX, bool TY1, bool TAB = false, bool FIRST = true>
B200MP_HD void planar_rhs(const DevParams<R> &P, const R D[4], const R y8[8], R sy, R cy, const WheelCtrl<R> &c,
                          const R Fz[4], R k[10], R &axc, R &ayc, R *out, const MuTableView &T = MuTableView(),
                          bool *okp = nullptr, MuRowCacheT<typename MuTab<R>::Row> *RC = nullptr)
{
    bool ok = true;
    const R U = y8[0], V = y8[1], wz = y8[2];
    const R hw = P.halfT * wz;                       // :261-271
    const R vxL = U - hw, vxR = U + hw;
    const R vyF = V + P.a * wz, vyR = V - P.b * wz;
    R fx[4], fy[4], fxt[4], fyt[4], s[4];
    // the tabulated FP64 step shares one hardware reciprocal between the four wheels (rcp4)
    constexpr bool kBatch = B200MP_BATCH_RCP != 0 && TAB && sizeof(R) == 8;
    R ri[4] = {(R)0, (R)0, (R)0, (R)0};
    if (kBatch) {
        const R v4[4] = {wheel_vx<R, 0, REAR0>(vxL, vyF, c.cd[0], c.sd[0]), wheel_vx<R, 1, REAR0>(vxR, vyF, c.cd[1], c.sd[1]),
                         wheel_vx<R, 2, REAR0>(vxL, vyR, c.cd[2], c.sd[2]), wheel_vx<R, 3, REAR0>(vxR, vyR, c.cd[3], c.sd[3])};
        rcp4(v4, ri);
    }
#define B200MP_WHEEL0 wheel_forces<R, 0, REAR0, TY1, TAB, FIRST, kBatch>(P, D[TY1 ? 0 : 0], vxL, vyF, y8[3], c.cd[0], c.sd[0], Fz[0], fx[0], fy[0], fxt[0], fyt[0], s[0], T, ok, RC, ri[0]);
#define B200MP_WHEEL1 wheel_forces<R, 1, REAR0, TY1, TAB, FIRST, kBatch>(P, D[TY1 ? 0 : 1], vxR, vyF, y8[4], c.cd[1], c.sd[1], Fz[1], fx[1], fy[1], fxt[1], fyt[1], s[1], T, ok, RC, ri[1]);
#define B200MP_WHEEL2 wheel_forces<R, 2, REAR0, TY1, TAB, FIRST, kBatch>(P, D[TY1 ? 0 : 2], vxL, vyR, y8[5], c.cd[2], c.sd[2], Fz[2], fx[2], fy[2], fxt[2], fyt[2], s[2], T, ok, RC, ri[2]);
#define B200MP_WHEEL3 wheel_forces<R, 3, REAR0, TY1, TAB, FIRST, kBatch>(P, D[TY1 ? 0 : 3], vxR, vyR, y8[6], c.cd[3], c.sd[3], Fz[3], fx[3], fy[3], fxt[3], fyt[3], s[3], T, ok, RC, ri[3]);
    // The four wheels are independent; the order they are written in only seeds ptxas's list scheduler (development A/B,
    // profiles/r02_k1_launch_shape.md: the step's time moves by a few per cent with it, the results do not move at all).
#if B200MP_WHEEL_ORDER == 1
    B200MP_WHEEL2 B200MP_WHEEL3 B200MP_WHEEL0 B200MP_WHEEL1
#elif B200MP_WHEEL_ORDER == 2
    B200MP_WHEEL0 B200MP_WHEEL2 B200MP_WHEEL1 B200MP_WHEEL3
#elif B200MP_WHEEL_ORDER == 3
    B200MP_WHEEL3 B200MP_WHEEL2 B200MP_WHEEL1 B200MP_WHEEL0
#elif B200MP_WHEEL_ORDER == 4
    B200MP_WHEEL2 B200MP_WHEEL0 B200MP_WHEEL3 B200MP_WHEEL1
#elif B200MP_WHEEL_ORDER == 5
    B200MP_WHEEL1 B200MP_WHEEL0 B200MP_WHEEL3 B200MP_WHEEL2
#else
    B200MP_WHEEL0 B200MP_WHEEL1 B200MP_WHEEL2 B200MP_WHEEL3
#endif
#undef B200MP_WHEEL0
#undef B200MP_WHEEL1
#undef B200MP_WHEEL2
#undef B200MP_WHEEL3

    const R Vwz = V * wz, Uwz = U * wz;
    // pair sums shared between the force balance and the yaw moment (the reference adds left to right, :376-378;
    // the difference is one rounding)
    const R fyF = fy[0] + fy[1], fyR = fy[2] + fy[3];
    const R fxL = fx[0] + fx[2], fxR = fx[1] + fx[3];
    const R U_dot = P.inv_m * (fxL + fxR) + Vwz;                          // :376-378
    const R V_dot = P.inv_m * (fyF + fyR) - Uwz;
    k[0] = U_dot;
    k[1] = V_dot;
    k[2] = P.inv_Izz * (P.a * fyF - P.b * fyR + P.halfT * (fxR - fxL));
    k[3] = c.tq[0] - P.rw_inv_Jw * fxt[0];                                // :379-382, (tq - rw f)/Jw
    k[4] = c.tq[1] - P.rw_inv_Jw * fxt[1];
    k[5] = c.tq[2] - P.rw_inv_Jw * fx[2];     // chassis-frame force on the rear axle, as the reference
    k[6] = c.tq[3] - P.rw_inv_Jw * fx[3];
    k[7] = wz;                                                            // :383-385
    k[8] = U * cy - V * sy;
    k[9] = U * sy + V * cy;
    axc = U_dot - Vwz;                                                    // :413-414
    ayc = V_dot + Uwz;
    if (TAB && okp) *okp = *okp && ok;
    if (AUX) {
        for (int i = 0; i < 4; ++i) {                                     // :420-423
            out[i] = fx[i];
            out[4 + i] = fy[i];
            out[8 + i] = Fz[i];
            out[12 + i] = s[i];
        }
        out[16] = fxt[0];
        out[17] = fyt[0];
    }
}

// One classic RK4 step (:427-445).  y[10] is advanced in place; ax, ay hold ax_prev, ay_prev on entry
// and the RK4-averaged axc, ayc on exit (the next step's ax_prev, ay_prev, drive.py:141).
// With AUX, sdot[10] and outs[18] receive the RK4-weighted means the reference returns (:440-441).
//
// SPEC = true: the heading trigonometry uses the branch-free "core" forms and the step is one basic
// block; nothing is committed and false is returned when an argument left their range (the caller then
// repeats the step with SPEC = false, which branches to the library where needed).
template <typename R, bool REAR0, bool AUX, bool TY1, bool SPEC, bool TAB = false>
B200MP_HD bool rk4_step_impl(const DevParams<R> &P, const R D[4], const WheelCtrl<R> &c, R h, R y[10], R &ax, R &ay,
                             R *sdot, R *outs, const MuTableView &T = MuTableView(), MuRowCacheT<typename MuTab<R>::Row> *RCX = nullptr)
{
    static_assert(!TAB || (SPEC && !AUX), "the tabulated friction path is speculative and does not log the slip");
    typedef Math<R> M;
    R Fz[4];
    normal_loads(P, ax, ay, Fz);
    if (TAB && !TY1) {   // tables of the generic path are normalised to D = 1: mu Fz = G(x) (D Fz), D per wheel
#pragma unroll
        for (int i = 0; i < 4; ++i) Fz[i] *= D[i];
    }
    R acc[10], ys[8], k[10], o[AUX ? 18 : 1], axc, ayc, sax, say;
    const R h2 = h * (R)0.5;
    // heading trigonometry: one sincos per step; the three later stage headings are yaw + e with
    // e = h/2*wz or h*wz (tiny), obtained by a small-angle rotation of (s0, c0).
    // FRAME (the speculative fast path): x_dot, y_dot of every stage are formed in the frame of the STEP's heading --
    // k8 = U cos e - V sin e, k9 = U sin e + V cos e -- and the RK4 sum is rotated into the world frame once at the end
    // (x, y never feed back), which spares composing (s0, c0) with (sin e, cos e) in stages 2-4 and the products of stage 1.
    constexpr bool FRAME = (B200MP_HEADING_FRAME != 0 || (B200MP_HEADING_FRAME_F32 != 0 && sizeof(R) == 4)) && SPEC && !AUX;
    R s0, c0, sj, cj;
    bool ok = true;
    if (SPEC)
        ok = M::sincos_core(y[7], &s0, &c0);
    else
        M::sincos(y[7], &s0, &c0);

    // rows cached across the stages (and, when the caller keeps the cache, across steps: RCX->k starts at -1)
    MuRowCacheT<typename MuTab<R>::Row> rc_store;
    MuRowCacheT<typename MuTab<R>::Row> *RC = (TAB && kMuCacheRows) ? (RCX ? RCX : &rc_store) : nullptr;
    const R s1 = FRAME ? (R)0 : s0, c1 = FRAME ? (R)1 : c0;
    if (RCX)
        planar_rhs<R, REAR0, AUX, TY1, TAB, false>(P, D, y, s1, c1, c, Fz, k, axc, ayc, o, T, &ok, RC);
    else
        planar_rhs<R, REAR0, AUX, TY1, TAB, true>(P, D, y, s1, c1, c, Fz, k, axc, ayc, o, T, &ok, RC);
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = k[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) ys[i] = y[i] + h2 * k[i];
    sax = axc;
    say = ayc;
    if (AUX)
        for (int i = 0; i < 18; ++i) outs[i] = o[i];
    if (FRAME)
        ok &= M::small_sincos_core(h2 * k[7], &sj, &cj);
    else if (SPEC)
        ok &= M::rotate_core(s0, c0, h2 * k[7], &sj, &cj);
    else if (!M::rotate_small(s0, c0, h2 * k[7], &sj, &cj))
        M::sincos(ys[7], &sj, &cj);

    planar_rhs<R, REAR0, AUX, TY1, TAB, false>(P, D, ys, sj, cj, c, Fz, k, axc, ayc, o, T, &ok, RC);
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] += (R)2 * k[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) ys[i] = y[i] + h2 * k[i];
    sax += (R)2 * axc;
    say += (R)2 * ayc;
    if (AUX)
        for (int i = 0; i < 18; ++i) outs[i] += (R)2 * o[i];
    if (FRAME)
        ok &= M::small_sincos_core(h2 * k[7], &sj, &cj);
    else if (SPEC)
        ok &= M::rotate_core(s0, c0, h2 * k[7], &sj, &cj);
    else if (!M::rotate_small(s0, c0, h2 * k[7], &sj, &cj))
        M::sincos(ys[7], &sj, &cj);

    planar_rhs<R, REAR0, AUX, TY1, TAB, false>(P, D, ys, sj, cj, c, Fz, k, axc, ayc, o, T, &ok, RC);
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] += (R)2 * k[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) ys[i] = y[i] + h * k[i];
    sax += (R)2 * axc;
    say += (R)2 * ayc;
    if (AUX)
        for (int i = 0; i < 18; ++i) outs[i] += (R)2 * o[i];
    if (FRAME)
        ok &= M::small_sincos_core(h * k[7], &sj, &cj);
    else if (SPEC)
        ok &= M::rotate_core(s0, c0, h * k[7], &sj, &cj);
    else if (!M::rotate_small(s0, c0, h * k[7], &sj, &cj))
        M::sincos(ys[7], &sj, &cj);

    planar_rhs<R, REAR0, AUX, TY1, TAB, false>(P, D, ys, sj, cj, c, Fz, k, axc, ayc, o, T, &ok, RC);
    if (SPEC && !ok) return false;
    const R h6 = (R)(1.0 / 6) * h;       // :438  state + 1/6*h*(K1+2K2+2K3+K4)
    const R sixth = (R)(1.0 / 6);
#pragma unroll
    for (int i = 0; i < (FRAME ? 8 : 10); ++i) {
        const R sum = acc[i] + k[i];
        y[i] = y[i] + h6 * sum;
        if (AUX) sdot[i] = sum * sixth;
    }
    if (FRAME) {   // the RK4 sum of (x_dot, y_dot) in the step's heading frame -> world frame
        const R sa = acc[8] + k[8], sb = acc[9] + k[9];
        y[8] = y[8] + h6 * (c0 * sa - s0 * sb);
        y[9] = y[9] + h6 * (s0 * sa + c0 * sb);
    }
    ax = (sax + axc) * sixth;            // :442-443
    ay = (say + ayc) * sixth;
    if (AUX)
        for (int i = 0; i < 18; ++i) outs[i] = (outs[i] + o[i]) * sixth;
    return true;
}

#if defined(__CUDACC__)
#define B200MP_NOINLINE __noinline__
#else
#define B200MP_NOINLINE
#endif

// the branching version, out of line: reached only when a heading or a stage rotation is out of range
template <typename R, bool REAR0, bool AUX, bool TY1>
#if defined(__CUDACC__)
__host__ __device__ B200MP_NOINLINE
#endif
void rk4_step_checked(const DevParams<R> &P, const R *D, const WheelCtrl<R> &c, R h, R *y, R *axay, R *sdot, R *outs)
{
    rk4_step_impl<R, REAR0, AUX, TY1, false>(P, D, c, h, y, axay[0], axay[1], sdot, outs);
}

// SPEC: run the straight-line speculative form first (used by the register-lean fast-path kernels; the
// generic / logging instantiations are already at the register ceiling and keep the branching form).
template <typename R, bool REAR0, bool AUX, bool TY1, bool SPEC, bool TAB = false>
B200MP_HD void rk4_step(const DevParams<R> &P, const R D[4], const WheelCtrl<R> &c, R h, R y[10], R &ax, R &ay,
                        R *sdot, R *outs, const MuTableView &T = MuTableView(), MuRowCacheT<typename MuTab<R>::Row> *RCX = nullptr)
{
    if (SPEC) {
        if (!rk4_step_impl<R, REAR0, AUX, TY1, true, TAB>(P, D, c, h, y, ax, ay, sdot, outs, T, RCX)) {
            // the out-of-line call takes addresses: hand it copies, so that the caller's y / c / D stay in
            // registers on the hot path instead of living in local memory across the step loop
            R axay[2] = {ax, ay}, yt[10], Dt[4];
            WheelCtrl<R> ct = c;
            const DevParams<R> Pt = P;
#pragma unroll
            for (int i = 0; i < 10; ++i) yt[i] = y[i];
#pragma unroll
            for (int i = 0; i < 4; ++i) Dt[i] = D[i];
            rk4_step_checked<R, REAR0, AUX, TY1>(Pt, Dt, ct, h, yt, axay, sdot, outs);
#pragma unroll
            for (int i = 0; i < 10; ++i) y[i] = yt[i];
            ax = axay[0];
            ay = axay[1];
        }
    } else {
        rk4_step_impl<R, REAR0, AUX, TY1, false>(P, D, c, h, y, ax, ay, sdot, outs);
    }
}

}  // namespace b200mp
