// Scalar math building blocks for the rollout kernels (sm_100a).
//
// Why not the CUDA math library: per RK4 step the model needs 16 atan, 16 sin, 8 sincos, 32 divisions and
// 16 square roots in FP64.  ncu on the first version of the kernel (profiles/r01_rollout_f64_v1.md)
// showed that with libm only 45 % of the issued instructions were FP64-pipe instructions: the rest were
// UMOV pairs materialising 64-bit polynomial coefficients, LDG loads of libm's coefficient tables,
// FSEL/branch quadrant logic and slow-path scaffolding, which capped the FP64 pipe at 50 % busy and
// overflowed the instruction cache.  The routines below
//   * keep every polynomial coefficient in __constant__ memory, so it is a constant-bank operand of the
//     DFMA itself (no UMOV, no load, no register),
//   * are branch-free on their whole working range,
//   * take the hardware seeds MUFU.RCP64H / MUFU.RSQ64H plus one third-order correction step for 1/x
//     and 1/sqrt(x) (1-2 ulp; no IEEE-rounding fix-up code),
//   * use argument ranges the model guarantees (atan of a real, sin of C*atan(.), sincos of a heading).
// Accuracy of each scheme is ~1-2 ulp (tools/gen_poly.py prints the measured bounds), three orders of
// magnitude inside the 1e-9 parity contract.
//
// The same header compiles for the host (tests/hostsim, plain g++) so the polynomials and the algebra
// of the kernels are checked against the oracle on a machine without a GPU.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define B200MP_HD __host__ __device__ __forceinline__
#define B200MP_TABLE static __constant__
#else
#define B200MP_HD inline
#define B200MP_TABLE static const
#endif

// custom polynomial paths: device code, and plain host compilers (hostsim); the host pass of nvcc
// (which cannot read __constant__ tables) uses libm and is never executed
#if defined(__CUDA_ARCH__) || !defined(__CUDACC__)
#define B200MP_POLY 1
#else
#define B200MP_POLY 0
#endif

namespace b200mp {

// ---- coefficient tables (tools/gen_poly.py; Chebyshev interpolants at 60 digits) -------------------
// sin(r) = r + r*u*P(u), u = r*r, |r| <= pi/2                      max abs err 2.2e-16
B200MP_TABLE double kSinHalfPi[8] = {-0.16666666666666666, 0.00833333333333331, -0.00019841269841251003,
                                     2.7557319218112503e-06, -2.5052107485885176e-08, 1.6058968982311192e-10,
                                     -7.643709637264542e-13, 2.728258980834901e-15};
// atan(t) = t + t*u*Q(u), u = t*t, |t| <= tan(pi/8)                max rel err 2.9e-16
B200MP_TABLE double kAtanPi8[10] = {-0.33333333333333226, 0.19999999999880183, -0.1428571426303133,
                                    0.11111109439473953, -0.09090846249880491, 0.07690942354266835,
                                    -0.06648434023802716, 0.05729171057606822, -0.04459975531678654,
                                    0.022434044895340084};
// scalar constants (uniform-register operands as well)
B200MP_TABLE double kTrig[8] = {
    0.3183098861837907,        // 0: 1/pi
    3.141592653589793,         // 1: pi (hi)
    1.2246467991473532e-16,    // 2: pi (lo)
    1.5707963267948966,        // 3: pi/2 (hi)
    6.123233995736766e-17,     // 4: pi/2 (lo)
    0.41421356237309503,       // 5: tan(pi/8)
    2.414213562373095,         // 6: tan(3 pi/8)
    0.7853981633974483,        // 7: pi/4
};
constexpr double kRoundMagic = 6755399441055744.0;   // 1.5 * 2^52: low word zero -> an immediate operand

template <typename R> struct Math;

template <> struct Math<double> {
    static B200MP_HD double rcp(double x)
    {
#if defined(__CUDA_ARCH__)
        double y;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        // one third-order step: 1/x = y (1 + e + e^2 + ...), e = 1 - x y ~ 2^-23  ->  error ~ e^3
        const double e = fma(-x, y, 1.0);
        const double t = fma(e, e, e);
        return fma(y, t, y);
#else
        return 1.0 / x;
#endif
    }
    // returns 1/sqrt(q); q == 0 gives +inf (callers guard the zero-slip case)
    static B200MP_HD double rsqrt(double q)
    {
#if defined(__CUDA_ARCH__)
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q));
        // one third-order step: q^-1/2 = y (1 - e)^-1/2 = y (1 + e/2 + 3 e^2/8 + ...), e = 1 - q y^2
        const double e = fma(-(q * y), y, 1.0);
        const double t = fma(e, 0.375, 0.5);
        return fma(y, t * e, y);
#else
        return 1.0 / ::sqrt(q);
#endif
    }
    // product rounded on its own (never contracted into a following add)
    static B200MP_HD double mul_rn(double a, double b)
    {
#if defined(__CUDA_ARCH__)
        return __dmul_rn(a, b);
#else
        return a * b;
#endif
    }
    static B200MP_HD double abs(double x) { return ::fabs(x); }

    // bit helpers
    static B200MP_HD int lo_word(double x)
    {
#if defined(__CUDA_ARCH__)
        return __double2loint(x);
#else
        uint64_t b;
        memcpy(&b, &x, 8);
        return (int)(uint32_t)b;
#endif
    }
    static B200MP_HD int hi_word(double x)
    {
#if defined(__CUDA_ARCH__)
        return __double2hiint(x);
#else
        uint64_t b;
        memcpy(&b, &x, 8);
        return (int)(uint32_t)(b >> 32);
#endif
    }
    static B200MP_HD double from_words(int hi, int lo)
    {
#if defined(__CUDA_ARCH__)
        return __hiloint2double(hi, lo);
#else
        const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
        double x;
        memcpy(&x, &b, 8);
        return x;
#endif
    }
    static B200MP_HD double xor_sign(double x, int bit0_source)
    {
        // flips the sign of x when bit 0 of bit0_source is set
#if defined(__CUDA_ARCH__)
        return __hiloint2double(__double2hiint(x) ^ (bit0_source << 31), __double2loint(x));
#else
        return (bit0_source & 1) ? -x : x;
#endif
    }

    // odd polynomial for sin on [-pi/2, pi/2]
    static B200MP_HD double sin_poly(double r)
    {
        const double u = r * r;
        double p = kSinHalfPi[7];
#pragma unroll
        for (int i = 6; i >= 0; --i) p = fma(p, u, kSinHalfPi[i]);
        return fma(r * u, p, r);
    }

    // sin(y) for moderate |y| (|y| < 2^30 pi): nearest multiple of pi removed with a two-term
    // Cody-Waite reduction, one odd polynomial on [-pi/2, pi/2], sign from the parity of the multiple.
    // The Pacejka argument C*atan(B*s) satisfies |y| < C*pi/2.
    static B200MP_HD double sin(double y)
    {
#if B200MP_POLY
        const double t = fma(y, kTrig[0], kRoundMagic);
        const double kf = t - kRoundMagic;
        double r = fma(-kf, kTrig[1], y);
        r = fma(-kf, kTrig[2], r);
        return xor_sign(sin_poly(r), lo_word(t));
#else
        return ::sin(y);
#endif
    }

    // atan(x), any finite x: three-way argument reduction
    //   |x| < tan(pi/8): t = |x|;  |x| > tan(3pi/8): t = -1/|x|, + pi/2;  else t = (|x|-1)/(|x|+1), + pi/4
    // then one odd polynomial on |t| <= tan(pi/8).  One reciprocal, no branch.
    static B200MP_HD double atan(double x)
    {
#if B200MP_POLY
        const double ax = ::fabs(x);
        const bool lo = ax < kTrig[5], hi = ax > kTrig[6];
        const double num = lo ? ax : (hi ? -1.0 : ax - 1.0);
        const double den = lo ? 1.0 : (hi ? ax : ax + 1.0);
        const double off = lo ? 0.0 : (hi ? kTrig[3] : kTrig[7]);
        const double t = num * rcp(den);
        const double u = t * t;
        double q = kAtanPi8[9];
#pragma unroll
        for (int i = 8; i >= 0; --i) q = fma(q, u, kAtanPi8[i]);
        const double res = off + fma(t * u, q, t);
        return ::copysign(res, x);
#else
        return ::atan(x);
#endif
    }

    // sin and cos of a heading, |x| <= 1e5: nearest multiple of pi removed (two-term Cody-Waite), then
    // sin(r) and cos(r) = sin(pi/2 - |r|) through the SAME polynomial as sin() above, so the per-step
    // heading trigonometry adds no coefficient table to the kernel; the parity of the multiple flips
    // both signs.  Larger or non-finite arguments fall back to the library.
    static B200MP_HD void sincos(double x, double *s, double *c)
    {
#if B200MP_POLY
        if (!(::fabs(x) <= 1.0e5)) {
#if defined(__CUDA_ARCH__)
            ::sincos(x, s, c);
#else
            *s = ::sin(x);
            *c = ::cos(x);
#endif
            return;
        }
        const double t = fma(x, kTrig[0], kRoundMagic);
        const double kf = t - kRoundMagic;
        double r = fma(-kf, kTrig[1], x);
        r = fma(-kf, kTrig[2], r);
        const double a = (kTrig[3] - ::fabs(r)) + kTrig[4];
        const int k = lo_word(t);
        *s = xor_sign(sin_poly(r), k);
        *c = xor_sign(sin_poly(a), k);
#else
        *s = ::sin(x);
        *c = ::cos(x);
#endif
    }

    // Speculative forms for straight-line code: same arithmetic without the range branch; the return value
    // says whether the argument was inside the range the fast scheme is valid for.  A caller that collects
    // the flags of a whole RK4 step and repeats the (rare) step on the branching versions keeps the four
    // stages in one basic block, which lets the scheduler overlap the serial tail of one stage with the
    // head of the next.
    static B200MP_HD bool sincos_core(double x, double *s, double *c)
    {
#if B200MP_POLY
        const double t = fma(x, kTrig[0], kRoundMagic);
        const double kf = t - kRoundMagic;
        double r = fma(-kf, kTrig[1], x);
        r = fma(-kf, kTrig[2], r);
        const double a = (kTrig[3] - ::fabs(r)) + kTrig[4];
        const int k = lo_word(t);
        *s = xor_sign(sin_poly(r), k);
        *c = xor_sign(sin_poly(a), k);
        return ::fabs(x) <= 1.0e5;
#else
        *s = ::sin(x);
        *c = ::cos(x);
        return true;
#endif
    }
    // sin e, cos e of a small increment alone (the caller works in the frame of the step's heading)
    static B200MP_HD bool small_sincos_core(double e, double *se, double *ce)
    {
        const double u = e * e;
        *se = fma(e * u, -1.0 / 6, e);
        *ce = fma(u, fma(u, 1.0 / 24, -0.5), 1.0);
        return ::fabs(e) <= 0.0009765625;
    }
    static B200MP_HD bool rotate_core(double sa, double ca, double e, double *s, double *c)
    {
        const double u = e * e;
        const double se = fma(e * u, -1.0 / 6, e);
        const double ce = fma(u, fma(u, 1.0 / 24, -0.5), 1.0);
        *s = fma(sa, ce, ca * se);
        *c = fma(ca, ce, -(sa * se));
        return ::fabs(e) <= 0.0009765625;
    }

    // (sin, cos) of (a + e) from (sin a, cos a) for a small increment e: used for the RK4 stage headings,
    // which differ from the step's heading by h/2*wz or h*wz.  sin e = e - e^3/6, cos e = 1 - e^2/2 + e^4/24
    // (|e| <= 2^-10: truncation < 8e-18); returns false when e is too large for the series (the caller
    // then evaluates sincos directly).
    static B200MP_HD bool rotate_small(double sa, double ca, double e, double *s, double *c)
    {
        if (!(::fabs(e) <= 0.0009765625)) return false;
        const double u = e * e;
        const double se = fma(e * u, -1.0 / 6, e);
        const double ce = fma(u, fma(u, 1.0 / 24, -0.5), 1.0);
        *s = fma(sa, ce, ca * se);
        *c = fma(ca, ce, -(sa * se));
        return true;
    }
};

// FP32 twin (K1f).  Same structure as the FP64 routines -- branch-free polynomials, hardware reciprocal /
// reciprocal-square-root seeds -- with single-precision targets (1-2 ulp, tools/gen_poly.py --f32):
//   * coefficients are literals, i.e. immediate operands of the FFMA;
//   * 1/x and 1/sqrt(x) are the MUFU.RCP / MUFU.RSQ results themselves (1-2 ulp);
//   * CUDA's sinf/atanf/sincosf carry range-reduction slow paths (Payne-Hanek) and quadrant selects that made
//     the FP32 kernel only 1.2x faster than the FP64 one; the model's arguments are bounded (see above).
template <> struct Math<float> {
    static B200MP_HD float rcp(float x)
    {
#if defined(__CUDA_ARCH__)
        float y;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
#else
        return 1.0f / x;
#endif
    }
    static B200MP_HD float rsqrt(float q)
    {
#if defined(__CUDA_ARCH__)
        float y;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(q));
        return y;
#else
        return 1.0f / ::sqrtf(q);
#endif
    }
    static B200MP_HD float mul_rn(float a, float b)
    {
#if defined(__CUDA_ARCH__)
        return __fmul_rn(a, b);
#else
        return a * b;
#endif
    }
    static B200MP_HD float abs(float x) { return ::fabsf(x); }
    static B200MP_HD int bits(float x)
    {
#if defined(__CUDA_ARCH__)
        return __float_as_int(x);
#else
        int b;
        memcpy(&b, &x, 4);
        return b;
#endif
    }
    static B200MP_HD float from_bits(int b)
    {
#if defined(__CUDA_ARCH__)
        return __int_as_float(b);
#else
        float x;
        memcpy(&x, &b, 4);
        return x;
#endif
    }
    // sin(r) = r + r*u*P(u), u = r*r, |r| <= pi/2 (5 coefficients, ~1 ulp of 1)
    static B200MP_HD float sin_poly(float r)
    {
        const float u = r * r;
        float p = -2.4061034054057018e-08f;
        p = fmaf(p, u, 2.753562739599147e-06f);
        p = fmaf(p, u, -0.00019841075118165463f);
        p = fmaf(p, u, 0.00833333283662796f);
        p = fmaf(p, u, -0.1666666716337204f);
        return fmaf(r * u, p, r);
    }
    // nearest multiple of pi removed (two-term Cody-Waite, exact products through the FMA); k = its parity source
    static B200MP_HD float reduce_pi(float y, int *k)
    {
        const float magic = 12582912.0f;                      // 1.5 * 2^23
        const float t = fmaf(y, 0.31830987334251404f, magic);
        const float kf = t - magic;
        float r = fmaf(-kf, 3.1415927410125732f, y);
        r = fmaf(-kf, -8.742277657347586e-08f, r);
        *k = bits(t);
        return r;
    }
    static B200MP_HD float sin(float y)
    {
#if B200MP_POLY
        int k;
        const float r = reduce_pi(y, &k);
        return from_bits(bits(sin_poly(r)) ^ (k << 31));
#else
        return ::sinf(y);
#endif
    }
    static B200MP_HD float atan(float x)
    {
#if B200MP_POLY
        const float ax = ::fabsf(x);
        const bool lo = ax < 0.4142135679721832f, hi = ax > 2.4142136573791504f;
        const float num = lo ? ax : (hi ? -1.0f : ax - 1.0f);
        const float den = lo ? 1.0f : (hi ? ax : ax + 1.0f);
        const float off = lo ? 0.0f : (hi ? 1.5707963705062866f : 0.7853981852531433f);
        const float t = num * rcp(den);
        const float u = t * t;
        float q = -0.06430986523628235f;
        q = fmaf(q, u, 0.10737381130456924f);
        q = fmaf(q, u, -0.14263364672660828f);
        q = fmaf(q, u, 0.1999952346086502f);
        q = fmaf(q, u, -0.3333333134651184f);
        const float res = off + fmaf(t * u, q, t);
        return ::copysignf(res, x);
#else
        return ::atanf(x);
#endif
    }
    static B200MP_HD bool sincos_core(float x, float *s, float *c)
    {
#if B200MP_POLY
        int k;
        const float r = reduce_pi(x, &k);
        const float a = (1.5707963705062866f - ::fabsf(r)) + -4.371138828673793e-08f;
        *s = from_bits(bits(sin_poly(r)) ^ (k << 31));
        *c = from_bits(bits(sin_poly(a)) ^ (k << 31));
        return ::fabsf(x) <= 8192.0f;                         // beyond: k*pi_lo is no longer negligible in FP32
#else
        *s = ::sinf(x);
        *c = ::cosf(x);
        return true;
#endif
    }
    static B200MP_HD void sincos(float x, float *s, float *c)
    {
        if (!sincos_core(x, s, c)) {
#if defined(__CUDA_ARCH__)
            ::sincosf(x, s, c);
#else
            *s = ::sinf(x);
            *c = ::cosf(x);
#endif
        }
    }
    static B200MP_HD bool small_sincos_core(float e, float *se, float *ce)
    {
        const float u = e * e;
        *se = fmaf(e * u, -1.0f / 6, e);
        *ce = fmaf(u, fmaf(u, 1.0f / 24, -0.5f), 1.0f);
        return ::fabsf(e) <= 0.015625f;
    }
    static B200MP_HD bool rotate_core(float sa, float ca, float e, float *s, float *c)
    {
        const float u = e * e;
        const float se = fmaf(e * u, -1.0f / 6, e);
        const float ce = fmaf(u, fmaf(u, 1.0f / 24, -0.5f), 1.0f);
        *s = fmaf(sa, ce, ca * se);
        *c = fmaf(ca, ce, -(sa * se));
        return ::fabsf(e) <= 0.015625f;
    }
    static B200MP_HD bool rotate_small(float sa, float ca, float e, float *s, float *c)
    {
        if (!(::fabsf(e) <= 0.015625f)) return false;
        rotate_core(sa, ca, e, s, c);
        return true;
    }
};

}  // namespace b200mp
