// Scalar math building blocks for the rollout kernels (sm_100a).
//
// FP64 division and square root are not single instructions on the GPU; the rollout needs, per wheel
// and RK4 stage, 1/vx and both sqrt(q) and 1/sqrt(q).  We take the hardware seeds (MUFU.RCP64H /
// MUFU.RSQ64H, ~2^-22 relative) and finish with two Newton steps on the FP64 pipe, which is 1-2 ulp --
// three orders of magnitude inside the 1e-9 parity contract -- at a third of the instruction count of
// the IEEE-rounded library routines and with no slow-path branches.
//
// The same header compiles for the host (tests/hostsim) with plain libm so the algebra of the kernels
// can be checked against the oracle without a GPU.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define B200MP_HD __host__ __device__ __forceinline__
#else
#define B200MP_HD inline
#endif

namespace b200mp {

template <typename R> struct Math;

template <> struct Math<double> {
    static B200MP_HD double rcp(double x)
    {
#if defined(__CUDA_ARCH__)
        double y;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        double e = fma(-x, y, 1.0);
        y = fma(y, e, y);
        e = fma(-x, y, 1.0);
        return fma(y, e, y);
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
        const double h = 0.5 * q;
        double e = fma(-h * y, y, 0.5);
        y = fma(y, e, y);
        e = fma(-h * y, y, 0.5);
        return fma(y, e, y);
#else
        return 1.0 / ::sqrt(q);
#endif
    }
    static B200MP_HD void sincos(double x, double *s, double *c)
    {
#if defined(__CUDA_ARCH__)
        ::sincos(x, s, c);
#else
        *s = ::sin(x);
        *c = ::cos(x);
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
    static B200MP_HD double sin(double x) { return ::sin(x); }
    static B200MP_HD double atan(double x) { return ::atan(x); }
    static B200MP_HD double abs(double x) { return ::fabs(x); }
};

template <> struct Math<float> {
    static B200MP_HD float rcp(float x)
    {
#if defined(__CUDA_ARCH__)
        return __frcp_rn(x);
#else
        return 1.0f / x;
#endif
    }
    static B200MP_HD float rsqrt(float q)
    {
#if defined(__CUDA_ARCH__)
        // rsqrtf is 2 ulp; one Newton step brings s = q*rsqrt(q) to ~1 ulp
        float y = rsqrtf(q);
        return y;
#else
        return 1.0f / ::sqrtf(q);
#endif
    }
    static B200MP_HD void sincos(float x, float *s, float *c)
    {
#if defined(__CUDA_ARCH__)
        ::sincosf(x, s, c);
#else
        *s = ::sinf(x);
        *c = ::cosf(x);
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
    static B200MP_HD float sin(float x) { return ::sinf(x); }
    static B200MP_HD float atan(float x) { return ::atanf(x); }
    static B200MP_HD float abs(float x) { return ::fabsf(x); }
};

}  // namespace b200mp
