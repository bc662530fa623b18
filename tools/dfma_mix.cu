// Microbenchmark (development tool): does a non-FP64 instruction issue in the shadow of a DFMA (a warp's DFMA holds the
// 16-lane FP64 pipe of its scheduler for 2 cycles) or does it cost its own issue cycle?  Eight independent DFMA chains
// with K independent FFMA (or IMAD) chains interleaved, straight-line, one or two warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
template <int K, bool INT> __global__ void mix(double *out, int iters, double a, double b, float fa, int ia)
{
    double x[8];
    float f[16];
    int n[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 16; ++i) { f[i] = threadIdx.x + i; n[i] = threadIdx.x + i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(a), "d"(b));
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int c = i * 2 + j;
                    if (c < K) {
                        if (INT) asm volatile("mad.lo.s32 %0, %0, %1, 7;" : "+r"(n[c]) : "r"(ia));
                        else asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(f[c]) : "f"(fa));
                    }
                }
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f[i] + n[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}
template <int K, bool INT> void run(int w)
{
    double *d; cudaMalloc(&d, 1 << 20);
    const int iters = 2000;
    for (int k = 0; k < 2; ++k) mix<K, INT><<<148, 128 * w>>>(d, iters, 1.0000001, 1e-9, 1.0001f, 3);
    double cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    const double groups = (double)iters * 4 * w;   // groups of (8 DFMA + K others) issued per scheduler
    printf("8 DFMA + %2d %s, %d warp(s)/scheduler: %6.2f cycles per group  (co-issue would give %4.1f, separate issue cycles %4.1f)\n", K,
           INT ? "IMAD" : "FFMA", w, cyc / groups, K > 8 ? 8.0 + K : 16.0, 16.0 + K);
    cudaFree(d);
}
int main()
{
    for (int w : {1, 2, 4}) {
        run<0, false>(w); run<2, false>(w); run<4, false>(w); run<8, false>(w); run<16, false>(w);
        run<4, true>(w); run<8, true>(w); run<16, true>(w);
    }
    return 0;
}
