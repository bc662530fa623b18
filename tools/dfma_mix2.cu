// Microbenchmark (development tool): which non-FP64 instruction classes issue in the shadow of a DFMA?  Eight independent
// DFMA chains with K independent integer chains of one kind interleaved (ALU-pipe xor / add / shift vs FMA-pipe IMAD / FFMA).
#include <cstdio>
#include <cuda_runtime.h>
template <int K, int KIND> __global__ void mix(double *out, int iters, double a, double b, float fa, int ia)
{
    double x[8];
    int n[16];
    float f[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 16; ++i) { n[i] = threadIdx.x + i; f[i] = threadIdx.x + i; }
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
                        if (KIND == 0) asm volatile("xor.b32 %0, %0, %1;" : "+r"(n[c]) : "r"(ia));
                        if (KIND == 1) asm volatile("add.s32 %0, %0, %1;" : "+r"(n[c]) : "r"(ia));
                        if (KIND == 2) asm volatile("shf.l.wrap.b32 %0, %0, %0, %1;" : "+r"(n[c]) : "r"(ia));
                        if (KIND == 3) asm volatile("mad.lo.s32 %0, %0, %1, 7;" : "+r"(n[c]) : "r"(ia));
                        if (KIND == 4) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(f[c]) : "f"(fa));
                        if (KIND == 5) asm volatile("max.s32 %0, %0, %1;" : "+r"(n[c]) : "r"(ia));
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
    for (int i = 0; i < 16; ++i) s += n[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}
template <int K, int KIND> void run(int w)
{
    static const char *names[] = {"LOP3 (xor)", "IADD (add)", "SHF", "IMAD", "FFMA", "VIMNMX"};
    double *d; cudaMalloc(&d, 1 << 20);
    const int iters = 2000;
    for (int k = 0; k < 2; ++k) mix<K, KIND><<<148, 128 * w>>>(d, iters, 1.0000001, 1e-9, 1.0001f, 3);
    double cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    printf("8 DFMA + %2d %-11s %d warp(s)/scheduler: %6.2f cycles per group (DFMA alone 17.4; +1 cycle each would be %4.1f)\n", K, names[KIND], w,
           cyc / ((double)iters * 4 * w), 17.4 + K);
    cudaFree(d);
}
int main()
{
    for (int w : {2}) {
        run<8, 0>(w); run<16, 0>(w); run<8, 1>(w); run<16, 1>(w); run<8, 2>(w); run<16, 2>(w);
        run<8, 3>(w); run<16, 3>(w); run<8, 4>(w); run<16, 4>(w); run<8, 5>(w); run<16, 5>(w);
    }
    return 0;
}
