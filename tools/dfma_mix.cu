// Microbenchmark (development tool): does a non-FP64 instruction issue in the shadow of a DFMA (2 cycles per warp on the
// 16-lane FP64 pipe) or does it cost its own issue cycle?  ILP-8 DFMA chains mixed with K independent FFMA / IMAD / LDS.
#include <cstdio>
#include <cuda_runtime.h>
template <int KF, int KI, int KL> __global__ void mix(double *out, int iters, double a, double b, float fa, int ia)
{
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i;
    __syncthreads();
    double x[8];
    float f[8];
    int n[8];
    float l = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3 + i; f[i] = threadIdx.x + i; n[i] = threadIdx.x + i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                x[i] = fma(x[i], a, b);
                if (i < KF) f[i] = fmaf(f[i], fa, 1.0f);
                if (i >= 4 && i - 4 < KF - 8 + 4 && KF > 8) f[i] = fmaf(f[i], fa, 2.0f);
                if (i < KI) n[i] = n[i] * ia + 7;
                if (i < KL) l += sm[(n[0] + i * 32 + threadIdx.x) & 1023];
            }
    }
    long long t1 = clock64();
    double s = l;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + f[i] + n[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}
template <int KF, int KI, int KL> void run(int w)
{
    double *d; cudaMalloc(&d, 1 << 20);
    const int iters = 2000;
    for (int k = 0; k < 2; ++k) mix<KF, KI, KL><<<148, 128 * w>>>(d, iters, 1.0000001, 1e-9, 1.0001f, 3);
    double cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    const double n = (double)iters * 32;
    printf("per 8 DFMA: +%d FFMA +%d IMAD +%d LDS, warps/SMSP %d: %.2f cycles per DFMA per SMSP\n", KF, KI, KL, w, cyc / (n * w));
    cudaFree(d);
}
int main()
{
    for (int w : {1, 2}) {
        run<0, 0, 0>(w); run<2, 0, 0>(w); run<4, 0, 0>(w); run<8, 0, 0>(w);
        run<0, 2, 0>(w); run<0, 4, 0>(w); run<0, 8, 0>(w); run<4, 4, 0>(w); run<8, 8, 0>(w);
        run<0, 0, 1>(w); run<0, 0, 2>(w); run<0, 0, 4>(w);
    }
    return 0;
}
