// Microbenchmark (development tool): sustained issue interval of FP64 instructions per scheduler as a function of how many
// DISTINCT register operands they read.  16 independent chains per thread, 1 or 2 warps per scheduler.
//   P0: x = fma(x, a, b)    a, b loop-invariant (operand reuse)      P1: x = fma(x, y_i, b)     P2: x = fma(x, y_i, z_i)
//   P3: x = fma(x, y_i, c[const bank])                                P4: x = x * y_i            P5: x = x + y_i
//   P6: x = fma(x, c[const], c[const])
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double kc[16];
template <int P> __global__ void ops(double *out, int iters, double a, double b)
{
    double x[16], y[16], z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] = threadIdx.x * 1e-3 + i; y[i] = 1.0 + 1e-9 * (threadIdx.x + i); z[i] = 1e-9 * (i + 1); }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (P == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(a), "d"(b));
                if (P == 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(y[i]), "d"(b));
                if (P == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(y[i]), "d"(z[i]));
                if (P == 3) x[i] = fma(x[i], y[i], kc[i]);
                if (P == 4) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x[i]) : "d"(y[i]));
                if (P == 5) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[i]) : "d"(y[i]));
                if (P == 6) x[i] = fma(x[i], kc[i], kc[(i + 1) & 15]);
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i] + y[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}
template <int P> void run(int w)
{
    static const char *names[] = {"fma(x, a, b)  shared a, b", "fma(x, y_i, b)", "fma(x, y_i, z_i)", "fma(x, y_i, c[bank])", "mul(x, y_i)", "add(x, y_i)", "fma(x, c[bank], c[bank])"};
    double *d; cudaMalloc(&d, 1 << 22);
    const int iters = 2000;
    for (int k = 0; k < 2; ++k) ops<P><<<148, 128 * w>>>(d, iters, 1.0000001, 1e-9);
    double cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %d warp(s)/scheduler: %5.2f cycles per warp-instruction per scheduler\n", names[P], w, cyc / ((double)iters * 64 * w));
    cudaFree(d);
}
int main()
{
    double h[16]; for (int i = 0; i < 16; ++i) h[i] = 1e-9 * (i + 1);
    cudaMemcpyToSymbol(kc, h, sizeof(h));
    for (int w : {1, 2}) { run<0>(w); run<1>(w); run<2>(w); run<3>(w); run<4>(w); run<5>(w); run<6>(w); }
    return 0;
}
