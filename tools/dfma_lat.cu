// Microbenchmark (development tool): DFMA issue/latency on sm_100a as a function of ILP and warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP> __global__ void chain(double *out, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}
template <int ILP> void run(int warps_per_smsp)
{
    double *d; cudaMalloc(&d, 1 << 20);
    const int iters = 2000, threads = 32 * 4 * warps_per_smsp;
    chain<ILP><<<148, threads>>>(d, iters, 1.0000001, 1e-9);
    chain<ILP><<<148, threads>>>(d, iters, 1.0000001, 1e-9);
    double cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    const double n = (double)iters * 8 * ILP;   // DFMA per warp
    printf("ILP %d warps/SMSP %d: %.2f cycles per DFMA per warp; %.2f cycles per DFMA per SMSP (pipe floor 2.0)\n", ILP, warps_per_smsp,
           cyc / n, cyc / (n * warps_per_smsp));
    cudaFree(d);
}
int main()
{
    for (int w : {1, 2, 3, 4}) { run<1>(w); run<2>(w); run<3>(w); run<4>(w); run<6>(w); run<8>(w); }
    return 0;
}
