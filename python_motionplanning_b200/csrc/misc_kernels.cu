// Reductions, MPC control sampling and the pipe-peak microbenchmark (sm_100a).
#include <math.h>
#include <stdint.h>

#include "b200mp_internal.h"

namespace b200mp {

// ------------------------------------------------------------------------------------------ argmin
// Lowest-index argmin, the tie convention of the reference's path selection (collision_checker.py:199,
// strict `<` while scanning upwards).  NaN is treated as +inf; all-+inf gives index -1 (the reference's None).
struct MinIdx {
    double v;
    long long i;
};

__device__ __forceinline__ MinIdx better(MinIdx a, MinIdx b)
{
    // an index of -1 marks "nothing finite seen"
    if (b.i < 0) return a;
    if (a.i < 0) return b;
    return (b.v < a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}

__device__ __forceinline__ MinIdx warp_min(MinIdx m)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        MinIdx o;
        o.v = __shfl_down_sync(0xffffffffu, m.v, d);
        o.i = __shfl_down_sync(0xffffffffu, m.i, d);
        m = better(m, o);
    }
    return m;
}

constexpr int kArgminBlock = 256;
constexpr int kArgminMaxBlocks = 1184;   // 148 SMs x 8

__device__ __forceinline__ MinIdx block_min(MinIdx m)
{
    __shared__ MinIdx part[kArgminBlock / 32];
    m = warp_min(m);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        MinIdx x;
        x.v = INFINITY;
        x.i = -1;
        if (threadIdx.x < kArgminBlock / 32) x = part[threadIdx.x];
        m = warp_min(x);
    }
    return m;
}

__global__ void __launch_bounds__(kArgminBlock)
argmin_partial_kernel(long long n, const double *__restrict__ cost, MinIdx *__restrict__ partial)
{
    MinIdx m;
    m.v = INFINITY;
    m.i = -1;
    for (long long i = (long long)blockIdx.x * kArgminBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kArgminBlock) {
        const double c = cost[i];
        if (c < INFINITY) {   // false for NaN and +inf
            MinIdx x;
            x.v = c;
            x.i = i;
            m = better(m, x);
        }
    }
    m = block_min(m);
    if (threadIdx.x == 0) partial[blockIdx.x] = m;
}

__global__ void __launch_bounds__(kArgminBlock)
argmin_final_kernel(int n_partial, const MinIdx *__restrict__ partial, long long index_offset, double *min_out,
                    long long *idx_out, int *idx32_out)
{
    MinIdx m;
    m.v = INFINITY;
    m.i = -1;
    for (int i = threadIdx.x; i < n_partial; i += kArgminBlock) m = better(m, partial[i]);
    m = block_min(m);
    if (threadIdx.x == 0) {
        const long long idx = m.i < 0 ? -1 : m.i + index_offset;
        if (min_out) *min_out = m.i < 0 ? INFINITY : m.v;
        if (idx_out) *idx_out = idx;
        if (idx32_out) *idx32_out = (int)idx;
    }
}

// argmin_final_kernel + gather of the winner's control sequence into one record (see b200mp_mpc_winner_f64)
__global__ void __launch_bounds__(kArgminBlock)
mpc_winner_final_kernel(int n_partial, const MinIdx *__restrict__ partial, long long index_offset, long long B, int n_seg,
                        const double *__restrict__ delta, const double *__restrict__ torque, double *__restrict__ record)
{
    __shared__ long long s_win;
    MinIdx m;
    m.v = INFINITY;
    m.i = -1;
    for (int i = threadIdx.x; i < n_partial; i += kArgminBlock) m = better(m, partial[i]);
    m = block_min(m);
    if (threadIdx.x == 0) {
        s_win = m.i;
        record[0] = m.i < 0 ? INFINITY : m.v;
        record[1] = m.i < 0 ? -1.0 : (double)(m.i + index_offset);
    }
    __syncthreads();
    const long long w = s_win;
    for (int s = threadIdx.x; s < n_seg; s += kArgminBlock) {
        record[2 + s] = w < 0 ? 0.0 : delta[(size_t)s * B + w];
        record[2 + n_seg + s] = w < 0 ? 0.0 : torque[(size_t)s * B + w];
    }
}

static int argmin_blocks(long long n)
{
    long long b = (n + kArgminBlock - 1) / kArgminBlock;
    if (b < 1) b = 1;
    if (b > kArgminMaxBlocks) b = kArgminMaxBlocks;
    return (int)b;
}

size_t argmin_scratch_bytes(long long n) { return sizeof(MinIdx) * (size_t)argmin_blocks(n); }

int argmin_launch(cudaStream_t st, long long n, const double *cost, long long index_offset, void *scratch,
                  double *min_out, long long *idx_out, int *idx32_out)
{
    const int nb = argmin_blocks(n);
    argmin_partial_kernel<<<nb, kArgminBlock, 0, st>>>(n, cost, (MinIdx *)scratch);
    B200MP_CUDA(cudaGetLastError());
    argmin_final_kernel<<<1, kArgminBlock, 0, st>>>(nb, (const MinIdx *)scratch, index_offset, min_out, idx_out, idx32_out);
    B200MP_CUDA(cudaGetLastError());
    return 0;
}

int launch_argmin_f64(int device, cudaStream_t st, long long n, const double *cost, long long index_offset,
                      double *min_out, long long *idx_out)
{
    if (n < 0 || (n > 0 && !cost) || (!min_out && !idx_out)) {
        set_error("argmin: bad arguments (n=%lld)", n);
        return B200MP_E_ARG;
    }
    void *scratch = nullptr;
    int rc = ensure_scratch(device, st, argmin_scratch_bytes(n), &scratch);
    if (rc) return rc;
    return argmin_launch(st, n, cost, index_offset, scratch, min_out, idx_out, nullptr);
}

int launch_mpc_winner_f64(int device, cudaStream_t st, long long B, int n_seg, const double *cost, const double *delta,
                          const double *torque, long long index_offset, double *record)
{
    if (B < 1 || n_seg < 0 || !cost || !record || (n_seg > 0 && (!delta || !torque))) {
        set_error("mpc_winner: bad arguments (B=%lld, n_seg=%d)", B, n_seg);
        return B200MP_E_ARG;
    }
    void *scratch = nullptr;
    int rc = ensure_scratch(device, st, argmin_scratch_bytes(B), &scratch);
    if (rc) return rc;
    const int nb = argmin_blocks(B);
    argmin_partial_kernel<<<nb, kArgminBlock, 0, st>>>(B, cost, (MinIdx *)scratch);
    B200MP_CUDA(cudaGetLastError());
    mpc_winner_final_kernel<<<1, kArgminBlock, 0, st>>>(nb, (const MinIdx *)scratch, index_offset, B, n_seg, delta, torque, record);
    B200MP_CUDA(cudaGetLastError());
    return 0;
}

// --------------------------------------------------------------------------- MPC control sampling
// Philox4x32-10 (Salmon et al., SC'11), counter = (rollout lo, rollout hi, segment, 0), key = seed.
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0;
        c[1] = lo1;
        c[2] = n2;
        c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

__global__ void __launch_bounds__(256)
mpc_sample_kernel(int B, int n_seg, uint32_t k0, uint32_t k1, long long rollout0, double delta_mean,
                  double delta_sigma, double delta_clip, double torque_mean, double torque_sigma,
                  double *__restrict__ delta, double *__restrict__ torque)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B) return;
    const unsigned long long gid = (unsigned long long)(rollout0 + r);
    for (int seg = 0; seg < n_seg; ++seg) {
        uint32_t c[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)seg, 0u};
        philox4x32_10(c, k0, k1);
        // two uniforms in (0, 1] and [0, 1) with 53 bits -> one Box-Muller pair
        const unsigned long long a = ((unsigned long long)c[0] << 32) | c[1];
        const unsigned long long b = ((unsigned long long)c[2] << 32) | c[3];
        const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);
        const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
        const double rad = sqrt(-2.0 * log(u1));
        double sn, cs;
        sincospi(2.0 * u2, &sn, &cs);
        const double e0 = rad * cs, e1 = rad * sn;
        double d = delta_mean + delta_sigma * e0;
        d = fmin(fmax(d, -delta_clip), delta_clip);
        delta[(size_t)seg * B + r] = d;
        torque[(size_t)seg * B + r] = torque_mean + torque_sigma * e1;
    }
}

int launch_mpc_sample_f64(cudaStream_t st, int B, int n_seg, unsigned long long seed, long long rollout0,
                          double delta_mean, double delta_sigma, double delta_clip, double torque_mean,
                          double torque_sigma, double *delta, double *torque)
{
    if (B < 0 || n_seg < 0 || !delta || !torque) {
        set_error("mpc_sample_controls: bad arguments (B=%d n_seg=%d)", B, n_seg);
        return B200MP_E_ARG;
    }
    if (B == 0 || n_seg == 0) return 0;
    mpc_sample_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, n_seg, (uint32_t)seed, (uint32_t)(seed >> 32), rollout0,
                                                      delta_mean, delta_sigma, delta_clip, torque_mean, torque_sigma,
                                                      delta, torque);
    B200MP_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------ pipe-peak benchmark
// Register-resident FMA chains: 8 independent accumulators per thread, no memory traffic.  This is the
// denominator of the rollout kernels' roofline (MEASURED_PEAKS.json has no FP64 / FP32 CUDA-core entry).
constexpr int kPeakChains = 8;
constexpr int kPeakInner = 64;

template <typename R>
__global__ void __launch_bounds__(256)
fma_peak_kernel(int iters, R a, R b, R *out)
{
    R x[kPeakChains];
#pragma unroll
    for (int k = 0; k < kPeakChains; ++k) x[k] = (R)(threadIdx.x + k) * (R)1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kPeakInner; ++u) {
#pragma unroll
            for (int k = 0; k < kPeakChains; ++k) x[k] = fma(x[k], a, b);
        }
    }
    R s = 0;
#pragma unroll
    for (int k = 0; k < kPeakChains; ++k) s += x[k];
    if (s == (R)123456789) out[0] = s;   // never true; keeps the chains alive
}

int run_fma_peak(int dtype_bits, int reps, double *tflops_out)
{
    if ((dtype_bits != 64 && dtype_bits != 32) || reps < 1 || !tflops_out) {
        set_error("fma_peak: dtype_bits must be 64 or 32, reps >= 1");
        return B200MP_E_ARG;
    }
    cudaDeviceProp prop;
    int dev = 0;
    B200MP_CUDA(cudaGetDevice(&dev));
    B200MP_CUDA(cudaGetDeviceProperties(&prop, dev));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    const int iters = dtype_bits == 64 ? 512 : 2048;
    void *out = nullptr;
    B200MP_CUDA(cudaMalloc(&out, 16));
    cudaEvent_t e0, e1;
    B200MP_CUDA(cudaEventCreate(&e0));
    B200MP_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int r = 0; r < reps + 1; ++r) {
        B200MP_CUDA(cudaEventRecord(e0, 0));
        if (dtype_bits == 64)
            fma_peak_kernel<double><<<blocks, threads>>>(iters, 0.999999, 1e-7, (double *)out);
        else
            fma_peak_kernel<float><<<blocks, threads>>>(iters, 0.9999f, 1e-5f, (float *)out);
        B200MP_CUDA(cudaEventRecord(e1, 0));
        B200MP_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        B200MP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flop = 2.0 * (double)blocks * threads * kPeakChains * kPeakInner * (double)iters;
        const double tf = flop / (ms * 1e-3) * 1e-12;
        if (r > 0 && tf > best) best = tf;   // first launch is the warm-up
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops_out = best;
    return 0;
}

}  // namespace b200mp
