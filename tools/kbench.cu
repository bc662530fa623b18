// Kernel-variant microbenchmark for the FP64 rollout kernel (development tool, not part of the product).
// Links the library's own translation units compiled with different -D tunables, runs BASELINE config 2
// (65,536 rollouts x 500 steps, full trajectory) and prints the event-timed duration and a checksum.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../include/b200mp.h"

#ifdef B200MP_SLICE_PROFILE
extern "C" int b200mp_debug_slice_prof(unsigned long long *, int);
#endif
static double lcg(unsigned long long &s) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(s >> 11) * (1.0 / 9007199254740992.0); }

int main(int argc, char **argv)
{
    const int B = argc > 1 ? atoi(argv[1]) : 65536, N = argc > 2 ? atoi(argv[2]) : 500, hold = getenv("KB_HOLD") ? atoi(getenv("KB_HOLD")) : 10;
    const int stride = argc > 3 ? atoi(argv[3]) : 1;
    if (argc > 5) b200mp_set_friction_mode(atoi(argv[5]));   // 1 = closed form
    B200mpVehicleParams p{};
    p.m = 987.89 + 869.93; p.b = 2.906 / 1.85; p.a = 2.906 - p.b; p.Izz = 0.5 * p.m * p.a * p.b; p.Jw = 1; p.hg = 0.55419;
    p.T = 1.536; p.wL = p.wR = p.T / 2; p.rw = 0.329 - (987.89 / 2 + 50) / 26290;
    for (int i = 0; i < 4; ++i) { p.B[i] = 20.6357; p.C[i] = 1.5047; p.D[i] = 1.0; }
    if (b200mp_set_params(0, &p, 1)) { printf("set_params: %s\n", b200mp_last_error()); return 1; }
    const int nseg = (N + hold - 1) / hold;
    std::vector<double> s0((size_t)12 * B), dl((size_t)nseg * B), tq((size_t)nseg * B);
    unsigned long long seed = 12345;
    const int coherent = argc > 6 ? atoi(argv[6]) : 0;   // 1: every rollout starts near one operating point (small slip)
    for (int r = 0; r < B; ++r) {
        if (coherent) {
            const double U = 25 + 0.5 * lcg(seed);
            s0[0 * (size_t)B + r] = U; s0[1 * (size_t)B + r] = 0.05 * (lcg(seed) - 0.5); s0[2 * (size_t)B + r] = 0.02 * (lcg(seed) - 0.5);
            for (int i = 0; i < 4; ++i) s0[(3 + i) * (size_t)B + r] = U / p.rw * (1 + 0.004 * (lcg(seed) - 0.5));
            s0[7 * (size_t)B + r] = -3.14 + 6.28 * lcg(seed); s0[8 * (size_t)B + r] = -100 + 200 * lcg(seed); s0[9 * (size_t)B + r] = -100 + 200 * lcg(seed);
            continue;
        }
        const double U = 5 + 35 * lcg(seed);
        s0[0 * (size_t)B + r] = U; s0[1 * (size_t)B + r] = -1 + 2 * lcg(seed); s0[2 * (size_t)B + r] = -0.5 + lcg(seed);
        for (int i = 0; i < 4; ++i) s0[(3 + i) * (size_t)B + r] = U / p.rw * (1 + 0.1 * (lcg(seed) - 0.5));
        s0[7 * (size_t)B + r] = -3.14 + 6.28 * lcg(seed); s0[8 * (size_t)B + r] = -100 + 200 * lcg(seed); s0[9 * (size_t)B + r] = -100 + 200 * lcg(seed);
    }
    for (size_t i = 0; i < dl.size(); ++i) { dl[i] = (coherent ? 0.1 : 1.0) * (-0.1 + 0.2 * lcg(seed)); tq[i] = -300 + 600 * lcg(seed); }
    double *d_s0, *d_dl, *d_tq, *d_traj = nullptr, *d_end;
    cudaMalloc(&d_s0, s0.size() * 8); cudaMalloc(&d_dl, dl.size() * 8); cudaMalloc(&d_tq, tq.size() * 8); cudaMalloc(&d_end, (size_t)12 * B * 8);
    if (stride) cudaMalloc(&d_traj, (size_t)(N / stride) * 10 * B * 8);
    cudaMemcpy(d_s0, s0.data(), s0.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_dl, dl.data(), dl.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_tq, tq.data(), tq.size() * 8, cudaMemcpyHostToDevice);
    B200mpRolloutArgs a{};
    a.B = B; a.n_steps = N; a.hold = hold; a.dt = 1e-4; a.state0 = d_s0; a.delta = d_dl; a.torque = d_tq; a.delta_ch = 1; a.torque_ch = 1;
    a.store_stride = stride; a.traj = d_traj; a.state_end = d_end;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f, sum = 0;
    const int reps = 8;
    for (int i = 0; i < 3 + reps; ++i) {
        cudaEventRecord(e0);
        if (b200mp_rk4_rollout_f64(0, nullptr, &a)) { printf("rollout: %s\n", b200mp_last_error()); return 1; }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (i >= 3) { sum += ms; if (ms < best) best = ms; }
    }
#ifdef B200MP_SLICE_PROFILE
    {
        unsigned long long q[8];
        b200mp_debug_slice_prof(nullptr, 1);
        b200mp_rk4_rollout_f64(0, nullptr, &a);
        cudaDeviceSynchronize();
        b200mp_debug_slice_prof(q, 0);
        const double it = (double)q[3];
        printf("   slice profile: %.0f items; per item: wait %.0f, prologue %.0f, steps %.0f, epilogue %.0f cycles; %.1f %% of the items found their predecessor unfinished\n",
               it, q[0] / it, q[1] / it, q[5] / it, q[2] / it, 100.0 * q[4] / it);
        printf("   SM clock seen by the CTAs: %.3f GHz (cycles / globaltimer ns over each CTA's life)\n", (double)q[6] / (double)q[7]);
    }
#endif
    std::vector<double> end((size_t)12 * B);
    cudaMemcpy(end.data(), d_end, end.size() * 8, cudaMemcpyDeviceToHost);
    double cs = 0; for (double v : end) cs += v;
    printf("%s B=%d N=%d stride=%d  mean %.3f ms  best %.3f ms  %.3e steps/s  checksum %.12e  err=%s\n", argc > 4 ? argv[4] : "", B, N, stride,
           sum / reps, best, (double)B * N / (sum / reps * 1e-3), cs, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
