// Collision-kernel microbenchmark (development tool, not part of the product): BASELINE config 3 shaped
// synthetic lattice (P paths x 49 points x 3 circles vs M obstacle points), all-FP64 kernel vs the
// FP32-filtered kernel; checks that the flags are identical and prints event-timed durations.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../include/b200mp.h"

static double lcg(unsigned long long &s) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(s >> 11) * (1.0 / 9007199254740992.0); }

int main(int argc, char **argv)
{
    const int P = argc > 1 ? atoi(argv[1]) : 4096, M = argc > 2 ? atoi(argv[2]) : 10000, n_pts = 49;
    const double shift = argc > 3 ? atof(argv[3]) : 0.0;   // moves the whole scene away from the origin
    const double oshift = argc > 4 ? atof(argv[4]) : 0.0;  // moves only the obstacles (large: nothing collides, no early exit)
    unsigned long long seed = 20261018;
    std::vector<double> px((size_t)P * n_pts), py(px.size()), pc(px.size()), ps(px.size()), obs((size_t)2 * M);
    for (int p = 0; p < P; ++p) {
        double x = 100 * lcg(seed) + shift, y = 100 * lcg(seed) + shift, yaw = -M_PI + 2 * M_PI * lcg(seed);
        const double k0 = -0.05 + 0.1 * lcg(seed), k1 = -0.05 + 0.1 * lcg(seed), ds = (20 + 20 * lcg(seed)) / 49;
        for (int j = 0; j < n_pts; ++j) {
            const size_t i = (size_t)p * n_pts + j;
            pc[i] = cos(yaw); ps[i] = sin(yaw);               // heading one sample behind the point (collision_checker.py:87-89)
            yaw += ds * (k0 + (k1 - k0) * j / 48.0);
            x += ds * cos(yaw); y += ds * sin(yaw);
            px[i] = x; py[i] = y;
        }
    }
    for (int m = 0; m < M;) {   // box outlines, 100 points each
        const double bx = 100 * lcg(seed) + shift + oshift, by = 100 * lcg(seed) + shift;
        for (int i = 0; i < 100 && m < M; ++i, ++m) {
            const double t = i / 100.0 * 21.0;
            double ox, oy;
            if (t < 6) { ox = t; oy = 0; } else if (t < 10.5) { ox = 6; oy = t - 6; } else if (t < 16.5) { ox = 16.5 - t; oy = 4.5; } else { ox = 0; oy = 21 - t; }
            obs[2 * (size_t)m] = bx + ox; obs[2 * (size_t)m + 1] = by + oy;
        }
    }
    double *d_px, *d_py, *d_pc, *d_ps, *d_obs; unsigned char *d_free;
    const size_t nb = px.size() * 8;
    cudaMalloc(&d_px, nb); cudaMalloc(&d_py, nb); cudaMalloc(&d_pc, nb); cudaMalloc(&d_ps, nb); cudaMalloc(&d_obs, obs.size() * 8); cudaMalloc(&d_free, P);
    cudaMemcpy(d_px, px.data(), nb, cudaMemcpyHostToDevice); cudaMemcpy(d_py, py.data(), nb, cudaMemcpyHostToDevice);
    cudaMemcpy(d_pc, pc.data(), nb, cudaMemcpyHostToDevice); cudaMemcpy(d_ps, ps.data(), nb, cudaMemcpyHostToDevice);
    cudaMemcpy(d_obs, obs.data(), obs.size() * 8, cudaMemcpyHostToDevice);
    const double off[3] = {-1.0, 1.0, 3.0}, rad[3] = {1.5, 1.5, 1.5};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<unsigned char> flags[3];
    for (int mode = 2; mode >= 0; --mode) {
        b200mp_set_collision_mode(mode);
        float best = 1e30f, sum = 0; const int reps = 10;
        for (int i = 0; i < 3 + reps; ++i) {
            cudaEventRecord(e0);
            if (b200mp_collision_check_f64(0, nullptr, P, n_pts, 3, off, rad, d_px, d_py, d_pc, d_ps, nullptr, n_pts, M, d_obs, d_free, nullptr)) { printf("collision: %s\n", b200mp_last_error()); return 1; }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (i >= 3) { sum += ms; if (ms < best) best = ms; }
        }
        if (mode == 0) { unsigned long long st2[2]; b200mp_collision_stats(0, nullptr, M, st2); printf("   broad phase: %llu warp-chunks screened of %lld, %llu thread-chunks rechecked\n", st2[0], (long long)((P * n_pts + 31) / 32) * ((M + 31) / 32), st2[1]); }
        flags[mode].resize(P);
        cudaMemcpy(flags[mode].data(), d_free, P, cudaMemcpyDeviceToHost);
        int nfree = 0; for (unsigned char f : flags[mode]) nfree += f;
        const double tests = (double)P * n_pts * 3 * M;
        printf("%s P=%d M=%d shift=%g  mean %.3f ms  best %.3f ms  %.3e nominal tests/s  free %d/%d  err=%s\n", mode == 1 ? "fp64_only" : (mode == 2 ? "screen   " : "cull     "), P, M, shift,
               sum / reps, best, tests / (sum / reps * 1e-3), nfree, P, cudaGetErrorString(cudaGetLastError()));
    }
    int diff = 0; for (int p = 0; p < P; ++p) diff += (flags[0][p] != flags[1][p]) + (flags[2][p] != flags[1][p]);
    printf("flag mismatches filtered vs fp64_only: %d\n", diff);
    return diff != 0;
}
