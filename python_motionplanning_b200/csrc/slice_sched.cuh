// Time-sliced scheduling shared by the rollout (K1) and tracking (K5) kernels.
#pragma once

namespace b200mp {

// Time-sliced scheduling.  Every thread's work is identical, so a batch whose warp count is not a
// multiple of what the GPU holds at once ends in a tail wave at low occupancy (65,536 rollouts at 12
// warps per SM = 1.15 waves: the last 15 % of the blocks run alone for a full rollout).  Rollouts are
// resumable, so instead the launch is cut into (block, time-chunk) items which a grid of persistent CTAs
// (as many as are resident at once) claims through an atomic ticket in chunk-major order; item (b, c)
// starts once done[b] == c, carrying the state through state_end (and the running cost through cost).
// An item's predecessor was always claimed earlier by a CTA that is running or done, so waiting cannot
// deadlock.
struct SliceSched {
    int *counter;   // next item to claim
    int *done;      // [n_blocks] chunks completed per rollout block
    int n_blocks, n_chunks, chunk;
};

__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

}  // namespace b200mp
