// K5: batched closed-loop path tracking -- Stanley lateral + PID longitudinal control around the RK4 step,
// one thread per vehicle (sm_100a, FP64).
//
// Replaces, for many vehicles at once, the control half of Car.drive (reference drive.py:126-151):
//   StanleyController.get_lookahead_index / stanley_control   libs/controllers/stanley_controller.py:56-129
//   LongitudinalController.long_control                       libs/controllers/stanley_controller.py:138-159
//   the first-order steering filter                            drive.py:137-138
// feeding VehicleModel.planar_model_RK4 (vehicle_rhs.cuh) without leaving the GPU, and writing the
// reference's 45-column DataLog row (drive.py:145-151) when asked to.
//
// The reference scans every waypoint (~3,000 at 1 cm spacing) with np.linalg.norm on each control update
// and then walks forward summing segment lengths until the look-ahead distance is reached.  Here:
//   * a prepare kernel evaluates, once per waypoint set, everything that does not depend on the vehicle:
//     segment lengths (same closed form as the host's norm, so the walk adds the same doubles), their
//     running sum (a search key only), segment headings, and one record per chunk of 8 ("fine"), 64 ("mid") and
//     512 ("coarse") consecutive waypoints: the chord from the chunk's first to its last waypoint and the largest
//     distance `eps` of any of its waypoints from that chord;
//   * the nearest-waypoint search is exact but sub-linear: the distance to the previous update's nearest
//     waypoint (or to the coarse chunks' first waypoints) is an upper bound UB on the minimum; every waypoint of a
//     chunk is at least dist(vehicle, chord) - eps away, so a chunk with dist(vehicle, chord) > UB + eps cannot
//     contain the minimum (rounding slack folded into eps and the comparison).  A vehicle that runs beside its
//     path at a lateral offset d sees the distance grow only quadratically along the path (sqrt(d^2 + s^2)), so the
//     former test "first waypoint of the chunk farther than UB + chunk radius" kept 2 sqrt(2 d r) / ds waypoints alive
//     (a dozen fine chunks at d = 0.3 m); the chord test keeps the one or two chunks around the foot point whatever d
//     is.  Coarse chunks are culled first, then the 8 mid chunks of a survivor, then the 8 fine chunks of a surviving
//     mid chunk, and the surviving waypoints are compared on the squared distance
//     in the host's rounding sequence (sqrt_rn is monotone).  The reference's "first strict minimum of the
//     rounded norms" is reproduced by a branch-free scan plus an exact replay when any two squares came
//     within 1e-15 of each other (ties, duplicate waypoints);
//   * the look-ahead walk is a binary search on the running sum with a rigorous rounding bound; only when
//     the bound cannot decide the index (|sum - lookahead| < ~1e-10) are the segment lengths added one by
//     one from the rounded minimum distance as the reference does.  Target indices are bit-identical.
// Waypoints and per-set tables are read through L1 (a set is ~120 KB; the vehicles of a CTA share one set).
#include <math.h>

#include <mutex>

#include "b200mp_internal.h"
#include "slice_sched.cuh"

namespace b200mp {

#ifndef B200MP_TRACK_BLOCK
#define B200MP_TRACK_BLOCK 128   /* measured: 64 / 128 / 256 threads with the rendezvous 2.54 / 2.51 / 2.61 ms (65,536 vehicles x 500 steps) */
#endif
#ifndef B200MP_TRACK_SYNC
#define B200MP_TRACK_SYNC 1
#endif
constexpr int kTrackBlock = B200MP_TRACK_BLOCK;
// CTA-wide rendezvous in front of every control update (arrival-counting barrier: the threads of a ragged last block that
// own no vehicle arrive from their own loop).  The controller's code (~25 KB hot) and the RK4 step loop (17 KB) do not fit
// the 32 KB instruction cache of an SM together; warps that drift apart keep both sets in flight (ncu: instruction-cache
// hit rate 88 %, `no_instruction` 14 % of the samples of a 500-step launch), warps that update together do not.
constexpr bool kTrackSync = B200MP_TRACK_SYNC != 0;
__device__ __forceinline__ void track_cta_rendezvous()
{
    // barrier.sync WITHOUT .aligned (bar.sync is the aligned form): threads of one warp arrive from two different loops when
    // the last block of a set is ragged, which the aligned form does not allow (it hangs)
    if (kTrackSync) asm volatile("barrier.sync 0;" ::: "memory");
}
constexpr int kFine = 8;             // waypoints per fine chunk
constexpr int kMid = 8 * kFine;      // one mid chunk = 8 fine chunks
constexpr int kCoarse = 8 * kMid;    // one coarse chunk = 8 mid chunks: every level is tested eight records at a time

// One chunk of consecutive waypoints [i0, i1] for the nearest-waypoint search: the chord a + t u (t in [0, 1]) from its
// first to its last waypoint, 1/|u|^2 (0 for a degenerate or non-finite chord: the "chord" is then the point a), and
// eps >= the distance of every waypoint of the chunk from the chord (+inf when the chunk holds a non-finite coordinate:
// such a chunk is never culled).  48 bytes = three 128-bit loads.
struct alignas(16) ChunkRec {
    double ax, ay, ux, uy, inv_l2, eps;
};

// squared distance from (x, y) to the chord of a record, as the device evaluates it (prepare and search share it)
__device__ __forceinline__ double chord_dist2(double ax, double ay, double ux, double uy, double inv_l2, double x, double y)
{
    const double vx = x - ax, vy = y - ay;
    double t = (vx * ux + vy * uy) * inv_l2;
    t = fmin(fmax(t, 0.0), 1.0);         // NaN -> 0
    const double dx = vx - t * ux, dy = vy - t * uy;
    return dx * dx + dy * dy;
}

__device__ __forceinline__ void track_prefetch_l1(const void *p)
{
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// false only when no waypoint of the chunk can be within ub of (x, y).  With d the true distance from the vehicle to the
// chord and s the evaluated one: |s - d| <= ~1e-15 (d + |u|) (the clamped foot point stays ON the chord, so a rounding of t
// can only lengthen the distance; the remaining operations are relatively accurate on operands bounded by d + |u|).  The
// relative part is covered by the factor below, the |u| part by the 1e-11 |u| the prepare kernel adds to eps (which also
// carries 1e-9 of each waypoint's deviation and 1e-11 of its distance from the chunk's first waypoint for the rounding of
// the deviations themselves).  NaN anywhere compares false: the chunk is searched.
__device__ __forceinline__ bool chunk_may_hold(const ChunkRec *__restrict__ rec, double x, double y, double ub)
{
    const double2 *__restrict__ p = reinterpret_cast<const double2 *>(rec);
    const double2 a = p[0], u = p[1], e = p[2];
    const double d2 = chord_dist2(a.x, a.y, u.x, u.y, e.x, x, y);
    const double reach = ub + e.y;
    return !(d2 > reach * reach * (1.0 + 1.0e-11));
}
constexpr double kPi = 3.141592653589793;

struct TrackDev {
    int V, n_steps, step0, ctrl_every, store_stride, n_sets, w_max, vps, blocks_per_set, norm_mode;
    double dt, target_vel, k, k_soft, max_steer, kp, ki, kd, lookahead, deadband, alpha;
    const double *state0, *ctrl0;
    const double2 *wp;
    const int *wp_count;
    const double *seg, *head, *cum, *clean;
    const ChunkRec *recs;   // [n_sets][recs_stride]: fine, mid, coarse chunk records of each waypoint set
    int recs_stride;
    double *traj, *log, *state_end, *ctrl_end;
    int *target_idx;
    const double *mu_table;   // friction table of parameter set 0 (vehicle_rhs.cuh), used by the no-log kernel
    double mu_B2;
    int2 *hint;               // [V] (nearest waypoint, previous look-ahead index) handed from one time slice to the next
};

__device__ __forceinline__ double host_sq(double v0, double v1, int mode)
{
    // the argument of the square root in the host's np.linalg.norm([v0, v1]) closed form (B200MP_NORM2_*)
    if (mode == B200MP_NORM2_FMA_V1) return __fma_rn(v1, v1, __dmul_rn(v0, v0));
    if (mode == B200MP_NORM2_FMA_V0) return __fma_rn(v0, v0, __dmul_rn(v1, v1));
    return __dadd_rn(__dmul_rn(v0, v0), __dmul_rn(v1, v1));
}

// numpy's float remainder for a positive divisor (np.float64.__mod__)
__device__ __forceinline__ double py_mod(double a, double b)
{
    double m = fabs(a) < b ? a : fmod(a, b);   // |a| < b: fmod returns a itself
    if (m != 0.0) {
        if (m < 0.0) m = __dadd_rn(m, b);
    } else {
        m = copysign(0.0, b);
    }
    return m;
}

__global__ void __launch_bounds__(256)
track_prepare_kernel(int w_max, const double2 *__restrict__ wp, const int *__restrict__ wp_count, int norm_mode,
                     double *__restrict__ seg, double *__restrict__ head, double *__restrict__ cum,
                     double *__restrict__ clean, ChunkRec *__restrict__ recs, int recs_stride)
{
    const int set = blockIdx.x;
    const int W = max(0, min(wp_count[set], w_max));   // a count beyond the array is clamped, never read out of bounds
    const double2 *w = wp + (size_t)set * w_max;
    // chunk records of the three levels, one compact array per set (fine, then mid, then coarse).  Three passes: the chords
    // (one thread per record), then every waypoint's distance from the chord of its fine, mid and coarse chunk folded into
    // the record with an atomic max (a coarse chunk has 512 waypoints: one thread per record took 50 us per launch), then
    // the slack terms.
    const int n_fine = (w_max + kFine - 1) / kFine, n_mid = (w_max + kMid - 1) / kMid, n_coarse = (w_max + kCoarse - 1) / kCoarse;
    ChunkRec *rset = recs + (size_t)set * recs_stride;
    for (int c = threadIdx.x; c < n_fine + n_mid + n_coarse; c += blockDim.x) {
        const int size = c < n_fine ? kFine : (c < n_fine + n_mid ? kMid : kCoarse);
        const int idx = c < n_fine ? c : (c < n_fine + n_mid ? c - n_fine : c - n_fine - n_mid);
        const int i0 = idx * size;
        ChunkRec r;
        r.ax = r.ay = r.ux = r.uy = r.inv_l2 = 0.0;
        r.eps = i0 < W ? 0.0 : INFINITY;   // records past the end are never read
        if (i0 < W) {
            const int i1 = min(W, i0 + size) - 1;
            const double2 a = w[i0], b = w[i1];
            const double ux = b.x - a.x, uy = b.y - a.y;
            const double l2 = ux * ux + uy * uy;
            const bool chord = l2 > 0.0 && l2 < INFINITY;
            r.ax = a.x;
            r.ay = a.y;
            r.ux = chord ? ux : 0.0;
            r.uy = chord ? uy : 0.0;
            r.inv_l2 = chord ? 1.0 / l2 : 0.0;
        }
        rset[c] = r;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < W; j += blockDim.x) {
        const double2 q = w[j];
        const bool finite = fabs(q.x) < INFINITY && fabs(q.y) < INFINITY;
#pragma unroll
        for (int lvl = 0; lvl < 3; ++lvl) {
            ChunkRec *r = rset + (lvl == 0 ? j / kFine : (lvl == 1 ? n_fine + j / kMid : n_fine + n_mid + j / kCoarse));
            const double dev = sqrt(chord_dist2(r->ax, r->ay, r->ux, r->uy, r->inv_l2, q.x, q.y));
            const double fx = q.x - r->ax, fy = q.y - r->ay;
            // rounding of dev itself (see chunk_may_hold): 1e-11 of this waypoint's distance from the chunk's first one
            double e = dev * (1.0 + 1.0e-9) + 1.0e-11 * sqrt(fx * fx + fy * fy);
            if (!(finite && e < INFINITY)) e = INFINITY;   // a chunk with a non-finite coordinate is never culled
            atomicMax(reinterpret_cast<unsigned long long *>(&r->eps), (unsigned long long)__double_as_longlong(e));   // e >= 0
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < n_fine + n_mid + n_coarse; c += blockDim.x) {
        // the search's own evaluation error, proportional to the chord length (chunk_may_hold)
        ChunkRec *r = rset + c;
        const double len = sqrt(r->ux * r->ux + r->uy * r->uy);
        const double eps = r->eps + 1.0e-11 * len + 1.0e-300;
        r->eps = eps < INFINITY ? eps : INFINITY;
    }
    double *sg = seg + (size_t)set * w_max, *hd = head + (size_t)set * w_max, *cm = cum + (size_t)set * w_max;
    int ok = 1;
    for (int i = threadIdx.x; i < W; i += blockDim.x) {
        const double2 p = w[i];
        double sl = 0.0;
        if (i > 0) {
            const double2 q = w[i - 1];
            sl = __dsqrt_rn(host_sq(__dsub_rn(p.x, q.x), __dsub_rn(p.y, q.y), norm_mode));   // stanley_controller.py:72-74
        }
        sg[i] = sl;
        ok &= (sl >= 0.0 && sl < INFINITY);
        const double2 a = (i < W - 1) ? w[i + 1] : w[0];     // :105-116, the last waypoint wraps to the first
        const double2 b = (i < W - 1) ? p : w[W - 1];
        hd[i] = atan2(a.y - b.y, a.x - b.x);
    }
    __shared__ int sok;
    __shared__ double part[256];
    if (threadIdx.x == 0) sok = 1;
    __syncthreads();   // also orders the sg[] stores above before the reads below (same block)
    if (!ok) atomicAnd(&sok, 0);
    // running arc length: only a SEARCH KEY for the look-ahead walk (the walk's own sums are re-derived from it with an
    // error bound that holds for any order of summation, see lookahead_index); a set with a NaN/Inf segment walks
    // sequentially.  Blocked scan: each thread sums a contiguous run, thread 0 scans the 256 partial sums.
    const int run = (W + (int)blockDim.x - 1) / (int)blockDim.x;
    const int j0 = min(W, (int)threadIdx.x * run), j1 = min(W, j0 + run);
    double acc = 0.0;
    for (int i = j0; i < j1; ++i) acc += sg[i];
    part[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double off = 0.0;
        for (int t = 0; t < (int)blockDim.x; ++t) {
            const double v = part[t];
            part[t] = off;
            off += v;
        }
        clean[set] = sok ? 1.0 : 0.0;
    }
    __syncthreads();
    acc = part[threadIdx.x];
    for (int i = j0; i < j1; ++i) {
        acc += sg[i];
        cm[i] = acc;
    }
}

struct SetView {
    const ChunkRec *__restrict__ fine, *__restrict__ mid, *__restrict__ coarse;   // chunk records of the three levels
    const double2 *__restrict__ w;
    const double *__restrict__ sg, *__restrict__ cm;
    int W;
    bool clean;
};

// first i in [lo, W) with base + (cm[i] - cm0) >= thr, else W   (cm is non-decreasing).  `hint` is where the answer
// was on the previous control update (a vehicle advances a few waypoints per update): the bracket is grown from
// there by doubling, so the usual cost is 3-4 probes instead of log2(W).
__device__ __forceinline__ int first_reaching(const double *__restrict__ cm, int lo, int W, double base, double cm0, double thr,
                                              int hint)
{
    int a = lo, b = W;      // invariant: every i < a fails the test, every i >= b passes it
    if (hint > lo && hint < W) {
        if (base + (cm[hint] - cm0) >= thr) {
            b = hint;
            for (int step = 1; b - step >= lo; step <<= 1) {
                if (base + (cm[b - step] - cm0) >= thr) {
                    b -= step;
                } else {
                    a = b - step + 1;
                    break;
                }
            }
        } else {
            a = hint + 1;
            for (int step = 1; a + step - 1 < W; step <<= 1) {
                const int j = a + step - 1;
                if (base + (cm[j] - cm0) >= thr) {
                    b = j;
                    break;
                }
                a = j + 1;
            }
        }
    }
    while (a < b) {
        const int m = (a + b) >> 1;
        if (base + (cm[m] - cm0) >= thr) b = m; else a = m + 1;
    }
    return a;
}

// Upper bound on the squared minimum distance for a vehicle's first update (no hint yet): the nearest of the coarse chunks'
// first waypoints, then of that chunk's mid and fine chunks' first waypoints -- any waypoint is a valid bound, a near one
// keeps the candidate set small.  Out of line: runs once per vehicle, and the controller's hot code has to share the SM's
// 32 KB instruction cache with the RK4 step loop.
__device__ __noinline__ double first_update_bound(const SetView &sv, double x, double y)
{
    const int W = sv.W;
    const int n_coarse = (W + kCoarse - 1) / kCoarse;
    double qub = INFINITY;
    int cbest = 0;
    for (int c = 0; c < n_coarse; ++c) {
        const double px = sv.coarse[c].ax, py = sv.coarse[c].ay;
        const double q = (px - x) * (px - x) + (py - y) * (py - y);
        if (q < qub) {
            qub = q;
            cbest = c;
        }
    }
    const int nm0 = min(8, (W - cbest * kCoarse + kMid - 1) / kMid);
    int mbest = 0;
    for (int m = 0; m < nm0; ++m) {
        const double px = sv.mid[cbest * 8 + m].ax, py = sv.mid[cbest * 8 + m].ay;
        const double q = (px - x) * (px - x) + (py - y) * (py - y);
        if (q < qub) {
            qub = q;
            mbest = m;
        }
    }
    const int nf0 = min(8, (W - cbest * kCoarse - mbest * kMid + kFine - 1) / kFine);
    for (int f = 0; f < nf0; ++f) {
        const double px = sv.fine[cbest * 64 + mbest * 8 + f].ax, py = sv.fine[cbest * 64 + mbest * 8 + f].ay;
        qub = fmin(qub, (px - x) * (px - x) + (py - y) * (py - y));
    }
    return qub;
}

// Tests up to eight consecutive chunk records (count >= 1): bit k of the result = record k may hold the nearest waypoint.
// No branches (an index past the end is clamped onto the last valid record and its bit masked off), so the 24 loads of a
// batch are in flight together; one copy of the code serves the three levels.
__device__ __noinline__ unsigned chunk_test8(const ChunkRec *__restrict__ rec, int count, double x, double y, double ub)
{
    unsigned mask = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const bool hold = chunk_may_hold(rec + min(k, count - 1), x, y, ub);
        mask |= (hold && k < count ? 1u : 0u) << k;
    }
    return mask;
}

// Exact replay of the reference's comparison sequence (first strict minimum of the ROUNDED norms) over a superset of the
// chunks the fast scan visited; reached only when two squared distances came within 1e-15 of each other.  Out of line.
__device__ __noinline__ int near_tie_replay(const SetView &sv, double x, double y, double ub, int mode, double *qbest_out)
{
    const double2 *__restrict__ w = sv.w;
    const int W = sv.W;
    double qbest = INFINITY, dbest = INFINITY;
    int ibest = 0;
    for (int c = 0; c * kFine < W; ++c) {
        if (!chunk_may_hold(sv.fine + c, x, y, ub)) continue;
        const int i1 = min(W, (c + 1) * kFine);
        for (int i = c * kFine; i < i1; ++i) {
            const double2 v = w[i];
            const double q = host_sq(__dsub_rn(v.x, x), __dsub_rn(v.y, y), mode);
            if (q < qbest) {
                const double d = __dsqrt_rn(q);
                if (d < dbest) {
                    dbest = d;
                    qbest = q;
                    ibest = i;
                }
            }
        }
    }
    *qbest_out = qbest;
    return ibest;
}

// The reference's look-ahead walk itself (stanley_controller.py:68-75), one addition per waypoint: only when the running-sum
// search of lookahead_index cannot decide the index within its rounding bound, or the set has a NaN/Inf segment.  Out of line.
__device__ __noinline__ int lookahead_walk(const double *__restrict__ sg, int W, int min_idx, double min_dist, double lookahead)
{
    double total = min_dist;
    int la = min_idx;
    for (int i = min_idx + 1; i < W; ++i) {
        if (total >= lookahead) break;
        total = __dadd_rn(total, sg[i]);
        la = i;
    }
    return la;
}

// get_lookahead_index (stanley_controller.py:56-76): exact nearest waypoint, then the look-ahead walk.
// hint = a waypoint index near the vehicle (the previous update's nearest index) or -1.
__device__ __forceinline__ int lookahead_index(const SetView &sv, double x, double y, double lookahead, int mode, int hint,
                                            int la_hint, int *nearest)
{
    const double2 *__restrict__ w = sv.w;
    const int W = sv.W;
    const int n_coarse = (W + kCoarse - 1) / kCoarse;
    // upper bound on the minimum distance: any waypoint will do; the coarse chunks' first waypoints when there is no hint
    double qub = INFINITY;
    if (hint >= 0 && hint < W) {
        const double2 p = w[hint];
        const double dx = p.x - x, dy = p.y - y;
        qub = dx * dx + dy * dy;
    }
    if (!(qub < INFINITY)) qub = first_update_bound(sv, x, y);
    const double ub = sqrt(qub) * (1.0 + 1.0e-12);   // +inf / NaN: nothing is culled

    // The reference updates on `dist < min_dist` with dist = sqrt_rn(q).  sqrt_rn is monotone, so a square that is
    // larger than the running minimum by more than a few ulp never updates and one that is smaller by more than a
    // few ulp always does (the rounded roots differ); the scan below is branch-free on that basis and only notes
    // whether any square ever fell inside the +-1e-15 band around the running minimum, in which case (rare: an
    // exact or near tie) the chunks are re-scanned comparing rounded roots one by one.  A chunk whose chord is farther
    // than ub + eps cannot hold the minimum or a tie with it (chunk_may_hold).
    double qbest = INFINITY;
    int ibest = 0;
    bool near_tie = false;
    for (int g0 = 0; g0 < n_coarse; g0 += 8) {
        unsigned cmask = chunk_test8(sv.coarse + g0, min(8, n_coarse - g0), x, y, ub);
        while (cmask) {
            const int g = g0 + __ffs(cmask) - 1;
            cmask &= cmask - 1;
            const int base = g * kCoarse;
            unsigned mmask = chunk_test8(sv.mid + g * 8, min(8, (W - base + kMid - 1) / kMid), x, y, ub);
            while (mmask) {
                const int m = __ffs(mmask) - 1;
                mmask &= mmask - 1;
                const int mbase = base + m * kMid;
                unsigned mask = chunk_test8(sv.fine + g * 64 + m * 8, min(8, (W - mbase + kFine - 1) / kFine), x, y, ub);
                while (mask) {
                    const int i0 = mbase + (__ffs(mask) - 1) * kFine;
                    mask &= mask - 1;
#pragma unroll
                    for (int k = 0; k < kFine; ++k) {
                        const int i = min(i0 + k, W - 1);               // the last chunk may be short: re-reading W-1 is harmless
                        const double2 v = w[i];
                        const double q = host_sq(__dsub_rn(v.x, x), __dsub_rn(v.y, y), mode);
                        const bool better = q < qbest * (1.0 - 1.0e-15);
                        near_tie |= !better && q <= qbest * (1.0 + 1.0e-15) && q < INFINITY && i0 + k < W;
                        qbest = better ? q : qbest;
                        ibest = better ? i : ibest;
                    }
                }
            }
        }
    }
    if (near_tie) ibest = near_tie_replay(sv, x, y, ub, mode, &qbest);   // rare: an exact or near tie
    const int min_idx = (qbest < INFINITY) ? ibest : 0;
    const double min_dist = (qbest < INFINITY) ? __dsqrt_rn(qbest) : INFINITY;
    *nearest = min_idx;

    // look-ahead walk (:68-75): la = first i >= min_idx whose running sum T_i (T_min_idx = min_dist, T_i = fl(T_{i-1} +
    // seg_i)) reaches the look-ahead distance, else W - 1.  T_i differs from A_i = min_dist + (cum_i - cum_min_idx) by
    // at most err (rounding of at most W additions each side), so when the first index reaching lookahead - err and
    // the first reaching lookahead + err coincide that index is the walk's; otherwise (or with a NaN/Inf segment) walk.
    if (sv.clean && min_dist < INFINITY) {
        const double cm0 = sv.cm[min_idx];
        const double err = 4.5e-16 * (double)(W + 2) * (sv.cm[W - 1] + min_dist + fabs(lookahead));
        const int lo = first_reaching(sv.cm, min_idx, W, min_dist, cm0, lookahead - err, la_hint);
        // the usual case: the same index also reaches lookahead + err (one probe, the line is already in L1)
        const int hi = (lo >= W || min_dist + (sv.cm[lo] - cm0) >= lookahead + err)
                           ? lo : first_reaching(sv.cm, lo, W, min_dist, cm0, lookahead + err, -1);
        if (lo == hi) return min(lo, W - 1);
    }
    return lookahead_walk(sv.sg, W, min_idx, min_dist, lookahead);
}

// The launch is cut into (vehicle block, time chunk) items claimed by persistent CTAs through a ticket, exactly as the
// rollout kernel does (slice_sched.cuh): 65,536 vehicles at 8 warps/SM are 1.73 waves, i.e. the second wave used to run at
// 73 % occupancy for a whole launch.  A chunk is a multiple of ctrl_every and of store_stride, so a slice starts on a
// control update and a stored step; the carried state goes through state_end / ctrl_end / hint in L2.  Launches that are not
// sliced run the SAME kernel with one block per CTA (sc.counter == NULL, a run-time switch): one compiled instruction
// sequence for every launch shape, so the compiler's FMA contraction choices -- and with them the last bits -- do not depend
// on the fleet size (a separately compiled plain kernel differed from the sliced one by 1 ulp in 6 % of the vehicles).
template <bool LOG, bool TAB>
__global__ void __launch_bounds__(kTrackBlock)
track_kernel(const __grid_constant__ TrackDev a, const __grid_constant__ DevParams<double> P0, const __grid_constant__ SliceSched sc)
{
    __shared__ int s_item;
    extern __shared__ __align__(16) unsigned char s_dyn[];   // the friction table (dynamic: 48 KB and more)
    double *s_mu = reinterpret_cast<double *>(s_dyn);
    // Controller state of each vehicle between two control updates (steering-filter state, speed-error integral, previous
    // speed, the two search hints): parked in shared memory while the RK4 steps run.  The step loop's schedule is paid for
    // in live registers (profiles/r02_k1_instruction_diet.md: capping K1 at 224 of its 246 costs 20 %); carried in
    // registers, this state and the per-set table pointers took ~30 of them and the same step ran 7-15 % slower here
    // than in the rollout kernel.  volatile: the values must really leave the register file.
    volatile double *s_ctl = reinterpret_cast<volatile double *>(s_dyn) + (TAB ? kMuTableDoubles : 0) + threadIdx.x;
    MuTableView T;
    T.c = s_mu;
    T.B2 = a.mu_B2;
    if (TAB) {   // 16 bytes per load, as the rollout kernel stages it
        const double2 *src = reinterpret_cast<const double2 *>(a.mu_table);
        double2 *dst = reinterpret_cast<double2 *>(s_mu);
#pragma unroll 8
        for (int i = threadIdx.x; i < kMuTableDoubles / 2; i += kTrackBlock) dst[i] = src[i];
        __syncthreads();
    }
    const size_t V = (size_t)a.V;
    const bool SLICED = sc.counter != nullptr;   // run-time: one-chunk launches take one block per CTA, no ticket
    int item = blockIdx.x;
    for (;;) {
        if (SLICED) {
            if (threadIdx.x == 0) s_item = atomicAdd(sc.counter, 1);
            __syncthreads();
            item = s_item;
            __syncthreads();   // s_item is rewritten by the next round
            if (item >= sc.n_blocks * sc.n_chunks) break;
        }
        const int chunk_idx = SLICED ? item / sc.n_blocks : 0;
        const int blk = SLICED ? item - chunk_idx * sc.n_blocks : item;
        const int n_begin = SLICED ? chunk_idx * sc.chunk : 0;
        const int n_stop = SLICED ? min(a.n_steps, n_begin + sc.chunk) : a.n_steps;
        if (SLICED && chunk_idx > 0) {   // wait for this block's previous time-chunk
            if (threadIdx.x == 0)
                while (ld_acquire(sc.done + blk) < chunk_idx) __nanosleep(100);
            __syncthreads();
        }
        const int set = blk / a.blocks_per_set;
        const int local = (blk - set * a.blocks_per_set) * kTrackBlock + threadIdx.x;
        const int r = set * a.vps + local;
        if (local < a.vps && r < a.V) {
            const int W = max(0, min(a.wp_count[set], a.w_max));
            {
                double x_del, e_int, prev_v;
                int nearest = -1, la_prev = -1;
                if (!SLICED || chunk_idx == 0) {
                    x_del = a.ctrl0[r];
                    e_int = a.ctrl0[V + r];
                    prev_v = a.ctrl0[2 * V + r];
                } else {   // carried state, written by another SM: read through L2
                    x_del = __ldcg(a.ctrl_end + r);
                    e_int = __ldcg(a.ctrl_end + V + r);
                    prev_v = __ldcg(a.ctrl_end + 2 * V + r);
                    const int2 h = __ldcg(a.hint + r);   // search hints only: the indices found do not depend on them
                    nearest = h.x;
                    la_prev = h.y;
                }
                s_ctl[0] = x_del;
                s_ctl[kTrackBlock] = e_int;
                s_ctl[2 * kTrackBlock] = prev_v;
                s_ctl[3 * kTrackBlock] = __hiloint2double(la_prev, nearest);
            }
            double y[10], ax, ay;
            if (!SLICED || chunk_idx == 0) {
#pragma unroll
                for (int c = 0; c < 10; ++c) y[c] = a.state0[c * V + r];
                ax = a.state0[10 * V + r];
                ay = a.state0[11 * V + r];
            } else {
#pragma unroll
                for (int c = 0; c < 10; ++c) y[c] = __ldcg(a.state_end + c * V + r);
                ax = __ldcg(a.state_end + 10 * V + r);
                ay = __ldcg(a.state_end + 11 * V + r);
            }
            double delta = 0.0, tau = 0.0, cte = 0.0;
            WheelCtrl<double> c;
            set_steer<double, true>(c, &delta);
            c.tq[0] = c.tq[1] = c.tq[2] = c.tq[3] = 0.0;

            const size_t out0 = a.store_stride > 0 ? (size_t)(n_begin / a.store_stride) : 0;
            double *tp = a.traj ? a.traj + out0 * 10 * V + r : nullptr;
            double *lp = (LOG && a.log) ? a.log + out0 * 45 * V + r : nullptr;
            int until_store = a.store_stride;

            int n = n_begin;
            while (n < n_stop) {
                track_cta_rendezvous();
                {   // ---- controllers (drive.py:128-138) on the current state
                    const double v = y[0], yaw = y[7], px = y[8], py = y[9];
                    double x_del = s_ctl[0], e_int = s_ctl[kTrackBlock], prev_v = s_ctl[2 * kTrackBlock];
                    int ce = 0;
                    double raw;
                    if (W > 0) {
                        const double hints = s_ctl[3 * kTrackBlock];
                        int nearest = __double2loint(hints);
                        const int la_prev = __double2hiint(hints);
                        const double2 *w = a.wp + (size_t)set * a.w_max;
                        const double *hd = a.head + (size_t)set * a.w_max;
                        SetView sv;
                        sv.w = w;
                        sv.fine = a.recs + (size_t)set * a.recs_stride;
                        sv.mid = sv.fine + (a.w_max + kFine - 1) / kFine;
                        sv.coarse = sv.mid + (a.w_max + kMid - 1) / kMid;
                        sv.sg = a.seg + (size_t)set * a.w_max;
                        sv.cm = a.cum + (size_t)set * a.w_max;
                        sv.W = W;
                        sv.clean = a.clean[set] != 0.0;
                        ce = lookahead_index(sv, px, py, a.lookahead, a.norm_mode, nearest, la_prev, &nearest);
                        s_ctl[3 * kTrackBlock] = __hiloint2double(ce, nearest);
                        {   // the next update reads a few waypoints further along both indices: start the following cache
                            // lines of every table it walks towards L1 now, ctrl_every steps ahead of their use
                            const int in = min(nearest + 8, W - 1), ic = min(ce + 8, W - 1);
                            track_prefetch_l1(w + in);
                            track_prefetch_l1(sv.cm + min(nearest + 16, W - 1));
                            track_prefetch_l1(sv.fine + min(nearest / kFine + 2, (W - 1) / kFine));
                            track_prefetch_l1(w + ic);
                            track_prefetch_l1(sv.cm + min(ce + 16, W - 1));
                            track_prefetch_l1(hd + min(ce + 16, W - 1));
                        }
                        double sn, cs;
                        sincos(yaw, &sn, &cs);
                        const double2 t = w[ce];
                        const double cv0 = __dsub_rn(__dsub_rn(t.x, px), __dmul_rn(a.lookahead, cs));   // :88-92
                        const double cv1 = __dsub_rn(__dsub_rn(t.y, py), __dmul_rn(a.lookahead, sn));
                        cte = __dsqrt_rn(host_sq(cv0, cv1, a.norm_mode));
                        if (cte < a.deadband) cte = 0.0;                                               // :95-96
                        double che = atan2(cv1, cv0) - yaw;                                            // :99-102
                        che = py_mod(che + kPi, 2.0 * kPi) - kPi;
                        const double sgn = che > 0.0 ? 1.0 : (che < 0.0 ? -1.0 : che);
                        double he = hd[ce] - yaw;                                                      // :107-121
                        he = py_mod(he + kPi, 2.0 * kPi) - kPi;
                        const double steer = he + atan(__ddiv_rn(__dmul_rn(__dmul_rn(a.k, sgn), cte), v + a.k_soft));   // :122-124
                        raw = fmin(fmax(steer, -a.max_steer), a.max_steer);                            // :126
                        if (steer != steer) raw = steer;                                               // np.clip keeps NaN
                    } else {
                        raw = 0.0;
                        cte = 0.0;
                    }
                    const double vel_error = a.target_vel - v;                                         // :149-155
                    e_int = __dadd_rn(e_int, __dmul_rn(vel_error, a.dt));
                    const double pp = __dmul_rn(a.kp, vel_error), ii = __dmul_rn(a.ki, e_int);
                    const double dd = __ddiv_rn(__dmul_rn(a.kd, v - prev_v), a.dt);
                    tau = __dadd_rn(__dadd_rn(pp, ii), dd);
                    if (v <= 0.01) tau = fabs(tau);                                                    // :157-158
                    prev_v = v;
                    x_del = __dadd_rn(__dmul_rn(1.0 - a.alpha, x_del), __dmul_rn(a.alpha, raw));       // drive.py:137
                    delta = x_del;
                    set_steer<double, true>(c, &delta);
                    c.tq[0] = c.tq[1] = c.tq[2] = c.tq[3] = tau * P0.inv_Jw;
                    if (a.target_idx) a.target_idx[(size_t)(n / a.ctrl_every) * V + r] = ce;
                    s_ctl[0] = x_del;
                    s_ctl[kTrackBlock] = e_int;
                    s_ctl[2 * kTrackBlock] = prev_v;
                }
                const int n_end = min(n_stop, n + a.ctrl_every);
#pragma unroll 1
                for (; n < n_end; ++n) {
                    double sdot[LOG ? 10 : 1], outs[LOG ? 18 : 1];
                    // the DataLog needs state_dot and the 18 outputs (combined slips included) only for the steps it stores:
                    // those take the closed-form logging step, every other step of a logging launch the tabulated one
                    if (LOG && !(TAB && !(a.store_stride > 0 && until_store == 1)))
                        rk4_step<double, true, true, true, false, false>(P0, P0.Dc, c, a.dt, y, ax, ay, sdot, outs);
                    else
                        rk4_step<double, true, false, true, true, TAB>(P0, P0.Dc, c, a.dt, y, ax, ay, sdot, outs, T);
                    if (a.store_stride > 0 && --until_store == 0) {
                        until_store = a.store_stride;
                        if (tp) {
#pragma unroll
                            for (int k = 0; k < 10; ++k) __stcs(tp + k * V, y[k]);   // written once: streaming stores, as in the rollout kernel
                            tp += 10 * V;
                        }
                        if (LOG && lp) {   // the DataLog row of drive.py:145-151
                            __stcs(lp, __dmul_rn((double)(a.step0 + n), a.dt));
#pragma unroll
                            for (int k = 0; k < 10; ++k) __stcs(lp + (1 + k) * V, y[k]);
#pragma unroll
                            for (int k = 0; k < 10; ++k) __stcs(lp + (11 + k) * V, sdot[k]);
                            __stcs(lp + 21 * V, delta);
#pragma unroll
                            for (int k = 0; k < 4; ++k) __stcs(lp + (22 + k) * V, tau);
#pragma unroll
                            for (int k = 0; k < 18; ++k) __stcs(lp + (26 + k) * V, outs[k]);
                            __stcs(lp + 44 * V, cte);
                            lp += 45 * V;
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 10; ++k) a.state_end[k * V + r] = y[k];
            a.state_end[10 * V + r] = ax;
            a.state_end[11 * V + r] = ay;
            a.ctrl_end[r] = s_ctl[0];
            a.ctrl_end[V + r] = s_ctl[kTrackBlock];
            a.ctrl_end[2 * V + r] = s_ctl[2 * kTrackBlock];
            if (SLICED && chunk_idx + 1 < sc.n_chunks) {
                const double hints = s_ctl[3 * kTrackBlock];
                a.hint[r] = make_int2(__double2loint(hints), __double2hiint(hints));
            }
        } else if (kTrackSync) {   // no vehicle: keep the rendezvous count of the block
            for (int n = n_begin; n < n_stop; n += a.ctrl_every) track_cta_rendezvous();
        }
        if (SLICED && chunk_idx + 1 < sc.n_chunks) {   // publish the carried state of this block
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) st_release(sc.done + blk, chunk_idx + 1);
        }
        if (!SLICED) break;
    }
}

int launch_track_f64(int device, cudaStream_t st, const B200mpTrackArgs &g)
{
    if (g.V < 0 || g.n_steps < 0 || g.ctrl_every < 1 || g.step0 < 0 || g.store_stride < 0 || g.n_sets < 1 || g.w_max < 0 ||
        g.vehicles_per_set < 1) {
        set_error("track: bad sizes V=%d n_steps=%d step0=%d ctrl_every=%d store_stride=%d n_sets=%d w_max=%d vehicles_per_set=%d",
                  g.V, g.n_steps, g.step0, g.ctrl_every, g.store_stride, g.n_sets, g.w_max, g.vehicles_per_set);
        return B200MP_E_ARG;
    }
    if (g.step0 % g.ctrl_every != 0 || (g.store_stride > 0 && g.step0 % g.store_stride != 0)) {
        set_error("track: step0 must be a multiple of ctrl_every and of store_stride");
        return B200MP_E_ARG;
    }
    if (g.V == 0) return 0;
    if (!g.state0 || !g.ctrl0 || !g.state_end || !g.ctrl_end || !g.wp_count || (g.w_max > 0 && !g.waypoints)) {
        set_error("track: state0, ctrl0, waypoints, wp_count, state_end and ctrl_end must be non-NULL");
        return B200MP_E_ARG;
    }
    if ((long long)g.n_sets * g.vehicles_per_set < g.V) {
        set_error("track: n_sets * vehicles_per_set (%d * %d) does not cover V = %d", g.n_sets, g.vehicles_per_set, g.V);
        return B200MP_E_ARG;
    }
    if (g.norm_mode < 0 || g.norm_mode > 2) {
        set_error("track: norm_mode %d", g.norm_mode);
        return B200MP_E_ARG;
    }
    if ((((size_t)g.waypoints) & 15) != 0) {
        set_error("track: waypoints must be 16-byte aligned");
        return B200MP_E_ARG;
    }
    DeviceState &ds = dev_state(device);
    if (ds.n_sets < 1) {
        set_error("track: no parameter table on device %d (call b200mp_set_params first)", device);
        return B200MP_E_PARAMS;
    }
    for (int i = 1; i < 4; ++i)
        if (ds.set0.B[i] != ds.set0.B[0] || ds.set0.C[i] != ds.set0.C[0] || ds.set0.D[i] != ds.set0.D[0]) {
            set_error("track: parameter set 0 must carry one (B, C, D) for the four tyres (vehicle_model.py:41-54)");
            return B200MP_E_PARAMS;
        }
    if (g.n_steps == 0) {
        if (g.state_end != g.state0)
            B200MP_CUDA(cudaMemcpyAsync(g.state_end, g.state0, sizeof(double) * 12 * (size_t)g.V, cudaMemcpyDeviceToDevice, st));
        if (g.ctrl_end != g.ctrl0)
            B200MP_CUDA(cudaMemcpyAsync(g.ctrl_end, g.ctrl0, sizeof(double) * 3 * (size_t)g.V, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    const size_t per = (size_t)g.n_sets * (size_t)(g.w_max > 0 ? g.w_max : 1);
    const int recs_stride = (g.w_max + kFine - 1) / kFine + (g.w_max + kMid - 1) / kMid + (g.w_max + kCoarse - 1) / kCoarse + 1;
    void *scratch = nullptr;
    const size_t recs_off = (sizeof(double) * (3 * per + (size_t)g.n_sets) + 15) & ~(size_t)15;   // records are 16-byte aligned
    const size_t tables_bytes = (recs_off + sizeof(ChunkRec) * (size_t)g.n_sets * (size_t)recs_stride + 15) & ~(size_t)15;
    int rc = ensure_scratch(device, st, tables_bytes + sizeof(int2) * (size_t)g.V, &scratch);
    if (rc) return rc;
    void *hint_area = (char *)scratch + tables_bytes;   // search hints handed between the time slices of a launch
    TrackDev a;
    a.V = g.V;
    a.n_steps = g.n_steps;
    a.step0 = g.step0;
    a.ctrl_every = g.ctrl_every;
    a.store_stride = (g.traj || g.log) ? g.store_stride : 0;
    a.n_sets = g.n_sets;
    a.w_max = g.w_max;
    a.vps = g.vehicles_per_set;
    a.blocks_per_set = (g.vehicles_per_set + kTrackBlock - 1) / kTrackBlock;
    a.norm_mode = g.norm_mode;
    a.dt = g.dt;
    a.target_vel = g.target_vel;
    a.k = g.k;
    a.k_soft = g.k_soft;
    a.max_steer = g.max_steer;
    a.kp = g.kp;
    a.ki = g.ki;
    a.kd = g.kd;
    a.lookahead = g.lookahead;
    a.deadband = g.deadband;
    a.alpha = g.steer_filter;
    a.state0 = g.state0;
    a.ctrl0 = g.ctrl0;
    a.wp = (const double2 *)g.waypoints;
    a.wp_count = g.wp_count;
    a.seg = (double *)scratch;
    a.head = (double *)scratch + per;
    a.cum = (double *)scratch + 2 * per;
    a.clean = (double *)scratch + 3 * per;
    a.recs = (const ChunkRec *)((char *)scratch + recs_off);
    a.recs_stride = recs_stride;
    a.traj = g.traj;
    a.log = g.log;
    a.state_end = g.state_end;
    a.ctrl_end = g.ctrl_end;
    a.target_idx = g.target_idx;
    a.hint = nullptr;
    track_prepare_kernel<<<g.n_sets, 256, 0, st>>>(g.w_max, a.wp, g.wp_count, g.norm_mode, (double *)a.seg, (double *)a.head,
                                                   (double *)a.cum, (double *)a.clean, (ChunkRec *)a.recs, recs_stride);
    B200MP_CUDA(cudaGetLastError());
    const DevParams<double> P0 = derive_params<double>(ds.set0);
    const long long grid = (long long)g.n_sets * a.blocks_per_set;
    if (grid > 0x7fffffffLL) {
        set_error("track: grid too large");
        return B200MP_E_ARG;
    }
    a.mu_table = ds.mu_table;
    a.mu_B2 = ds.mu_table_B2;
    const int fmode = (g.friction_override >= 1 && g.friction_override <= 2) ? g.friction_override - 1 : friction_mode();
    const bool tab = ds.mu_table && ds.mu_table_B2 > 0.0 && fmode == B200MP_FRICTION_AUTO;
    const size_t smem_tab = sizeof(double) * kMuTableDoubles;
    const bool log = g.log != nullptr;
    const bool tab_used = tab && (!log || g.store_stride > 1);   // a logging launch that stores every step never takes the table
    const size_t smem = (tab_used ? smem_tab : 0) + sizeof(double) * 4 * kTrackBlock;   // + the controller-state stash
    // the four (LOG, TAB) combinations
    typedef void (*Kern)(const TrackDev, const DevParams<double>, const SliceSched);
    const Kern kern = log ? (tab_used ? (Kern)track_kernel<true, true> : (Kern)track_kernel<true, false>)
                          : (tab_used ? (Kern)track_kernel<false, true> : (Kern)track_kernel<false, false>);
    // opt-in for dynamic shared memory beyond 48 KB and the resident-CTA count: once per (kernel, device), cached
    static std::mutex info_mu;
    static struct { const void *fn; int dev; int resident; } info[16];
    static int n_info = 0;
    int dev_id = 0, resident = 0;
    B200MP_CUDA(cudaGetDevice(&dev_id));
    {
        std::lock_guard<std::mutex> lock(info_mu);
        for (int i = 0; i < n_info; ++i)
            if (info[i].fn == (const void *)kern && info[i].dev == dev_id) resident = info[i].resident;
    }
    if (!resident) {
        if (smem) B200MP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int sms = 0, occ = 0;
        B200MP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id));
        B200MP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kTrackBlock, smem));
        resident = sms * (occ > 0 ? occ : 1);
        std::lock_guard<std::mutex> lock(info_mu);
        if (n_info < 16) {
            info[n_info].fn = (const void *)kern;
            info[n_info].dev = dev_id;
            info[n_info].resident = resident;
            ++n_info;
        }
    }
    // time slicing: only when the launch is more than one wave but too few waves for the tail to vanish
    const int n_blocks = (int)grid;
    SliceSched sc{nullptr, nullptr, n_blocks, 1, g.n_steps};
    int unit = g.ctrl_every;   // a slice starts on a control update and on a stored step
    if (a.store_stride > 1) {
        int x = unit, y = a.store_stride;
        while (y) { const int t = x % y; x = y; y = t; }
        unit = unit / x * a.store_stride;
    }
    if (n_blocks > resident && n_blocks < 8 * resident && g.n_steps >= 4 * unit &&
        sizeof(int) * ((size_t)n_blocks + 1) <= kSchedSlotBytes) {
        long long want = (12LL * resident + n_blocks - 1) / n_blocks;      // ~12 rounds of items
        int chunk = (int)((g.n_steps + want - 1) / want);
        chunk = (chunk + unit - 1) / unit * unit;
        if (chunk < 2 * unit) chunk = 2 * unit;
        const int n_chunks = (g.n_steps + chunk - 1) / chunk;
        if (n_chunks > 1) {
            sc.n_chunks = n_chunks;
            sc.chunk = chunk;
        }
    }
    a.hint = (int2 *)hint_area;
    if (sc.n_chunks > 1) {
        void *sched_mem = nullptr;
        cudaEvent_t sched_done = nullptr;
        rc = acquire_sched_slot(device, &sched_mem, &sched_done);
        if (rc) return rc;
        B200MP_CUDA(cudaMemsetAsync(sched_mem, 0, sizeof(int) * ((size_t)n_blocks + 1), st));
        sc.counter = (int *)sched_mem;
        sc.done = (int *)sched_mem + 1;
        const long long items = (long long)n_blocks * sc.n_chunks;
        kern<<<(int)(items < resident ? items : resident), kTrackBlock, smem, st>>>(a, P0, sc);
        cudaError_t e = cudaGetLastError();
        cudaError_t e2 = cudaEventRecord(sched_done, st);
        if (e == cudaSuccess) e = e2;
        if (e != cudaSuccess) return cuda_fail(e, "track_kernel launch");
        return 0;
    }
    kern<<<n_blocks, kTrackBlock, smem, st>>>(a, P0, sc);   // one block per CTA, one chunk: the same kernel without the ticket
    B200MP_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200mp
