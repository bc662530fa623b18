// C-ABI entry points of libb200mp.so (declared in include/b200mp.h).
//
// Thin, exception-free shims: validate, make the device current, launch on the caller's stream.
// Ownership: callers own every buffer they pass; the library owns only the per-device parameter
// table and a reduction scratch area, both released by b200mp_shutdown().
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <thread>
#include <vector>
#include <cstdint>

#include "b200mp_internal.h"

namespace b200mp {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what)
{
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return (int)e;
}

static DeviceState g_states[kMaxDevices];
static std::mutex g_mutex;

static std::atomic<int> g_friction_mode{B200MP_FRICTION_AUTO};
int friction_mode() { return g_friction_mode.load(std::memory_order_relaxed); }
static std::atomic<int> g_collision_mode{B200MP_COLLISION_AUTO};
int collision_mode() { return g_collision_mode.load(std::memory_order_relaxed); }

DeviceState &dev_state(int device) { return g_states[(device >= 0 && device < kMaxDevices) ? device : 0]; }

int ensure_scratch(int device, cudaStream_t stream, size_t bytes, void **out)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceState &ds = dev_state(device);
    if (bytes < 4096) bytes = 4096;
    if (ds.cull_stats_stream == stream) ds.cull_stats_ptr = nullptr;   // this stream's area is about to be overwritten
    DeviceState::Scratch *slot = nullptr, *lru = nullptr;
    for (auto &sc : ds.scratch) {
        if (sc.used && sc.stream == stream) slot = &sc;
        if (!lru || (!sc.used && lru->used) || (sc.used == lru->used && sc.last_use < lru->last_use)) lru = &sc;
    }
    if (!slot) {
        slot = lru;
        if (slot->used) {
            // more streams than slots: the least recently used area changes hands once its stream has drained
            B200MP_CUDA(cudaStreamSynchronize(slot->stream));
            if (ds.cull_stats_stream == slot->stream) ds.cull_stats_ptr = nullptr;
        }
        slot->stream = stream;
        slot->used = true;
    }
    slot->last_use = ++ds.scratch_clock;
    if (slot->bytes < bytes) {
        if (slot->ptr) {
            // a larger area is needed: wait for the work of this stream that may still use the old one
            B200MP_CUDA(cudaStreamSynchronize(stream));
            B200MP_CUDA(cudaFree(slot->ptr));
            slot->ptr = nullptr;
            slot->bytes = 0;
        }
        const size_t want = bytes + bytes / 2;
        B200MP_CUDA(cudaMalloc(&slot->ptr, want));
        slot->bytes = want;
    }
    *out = slot->ptr;
    return 0;
}

int acquire_sched_slot(int device, void **area, cudaEvent_t *done)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceState &ds = dev_state(device);
    if (!ds.sched_ring) {
        B200MP_CUDA(cudaMalloc(&ds.sched_ring, kSchedSlots * kSchedSlotBytes));
        for (int i = 0; i < kSchedSlots; ++i) B200MP_CUDA(cudaEventCreateWithFlags(&ds.sched_event[i], cudaEventDisableTiming));
    }
    const int slot = ds.sched_next;
    ds.sched_next = (slot + 1) % kSchedSlots;
    if (ds.sched_used[slot]) B200MP_CUDA(cudaEventSynchronize(ds.sched_event[slot]));   // 32 launches ago: long done
    ds.sched_used[slot] = true;
    *area = (char *)ds.sched_ring + (size_t)slot * kSchedSlotBytes;
    *done = ds.sched_event[slot];
    return 0;
}

DeviceGuard::DeviceGuard(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n < 1) {
        set_error("no usable CUDA device (%s); libb200mp has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        (void)cudaGetLastError();
        status_ = B200MP_E_NODEVICE;
        return;
    }
    if (device < 0 || device >= n || device >= kMaxDevices) {
        set_error("device %d out of range (have %d)", device, n);
        status_ = B200MP_E_ARG;
        return;
    }
    e = cudaGetDevice(&prev_);
    if (e != cudaSuccess) prev_ = -1;
    if (prev_ != device) {
        e = cudaSetDevice(device);
        if (e != cudaSuccess) status_ = cuda_fail(e, "cudaSetDevice");
    }
}

DeviceGuard::~DeviceGuard()
{
    if (status_ == 0 && prev_ >= 0) {
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != prev_) (void)cudaSetDevice(prev_);
    }
}

}  // namespace b200mp

using namespace b200mp;

#define B200MP_ENTER(device)          \
    g_err[0] = 0;                     \
    DeviceGuard guard__(device);      \
    if (guard__.status() != 0) return guard__.status()

// D = 1 friction tables for every parameter set with four equal (B, C) (see DeviceState::set_tables), built on
// host threads (23 ms of long-double trigonometry per table).
static cudaError_t upload_set_tables(DeviceState &ds, const B200mpVehicleParams *sets, int n_sets)
{
    if (ds.set_tables) (void)cudaFree(ds.set_tables);
    if (ds.set_B2) (void)cudaFree(ds.set_B2);
    ds.set_tables = nullptr;
    ds.set_B2 = nullptr;
    ds.set_tables_n = 0;
    if (ds.set_tables_f32) (void)cudaFree(ds.set_tables_f32);
    ds.set_tables_f32 = nullptr;
    std::vector<double> tables((size_t)n_sets * kMuTableDoubles + 2), B2(n_sets, 0.0);
    std::vector<float> tables32((size_t)n_sets * kMuTableFloats + 4);
    double *base = tables.data();
    if (reinterpret_cast<uintptr_t>(base) % 16) ++base;   // MuRow is 16-byte aligned
    float *base32 = tables32.data();
    while (reinterpret_cast<uintptr_t>(base32) % 16) ++base32;
    std::atomic<int> next(0);
    auto work = [&]() {
        for (int i = next.fetch_add(1); i < n_sets; i = next.fetch_add(1)) {
            const B200mpVehicleParams &h = sets[i];
            bool uniform = true;
            for (int w = 1; w < 4; ++w) uniform = uniform && h.B[w] == h.B[0] && h.C[w] == h.C[0];
            if (!(uniform && h.B[0] > 0.0 && h.C[0] > 0.0 && h.C[0] < 4.0)) continue;
            const double err = build_mu_table(h.B[0], h.C[0], 1.0, base + (size_t)i * kMuTableDoubles);
            const double err32 = build_mu_table_f32(h.B[0], h.C[0], 1.0, base32 + (size_t)i * kMuTableFloats);
            if (err < 1.0e-15 && err32 < 1.0e-6) B2[i] = h.B[0] * h.B[0];
        }
    };
    unsigned nt = std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > (unsigned)n_sets) nt = (unsigned)n_sets;
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    cudaError_t e = cudaMalloc((void **)&ds.set_tables, sizeof(double) * (size_t)n_sets * kMuTableDoubles);
    if (e == cudaSuccess) e = cudaMalloc((void **)&ds.set_B2, sizeof(double) * n_sets);
    if (e == cudaSuccess)
        e = cudaMemcpy(ds.set_tables, base, sizeof(double) * (size_t)n_sets * kMuTableDoubles, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(ds.set_B2, B2.data(), sizeof(double) * n_sets, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc((void **)&ds.set_tables_f32, sizeof(float) * (size_t)n_sets * kMuTableFloats);
    if (e == cudaSuccess)
        e = cudaMemcpy(ds.set_tables_f32, base32, sizeof(float) * (size_t)n_sets * kMuTableFloats, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) ds.set_tables_n = n_sets;
    return e;
}

extern "C" {

int b200mp_version(void) { return B200MP_VERSION; }

const char *b200mp_last_error(void) { return g_err; }

int b200mp_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
        return B200MP_E_NODEVICE;
    }
    return n;
}

int b200mp_set_params(int device, const B200mpVehicleParams *host_sets, int n_sets)
{
    B200MP_ENTER(device);
    if (!host_sets || n_sets < 1) {
        set_error("set_params: need at least one parameter set");
        return B200MP_E_ARG;
    }
    static_assert(sizeof(HostParams) == sizeof(B200mpVehicleParams), "parameter struct layout");
    DevParams<double> *h64 = new DevParams<double>[n_sets];
    DevParams<float> *h32 = new DevParams<float>[n_sets];
    for (int i = 0; i < n_sets; ++i) {
        HostParams hp;
        memcpy(&hp, &host_sets[i], sizeof(hp));
        h64[i] = derive_params<double>(hp);
        h32[i] = derive_params<float>(hp);
    }
    int rc = 0;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        DeviceState &ds = dev_state(device);
        cudaError_t e = cudaDeviceSynchronize();   // rollouts in flight may still read the old table
        if (e == cudaSuccess && ds.table64) e = cudaFree(ds.table64);
        if (e == cudaSuccess && ds.table32) e = cudaFree(ds.table32);
        ds.table64 = nullptr;
        ds.table32 = nullptr;
        ds.n_sets = 0;
        if (e == cudaSuccess) e = cudaMalloc((void **)&ds.table64, sizeof(DevParams<double>) * n_sets);
        if (e == cudaSuccess) e = cudaMalloc((void **)&ds.table32, sizeof(DevParams<float>) * n_sets);
        if (e == cudaSuccess) e = cudaMemcpy(ds.table64, h64, sizeof(DevParams<double>) * n_sets, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(ds.table32, h32, sizeof(DevParams<float>) * n_sets, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            memcpy(&ds.set0, &host_sets[0], sizeof(HostParams));
            ds.n_sets = n_sets;
            // friction table of set 0 (fast-path kernels); skipped when the four tyres differ or the fit is not at
            // rounding level (then the closed-form path is used)
            const HostParams &h = ds.set0;
            bool uniform = true;
            for (int i = 1; i < 4; ++i) uniform = uniform && h.B[i] == h.B[0] && h.C[i] == h.C[0] && h.D[i] == h.D[0];
            ds.mu_table_B2 = 0.0;
            ds.mu_table_f32_ok = false;
            if (uniform && h.B[0] > 0.0 && h.C[0] > 0.0 && h.C[0] < 4.0 && h.D[0] == h.D[0]) {
                alignas(16) static float host_table_f32[kMuTableFloats];
                if (build_mu_table_f32(h.B[0], h.C[0], h.D[0], host_table_f32) < 1.0e-6) {
                    if (!ds.mu_table_f32) e = cudaMalloc((void **)&ds.mu_table_f32, sizeof(host_table_f32));
                    if (e == cudaSuccess) e = cudaMemcpy(ds.mu_table_f32, host_table_f32, sizeof(host_table_f32), cudaMemcpyHostToDevice);
                    ds.mu_table_f32_ok = e == cudaSuccess;
                }
            }
            if (e == cudaSuccess && uniform && h.B[0] > 0.0 && h.C[0] > 0.0 && h.C[0] < 4.0 && h.D[0] == h.D[0]) {
                alignas(16) static double host_table[kMuTableDoubles];
                ds.mu_table_err = build_mu_table(h.B[0], h.C[0], h.D[0], host_table);
                if (ds.mu_table_err < 1.0e-15) {
                    if (!ds.mu_table) e = cudaMalloc((void **)&ds.mu_table, sizeof(host_table));
                    if (e == cudaSuccess) e = cudaMemcpy(ds.mu_table, host_table, sizeof(host_table), cudaMemcpyHostToDevice);
                    if (e == cudaSuccess) ds.mu_table_B2 = h.B[0] * h.B[0];
                }
            }
            if (e == cudaSuccess) e = upload_set_tables(ds, host_sets, n_sets);
            if (e != cudaSuccess) rc = cuda_fail(e, "set_params (friction table)");
        } else {
            rc = cuda_fail(e, "set_params");
        }
    }
    delete[] h64;
    delete[] h32;
    return rc;
}

int b200mp_rk4_rollout_f64(int device, void *stream, const B200mpRolloutArgs *args)
{
    B200MP_ENTER(device);
    if (!args) {
        set_error("rk4_rollout: args is NULL");
        return B200MP_E_ARG;
    }
    return launch_rollout_f64(device, (cudaStream_t)stream, *args);
}

int b200mp_rk4_rollout_f32(int device, void *stream, const B200mpRolloutArgs *args)
{
    B200MP_ENTER(device);
    if (!args) {
        set_error("rk4_rollout: args is NULL");
        return B200MP_E_ARG;
    }
    return launch_rollout_f32(device, (cudaStream_t)stream, *args);
}

int b200mp_planar_model_f64(int device, void *stream, int B, const double *state, const double *torque,
                            const double *mu, const double *delta, const double *axay, const int *param_set,
                            double *state_dot, double *misc, double *outputs)
{
    B200MP_ENTER(device);
    return launch_planar_model_f64(device, (cudaStream_t)stream, B, state, torque, mu, delta, axay, param_set, state_dot,
                                   misc, outputs);
}

int b200mp_mpc_sample_controls_f64(int device, void *stream, int B, int n_seg, unsigned long long seed,
                                   long long rollout0, double delta_mean, double delta_sigma, double delta_clip,
                                   double torque_mean, double torque_sigma, double *delta, double *torque)
{
    B200MP_ENTER(device);
    return launch_mpc_sample_f64((cudaStream_t)stream, B, n_seg, seed, rollout0, delta_mean, delta_sigma, delta_clip,
                                 torque_mean, torque_sigma, delta, torque);
}

int b200mp_argmin_f64(int device, void *stream, long long n, const double *cost, long long index_offset,
                      double *min_out, long long *idx_out)
{
    B200MP_ENTER(device);
    return launch_argmin_f64(device, (cudaStream_t)stream, n, cost, index_offset, min_out, idx_out);
}

int b200mp_mpc_winner_f64(int device, void *stream, long long B, int n_seg, const double *cost, const double *delta,
                          const double *torque, long long index_offset, double *record)
{
    B200MP_ENTER(device);
    return launch_mpc_winner_f64(device, (cudaStream_t)stream, B, n_seg, cost, delta, torque, index_offset, record);
}

int b200mp_collision_check_f64(int device, void *stream, int P, int n_pts, int n_circ, const double *off,
                               const double *rad, const double *px, const double *py, const double *pcos,
                               const double *psin, const double *pyaw, int yaw_stride, int M, const double *obs,
                               unsigned char *free_out, double *min_clear)
{
    B200MP_ENTER(device);
    return launch_collision_f64(device, (cudaStream_t)stream, P, n_pts, n_circ, off, rad, px, py, pcos, psin, pyaw,
                                yaw_stride, M, obs, free_out, min_clear, nullptr, 0, -1);
}

int b200mp_collision_check_yaw_f64(int device, void *stream, int P, int n_pts, int n_circ, const double *off,
                                   const double *rad, const double *px, const double *py, const double *pyaw,
                                   int yaw_stride, int M, const double *obs, unsigned char *free_out, int *undecided,
                                   int undecided_capacity, int mode)
{
    B200MP_ENTER(device);
    if (!undecided) {
        set_error("collision_check_yaw: undecided must be a device array of 1 + capacity ints");
        return B200MP_E_ARG;
    }
    if (mode != -1 && mode != B200MP_COLLISION_AUTO && mode != B200MP_COLLISION_FP64_ONLY && mode != B200MP_COLLISION_SCREEN_ONLY) {
        set_error("collision_check_yaw: unknown mode %d", mode);
        return B200MP_E_ARG;
    }
    return launch_collision_f64(device, (cudaStream_t)stream, P, n_pts, n_circ, off, rad, px, py, nullptr, nullptr, pyaw,
                                yaw_stride, M, obs, free_out, nullptr, undecided, undecided_capacity, mode);
}

int b200mp_collision_resolve_f64(int device, void *stream, int n_list, const int *items, const double *cos_sin, int P,
                                 int n_pts, int n_circ, const double *off, const double *rad, const double *px,
                                 const double *py, int M, const double *obs, unsigned char *free_out)
{
    B200MP_ENTER(device);
    return launch_collision_resolve_f64(device, (cudaStream_t)stream, n_list, items, cos_sin, P, n_pts, n_circ, off, rad,
                                        px, py, M, obs, free_out);
}

int b200mp_set_friction_mode(int mode)
{
    g_err[0] = 0;
    if (mode != B200MP_FRICTION_AUTO && mode != B200MP_FRICTION_CLOSED_FORM) {
        set_error("set_friction_mode: unknown mode %d", mode);
        return B200MP_E_ARG;
    }
    return g_friction_mode.exchange(mode);
}

int b200mp_collision_stats(int device, void *stream, int M, unsigned long long *out2)
{
    B200MP_ENTER(device);
    if (!out2) {
        set_error("collision_stats: out2 is NULL");
        return B200MP_E_ARG;
    }
    return collision_stats(device, (cudaStream_t)stream, M, out2);
}

int b200mp_set_collision_mode(int mode)
{
    g_err[0] = 0;
    if (mode != B200MP_COLLISION_AUTO && mode != B200MP_COLLISION_FP64_ONLY && mode != B200MP_COLLISION_SCREEN_ONLY) {
        set_error("set_collision_mode: unknown mode %d", mode);
        return B200MP_E_ARG;
    }
    return g_collision_mode.exchange(mode);
}

int b200mp_select_best_f64(int device, void *stream, int P, const double *ex, const double *ey,
                           const unsigned char *free_in, double gx, double gy, double weight, int norm_mode,
                           double *scores_out, int *best_out)
{
    B200MP_ENTER(device);
    return launch_select_best_f64(device, (cudaStream_t)stream, P, ex, ey, free_in, gx, gy, weight, norm_mode, scores_out,
                                  best_out);
}

int b200mp_track_closed_loop_f64(int device, void *stream, const B200mpTrackArgs *args)
{
    B200MP_ENTER(device);
    if (!args) {
        set_error("track: args is NULL");
        return B200MP_E_ARG;
    }
    return launch_track_f64(device, (cudaStream_t)stream, *args);
}

int b200mp_sample_lattice_f64(int device, void *stream, int P, int n_samples, const double *kappa1,
                              const double *kappa2, const double *sf, const double *ego_x, const double *ego_y,
                              const double *ego_yaw, int ego_broadcast, double *px, double *py, double *pyaw,
                              double *pcos, double *psin, double *end_xy)
{
    B200MP_ENTER(device);
    return launch_lattice_f64(device, (cudaStream_t)stream, P, n_samples, kappa1, kappa2, sf, ego_x, ego_y, ego_yaw,
                              ego_broadcast, px, py, pyaw, pcos, psin, end_xy);
}

int b200mp_optimize_spirals_f64(int device, void *stream, int P, int n_samples, const double *xf, const double *yf,
                                const double *tf, double *p_out, double *f_out, int *iters_out, unsigned char *valid_out)
{
    B200MP_ENTER(device);
    return launch_spiral_opt_f64(device, (cudaStream_t)stream, P, n_samples, xf, yf, tf, p_out, f_out, iters_out, valid_out);
}

int b200mp_fma_peak(int device, int dtype_bits, int reps, double *tflops_out)
{
    B200MP_ENTER(device);
    return run_fma_peak(dtype_bits, reps, tflops_out);
}

int b200mp_shutdown(void)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    int prev = -1;
    (void)cudaGetDevice(&prev);
    for (int d = 0; d < n && d < kMaxDevices; ++d) {
        DeviceState &ds = g_states[d];
        bool any_scratch = false;
        for (auto &sc : ds.scratch) any_scratch = any_scratch || sc.ptr;
        if (!ds.table64 && !ds.table32 && !any_scratch && !ds.sched_ring && !ds.mu_table && !ds.set_tables) continue;
        if (cudaSetDevice(d) != cudaSuccess) continue;
        (void)cudaDeviceSynchronize();
        if (ds.table64) (void)cudaFree(ds.table64);
        if (ds.table32) (void)cudaFree(ds.table32);
        for (auto &sc : ds.scratch)
            if (sc.ptr) (void)cudaFree(sc.ptr);
        if (ds.mu_table) (void)cudaFree(ds.mu_table);
        if (ds.mu_table_f32) (void)cudaFree(ds.mu_table_f32);
        if (ds.set_tables) (void)cudaFree(ds.set_tables);
        if (ds.set_B2) (void)cudaFree(ds.set_B2);
        if (ds.set_tables_f32) (void)cudaFree(ds.set_tables_f32);
        if (ds.sched_ring) {
            (void)cudaFree(ds.sched_ring);
            for (int i = 0; i < kSchedSlots; ++i) (void)cudaEventDestroy(ds.sched_event[i]);
        }
        ds = DeviceState{};
    }
    if (prev >= 0) (void)cudaSetDevice(prev);
    return 0;
}

}  // extern "C"
