// Internal glue shared by the translation units of libb200mp.so (not part of the ABI).
#pragma once

#include <cuda_runtime.h>

#include "../../include/b200mp.h"
#include "vehicle_rhs.cuh"

namespace b200mp {

constexpr int kMaxDevices = 64;
constexpr int kSchedSlots = 32;
constexpr size_t kSchedSlotBytes = 64 * 1024;   // counter + one int per rollout block of a sliced launch

// printf-style message for b200mp_last_error() (thread-local)
void set_error(const char *fmt, ...);
// records the CUDA error text and returns it as the ABI return code
int cuda_fail(cudaError_t e, const char *what);

#define B200MP_CUDA(call)                                         \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return ::b200mp::cuda_fail(e__, #call); \
    } while (0)

// Library-owned per-device state: the uploaded parameter table and reduction scratch.
struct DeviceState {
    int n_sets = 0;
    HostParams set0{};
    DevParams<double> *table64 = nullptr;
    DevParams<float> *table32 = nullptr;
    // tabulated friction function of parameter set 0 (vehicle_rhs.cuh: build_mu_table), when its four tyres are equal
    double *mu_table = nullptr;
    double mu_table_B2 = 0.0;
    double mu_table_err = 0.0;
    float *mu_table_f32 = nullptr;   // FP32 twin (cubic per interval, vehicle_rhs.cuh: build_mu_table_f32); valid iff mu_table_f32_ok
    bool mu_table_f32_ok = false;
    // the same tables normalised to D = 1 for EVERY parameter set whose four tyres share (B, C): [n_sets][kMuTableDoubles],
    // and B^2 per set (0 = no table: the set is evaluated in closed form).  Used by the generic kernels when all
    // rollouts of a CTA share one set (parameter sweeps, per-rollout mu_max).
    double *set_tables = nullptr;
    float *set_tables_f32 = nullptr;   // FP32 twins [n_sets][kMuTableFloats]; a set has one iff set_B2 > 0 (both are built or neither)
    double *set_B2 = nullptr;
    int set_tables_n = 0;
    // scratch areas, ONE PER STREAM that has asked for one (the ABI is asynchronous on the caller's stream, so two calls
    // on different streams of one device must not share an area while their kernels are in flight); work on one stream
    // is ordered, so an area is reused by the next call on the same stream without synchronisation
    static constexpr int kScratchStreams = 8;
    struct Scratch {
        cudaStream_t stream = nullptr;
        void *ptr = nullptr;
        size_t bytes = 0;
        unsigned long long last_use = 0;
        bool used = false;
    };
    Scratch scratch[kScratchStreams];
    unsigned long long scratch_clock = 0;
    // where the broad-phase statistics of the last collision launch sit; cleared by the next scratch user of that stream
    void *cull_stats_ptr = nullptr;
    cudaStream_t cull_stats_stream = nullptr;
    // ring of small work-queue areas for time-sliced rollout launches (one per launch in flight)
    void *sched_ring = nullptr;
    cudaEvent_t sched_event[kSchedSlots] = {};
    bool sched_used[kSchedSlots] = {};
    int sched_next = 0;
};
DeviceState &dev_state(int device);
// a device area of at least `bytes` owned by (device, stream) until the next ensure_scratch on that stream
int ensure_scratch(int device, cudaStream_t stream, size_t bytes, void **out);
// A zero-initialisable kSchedSlotBytes device area that no launch still in flight is using; the caller
// records `*done` on its stream after the launch that uses the area.
int acquire_sched_slot(int device, void **area, cudaEvent_t *done);

// B200MP_FRICTION_* (process-wide, b200mp_set_friction_mode)
int friction_mode();
// B200MP_COLLISION_* (process-wide, b200mp_set_collision_mode)
int collision_mode();

// Sets `device` current for the scope of an ABI call and restores the caller's device afterwards.
class DeviceGuard {
  public:
    explicit DeviceGuard(int device);
    ~DeviceGuard();
    int status() const { return status_; }

  private:
    int prev_ = -1;
    int status_ = 0;
};

// kernel launchers (defined next to their kernels)
int launch_rollout_f64(int device, cudaStream_t st, const B200mpRolloutArgs &a);
int launch_rollout_f32(int device, cudaStream_t st, const B200mpRolloutArgs &a);
int launch_planar_model_f64(int device, cudaStream_t st, int B, const double *state, const double *torque,
                            const double *mu, const double *delta, const double *axay, const int *param_set,
                            double *state_dot, double *misc, double *outputs);
int launch_collision_f64(int device, cudaStream_t st, int P, int n_pts, int n_circ, const double *off,
                         const double *rad, const double *px, const double *py, const double *pcos,
                         const double *psin, const double *pyaw, int yaw_stride, int M, const double *obs,
                         unsigned char *free_out, double *min_clear, int *undecided, int undecided_capacity, int mode);
int launch_collision_resolve_f64(int device, cudaStream_t st, int n_list, const int *items, const double *cos_sin, int P,
                                 int n_pts, int n_circ, const double *off, const double *rad, const double *px,
                                 const double *py, int M, const double *obs, unsigned char *free_out);
int collision_stats(int device, cudaStream_t st, int M, unsigned long long out[2]);
int launch_select_best_f64(int device, cudaStream_t st, int P, const double *ex, const double *ey,
                           const unsigned char *free_in, double gx, double gy, double weight, int norm_mode,
                           double *scores_out, int *best_out);
int launch_argmin_f64(int device, cudaStream_t st, long long n, const double *cost, long long index_offset,
                      double *min_out, long long *idx_out);
// two-phase argmin (lowest index wins ties, NaN = +inf); scratch must hold argmin_scratch_bytes(n)
size_t argmin_scratch_bytes(long long n);
int argmin_launch(cudaStream_t st, long long n, const double *cost, long long index_offset, void *scratch,
                  double *min_out, long long *idx_out, int *idx32_out);
int launch_mpc_sample_f64(cudaStream_t st, int B, int n_seg, unsigned long long seed, long long rollout0,
                          double delta_mean, double delta_sigma, double delta_clip, double torque_mean,
                          double torque_sigma, double *delta, double *torque);
int launch_mpc_winner_f64(int device, cudaStream_t st, long long B, int n_seg, const double *cost, const double *delta,
                          const double *torque, long long index_offset, double *record);
int run_fma_peak(int dtype_bits, int reps, double *tflops_out);
int launch_lattice_f64(int device, cudaStream_t st, int P, int n_samples, const double *k1, const double *k2,
                       const double *sf, const double *ego_x, const double *ego_y, const double *ego_yaw,
                       int ego_broadcast, double *px, double *py, double *pyaw, double *pcos, double *psin, double *end_xy);
int launch_spiral_opt_f64(int device, cudaStream_t st, int P, int n_samples, const double *xf, const double *yf,
                          const double *tf, double *p_out, double *f_out, int *it_out, unsigned char *valid_out);
int launch_track_f64(int device, cudaStream_t st, const B200mpTrackArgs &a);

}  // namespace b200mp
