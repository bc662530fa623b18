/* b200mp.h -- C ABI of the B200-native batched vehicle-dynamics / lattice-evaluation engine.
 *
 * This is the drop-in boundary for the one data-parallel hot path of earasteh/Python-Motionplanning:
 *   VehicleModel.planar_model / planar_model_RK4   (reference libs/vehicle_model/vehicle_model.py:220-445)
 *   CollisionChecker.collision_check               (reference libs/motionplanner/collision_checker.py:32-117)
 *   CollisionChecker.select_best_path_index        (reference libs/motionplanner/collision_checker.py:134-203)
 *   the multiprocessing fan-out they are called through (reference libs/motionplanner/local_planner.py:369-374)
 * The reference is pure Python and has no FFI of its own; these are the entry points a ctypes binding
 * on the reference side calls (INTEGRATION.md shows that binding).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes.  No C++ or torch types cross this boundary.
 *   - "dev" pointers are device memory on `device`, allocated and owned by the CALLER (cudaMalloc or a
 *     framework tensor); the library never frees them and keeps no reference after the call returns.
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream)
 *     unless stated otherwise.  Host scalars / small host arrays are consumed before the call returns.
 *   - Return value: 0 = OK, > 0 = a cudaError_t, < 0 = argument error (B200MP_E_*).  A message for the
 *     calling thread is available from b200mp_last_error().  No exception crosses the ABI.
 *   - Numerical exceptional values (NaN/Inf from a wheel speed of exactly 0) are not errors; they
 *     propagate exactly like the reference's numpy scalars.
 *   - Arrays are structure-of-arrays with the rollout / path index fastest, so that a warp touches
 *     32 consecutive elements per component.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef B200MP_H
#define B200MP_H

#ifdef __cplusplus
extern "C" {
#endif

#define B200MP_VERSION 100

#define B200MP_E_ARG (-1)      /* bad argument (NULL pointer, negative size, unsupported channel count) */
#define B200MP_E_PARAMS (-2)   /* no parameter table uploaded for this device / set index out of range */
#define B200MP_E_NODEVICE (-3) /* no usable CUDA device */

#define B200MP_N_STATE 10   /* [U, V, wz, wFL, wFR, wRL, wRR, yaw, x, y]            vehicle_model.py:224 */
#define B200MP_N_CARRY 12   /* the 10 states + ax_prev, ay_prev carried step to step  drive.py:141        */
#define B200MP_N_OUTPUTS 18 /* [fx(4) fy(4) Fz(4) s(4) fxtFL fytFL]                   vehicle_model.py:420 */
#define B200MP_N_AUX 28     /* RK4-averaged state_dot(10) + outputs(18)               vehicle_model.py:440-441 */
#define B200MP_N_MISC 6     /* [vx, vy, ax, ay, axc, ayc]                             vehicle_model.py:410-416 */

/* np.linalg.norm([v0, v1]) closed forms (host-BLAS dependent) */
#define B200MP_NORM2_NOFMA 0  /* sqrt(fl(v0*v0) + fl(v1*v1))    */
#define B200MP_NORM2_FMA_V1 1 /* sqrt(fma(v1, v1, fl(v0*v0)))    */
#define B200MP_NORM2_FMA_V0 2 /* sqrt(fma(v0, v0, fl(v1*v1)))    */

/* Parameter block: the attributes of the reference's VehicleParameters that the planar model reads
 * (vehicle_model.py:17-61).  Wheel order FL, FR, RL, RR.  D is the Pacejka peak used when no mu_max
 * array is supplied (the reference overwrites D with mu_max on every call, vehicle_model.py:232-235). */
typedef struct B200mpVehicleParams {
    double m, a, b, Izz, Jw, hg, T, wL, wR, rw;
    double B[4], C[4], D[4];
} B200mpVehicleParams;

/* Arguments of one batched RK4 rollout launch (replaces the per-vehicle loop of drive.py:141-143
 * around VehicleModel.planar_model_RK4, vehicle_model.py:427-445).  All pointers are dev pointers of
 * the launch's element type (double for _f64, float for _f32).
 *   state0     [12][B]   states + ax_prev, ay_prev
 *   delta      [n_seg][delta_ch][Bc]   delta_ch = 1: front steer on FL = FR, rear 0 (drive.py:143); 4: per wheel
 *   torque     [n_seg][torque_ch][Bc]  torque_ch = 1: equal on 4 wheels (stanley_controller.py:159); 4: per wheel
 *              Bc = B, or 1 when ctrl_broadcast != 0 (one control sequence for every rollout)
 *              segment index of step n = (step0 + n) / hold   (zero-order hold, drive.py:128)
 *   mu         [4][B] mu_max per wheel, or NULL -> parameter-set D
 *   param_set  [B] int index into the table of b200mp_set_params, or NULL -> set 0
 *   traj       [n_steps/store_stride][10][B] state after every store_stride-th step, or NULL (store_stride 0)
 *   aux        [n_steps/store_stride][28][B] RK4-averaged state_dot + outputs at the same steps, or NULL
 *   state_end  [12][B]  (may alias state0)
 *   cost       [B] or NULL: J = sum_n (x-xr_n)^2 + (y-yr_n)^2 + w_u (U-u_ref)^2 accumulated in step order
 *   cost_ref   [n_total][2] (xr, yr) indexed by step0 + n;  cost_in [B] or NULL = J carried in from a previous launch
 */
typedef struct B200mpRolloutArgs {
    int B;
    int n_steps;
    int step0;
    int hold;
    double dt;
    const void *state0;
    const void *delta;
    const void *torque;
    int delta_ch;
    int torque_ch;
    int ctrl_broadcast;
    int store_stride;
    const void *mu;
    const int *param_set;
    void *traj;
    void *aux;
    void *state_end;
    void *cost;
    const void *cost_in;
    const void *cost_ref;
    double w_u;
    double u_ref;
    int state_broadcast;   /* 1: state0 is [12][1], one start state for every rollout (sampling MPC: no [12][B] copy) */
    int friction_override; /* 0: the process-wide mode of b200mp_set_friction_mode; else 1 + B200MP_FRICTION_* for THIS launch */
} B200mpRolloutArgs;

int b200mp_version(void);
const char *b200mp_last_error(void);

/* Number of CUDA devices, or a negative error. */
int b200mp_device_count(void);

/* Upload n_sets parameter sets for `device` (synchronous; must not race with rollouts in flight on
 * that device) and build their friction tables (see b200mp_set_friction_mode; ~25 ms of host work per set, spread
 * over the host threads).  Replaces constructing VehicleParameters (vehicle_model.py:17-61, drive.py:37). */
int b200mp_set_params(int device, const B200mpVehicleParams *host_sets, int n_sets);

/* Batched RK4 rollouts, one thread per rollout.  vehicle_model.py:427-445 looped as drive.py:141-143. */
int b200mp_rk4_rollout_f64(int device, void *stream, const B200mpRolloutArgs *args);
int b200mp_rk4_rollout_f32(int device, void *stream, const B200mpRolloutArgs *args);

/* Batched single right-hand-side evaluation, VehicleModel.planar_model (vehicle_model.py:220-425).
 *   state [10][B], torque [4][B], mu [4][B] or NULL, delta [4][B], axay [2][B] (ax_prev, ay_prev),
 *   outputs: state_dot [10][B], misc [6][B] = (vx, vy, ax, ay, axc, ayc), outputs [18][B] (each may be NULL). */
int b200mp_planar_model_f64(int device, void *stream, int B, const double *state, const double *torque,
                            const double *mu, const double *delta, const double *axay, const int *param_set,
                            double *state_dot, double *misc, double *outputs);

/* Sampling-MPC control sequences (BASELINE.json config 4):
 *   delta[seg][0][r]  = clip(delta_mean + delta_sigma * eps,  +-delta_clip)
 *   torque[seg][0][r] = torque_mean + torque_sigma * eps'
 * eps, eps' iid N(0,1) from Philox4x32-10 keyed by seed with counter (rollout0 + r, seg): a shard that
 * passes its global offset in rollout0 reproduces the unsharded sequences. */
int b200mp_mpc_sample_controls_f64(int device, void *stream, int B, int n_seg, unsigned long long seed,
                                   long long rollout0, double delta_mean, double delta_sigma, double delta_clip,
                                   double torque_mean, double torque_sigma, double *delta, double *torque);

/* argmin over cost[n] with lowest-index tie-break (the convention of collision_checker.py:199); NaN
 * counts as +inf.  Writes min_out[0] (dev double) and idx_out[0] (dev long long, index_offset added;
 * -1 when nothing is finite). */
int b200mp_argmin_f64(int device, void *stream, long long n, const double *cost, long long index_offset,
                      double *min_out, long long *idx_out);

/* Winner record of a sampling-MPC plan in ONE call: lowest-index argmin over cost[B] (NaN = +inf) and a gather of the
 * winner's control sequence.  record dev [2 + 2 n_seg] = [min cost, global index = local index + index_offset (as a
 * double, exact below 2^53; -1 when nothing is finite), delta[n_seg], torque[n_seg]] with delta / torque dev
 * [n_seg][1][B] as written by b200mp_mpc_sample_controls_f64.  The records of the ranks are what one all-gather
 * exchanges (python_motionplanning_b200/distributed.py). */
int b200mp_mpc_winner_f64(int device, void *stream, long long B, int n_seg, const double *cost, const double *delta,
                          const double *torque, long long index_offset, double *record);

/* Circle-offset collision test for P paths at once: replaces Pool.starmap(collision_check, ...)
 * (local_planner.py:369-372) over CollisionChecker.collision_check (collision_checker.py:32-117).
 *   off, rad   HOST arrays [n_circ] (n_circ <= 8)
 *   px, py     dev [P][n_pts]
 *   pcos, psin dev [P][n_pts] cos / sin of path[2][j] computed by the caller on the host (numpy), which is
 *              what makes the booleans and min_clear bit-exact; or NULL with pyaw dev [P][yaw_stride] (first
 *              n_pts of each row used) to evaluate sincos on the device (<= 2 ulp: a verdict can then differ from the
 *              reference's only for an obstacle point within ~1e-13 m of a circle -- use b200mp_collision_check_yaw_f64
 *              for proven verdicts)
 *   obs        dev [M][2]
 *   free_out   dev [P] bytes, 1 = collision-free (the reference's polarity)
 *   min_clear  dev [P] min over all tests of (distance - radius), or NULL
 * Distances follow scipy cdist bit for bit: sqrt(fl(dx*dx) + fl(dy*dy)), no FMA; d == r is free. */
int b200mp_collision_check_f64(int device, void *stream, int P, int n_pts, int n_circ, const double *off,
                               const double *rad, const double *px, const double *py, const double *pcos,
                               const double *psin, const double *pyaw, int yaw_stride, int M, const double *obs,
                               unsigned char *free_out, double *min_clear);

/* The same test from path YAWS, bit-exact without host trigonometry on the bulk of the data (collision_checker.py:88-89
 * evaluates np.cos / np.sin of every yaw; here the device does).  Every verdict the device writes is PROVEN to equal the
 * verdict of the reference's arithmetic for any host cos / sin within 2^-44 of the exact values (numpy: <= 0.52 ulp;
 * CUDA sincos: <= 2 ulp): the host's and the device's circle centres then differ by at most
 *   ec = 2 |off| (2^-44 + 2^-50) + 2^-51 (|cx| + |cy|)   (~1e-13 m),
 * the FP32 screen's band is widened by ec, and the exact FP64 recheck declares a collision only for
 * q64 < ((r (1 - 2^-50) - ec) / (1 + 2^-51))^2 and "free" only for q64 >= ((r (1 + 2^-50) + ec) / (1 - 2^-51))^2.
 * A path point with an obstacle point inside that band is NOT decided on the device: its item index p * n_pts + j is
 * appended to `undecided` (dev int [1 + capacity]: [0] = number of appended items, which may exceed capacity and may
 * contain duplicates; [1..] = items) and free_out[p] is left as the other points of the path decide it.  The caller reads
 * the count, evaluates cos / sin of the listed yaws on the HOST (numpy) and calls b200mp_collision_resolve_f64; if the
 * count exceeds capacity it falls back to b200mp_collision_check_f64 with host pcos / psin.  On BASELINE config 3 the
 * list is empty; a list entry needs an obstacle point within ~1e-13 m of a circle.
 *   pyaw  dev [P][yaw_stride], the first n_pts of each row are used (the reference's 49/50 quirk)
 *   mode  B200MP_COLLISION_* for THIS call, or -1 for the process-wide default */
int b200mp_collision_check_yaw_f64(int device, void *stream, int P, int n_pts, int n_circ, const double *off,
                                   const double *rad, const double *px, const double *py, const double *pyaw,
                                   int yaw_stride, int M, const double *obs, unsigned char *free_out, int *undecided,
                                   int undecided_capacity, int mode);

/* Decides the listed path points exactly: items dev [n_list] (as written by b200mp_collision_check_yaw_f64), cos_sin dev
 * [2][n_list] = the caller's host-evaluated cos and sin of the items' yaws; every obstacle point is tested against the
 * items' circles with the FP64 sequence of b200mp_collision_check_f64 and free_out[p] is cleared on a collision. */
int b200mp_collision_resolve_f64(int device, void *stream, int n_list, const int *items, const double *cos_sin, int P,
                                 int n_pts, int n_circ, const double *off, const double *rad, const double *px,
                                 const double *py, int M, const double *obs, unsigned char *free_out);

/* How the rollout / tracking kernels evaluate the combined-slip friction D sin(C atan(B s)) / s
 * (vehicle_model.py:296-348).  AUTO (default): from tables of piecewise polynomials in 1 + (B s)^2 built on the host by
 * b200mp_set_params (long double fits, audited: FP64 <= 4e-16 relative with degree 6 on 64 intervals per binade, FP32 a
 * few ulps with cubics on 32), no square root / reciprocal / atan / sin in the kernel; slips beyond the table and
 * non-finite values repeat the step on the closed form.  Set 0 has a table with D folded in (launches without param_set
 * and mu); every set whose four tyres share (B, C) has a D = 1 table that the generic kernels use for blocks of 64
 * rollouts sharing one set, with D (or mu_max) scaling the normal load; a set that fails the audit (e.g. C > 2, where
 * the function changes sign) and logging steps are evaluated in closed form.  CLOSED_FORM: always the sqrt / atan / sin
 * sequence (A/B checks).  Both are within the 1e-9 parity contract with ~1e-13 to spare.  Process-wide; returns the
 * previous mode, or B200MP_E_ARG. */
#define B200MP_FRICTION_AUTO 0
#define B200MP_FRICTION_CLOSED_FORM 1
int b200mp_set_friction_mode(int mode);

/* Arithmetic of the boolean-only collision test (min_clear == NULL).  AUTO (default): every circle/point
 * pair is screened in FP32 with a proven error bound and only undecided pairs repeat the exact FP64
 * sequence above -- the flags are bit-identical to FP64_ONLY for every input, the FP64 pipe is simply not
 * spent on pairs single precision already decides.  FP64_ONLY forces the all-FP64 kernel (A/B checks).
 * Process-wide; returns the previous mode, or B200MP_E_ARG. */
#define B200MP_COLLISION_AUTO 0        /* bounding-box broad phase over 32-point obstacle chunks + FP32 screen + exact recheck */
#define B200MP_COLLISION_FP64_ONLY 1   /* every pair in FP64 */
#define B200MP_COLLISION_SCREEN_ONLY 2 /* FP32 screen + exact recheck on every pair (no broad phase) */
int b200mp_set_collision_mode(int mode);

/* Broad-phase statistics of the LAST b200mp_collision_check_f64 call on `device` in AUTO mode with M obstacle points
 * (synchronises `stream`): out2[0] = (warp, 32-point chunk) pairs that survived the bounding-box test and were
 * screened point by point (x 32 path points x n_circ x 32 = circle/point tests executed), out2[1] = (thread, chunk)
 * pairs repeated in FP64.  For honest "executed vs nominal" accounting in benchmarks. */
int b200mp_collision_stats(int device, void *stream, int M, unsigned long long *out2);

/* select_best_path_index on the path end points (collision_checker.py:134-203):
 *   score_i = norm([ex_i-gx, ey_i-gy]) + sum over colliding j (ascending) of weight*norm([ex_i-ex_j, ey_i-ey_j])
 * colliding i -> +inf; first strict minimum wins; best_out[0] = -1 for the reference's None.
 * free_in dev [P] bytes: 1 = collision-free (a candidate), 0 = colliding (a penalty term for the others), any other value
 * = excluded -- neither candidate nor penalty, which is how a path the planner dropped before the selection
 * (local_planner.py:317-323: spiral end point farther than 0.1 from its goal) is passed without compacting the arrays;
 * the relative order of the remaining paths, and with it the tie-break, is the filtered list's.
 * norm_mode picks the closed form of np.linalg.norm the host follows (B200MP_NORM2_*).
 * scores_out dev [P] or NULL. */
int b200mp_select_best_f64(int device, void *stream, int P, const double *ex, const double *ey,
                           const unsigned char *free_in, double gx, double gy, double weight, int norm_mode,
                           double *scores_out, int *best_out);

/* Batched closed-loop path tracking: for every vehicle, every ctrl_every-th sub-step,
 * StanleyController.stanley_control (reference libs/controllers/stanley_controller.py:56-129) and
 * LongitudinalController.long_control (:138-159) on the current state, the first-order steering filter of
 * drive.py:137-138, then VehicleModel.planar_model_RK4 with delta = [d, d, 0, 0], torque = [tau]*4
 * (drive.py:141-143; the parameter set's D plays mu_max = [1,1,1,1] of drive.py:142).
 *   state0     dev [12][V]
 *   ctrl0      dev [3][V]   steering-filter state x_del, integral of the speed error, previous speed
 *                           (drive.py: x_del[-1], total_vel_error, prev_vel)
 *   waypoints  dev [n_sets][w_max][2] (16-byte aligned), wp_count dev [n_sets]; vehicle r tracks set
 *              r / vehicles_per_set (the planner hands one waypoint list to each tracker, local_planner.py:419)
 *   traj       dev [n_steps/store_stride][10][V] or NULL
 *   log        dev [n_steps/store_stride][45][V] or NULL: the DataLog row of drive.py:145-151
 *              [time, state(10), state_dot(10), delta, torque(4), outputs(18), crosstrack]; time = (step0+n)*dt
 *   target_idx dev [ceil(n_steps/ctrl_every)][V] or NULL: the look-ahead waypoint index of every control update
 *   state_end  dev [12][V], ctrl_end dev [3][V] (may alias the inputs): the loop is resumable
 * step0 must be a multiple of ctrl_every (a launch starts on a control update).  norm_mode as in
 * b200mp_select_best_f64.  Target indices and crosstrack errors follow the reference bit for bit given the
 * same states; headings use CUDA atan2/sincos (<= 1-2 ulp from numpy's). */
typedef struct B200mpTrackArgs {
    int V;
    int n_steps;
    int step0;
    int ctrl_every;
    int store_stride;
    int n_sets;
    int w_max;
    int vehicles_per_set;
    int norm_mode;
    double dt;
    double target_vel;
    double k, k_soft, max_steer;       /* StanleyController gains (drive.py:71-74) */
    double kp, ki, kd;                 /* LongitudinalController gains (drive.py:82-84) */
    double lookahead, deadband;        /* stanley_controller.py:44-45 */
    double steer_filter;               /* 1e-5 / (2 * 0.001), drive.py:137 */
    const double *state0;
    const double *ctrl0;
    const double *waypoints;
    const int *wp_count;
    double *traj;
    double *log;
    int *target_idx;
    double *state_end;
    double *ctrl_end;
    int friction_override; /* 0: the process-wide friction mode; else 1 + B200MP_FRICTION_* for THIS launch */
} B200mpTrackArgs;
#define B200MP_N_LOG 45 /* columns of the reference's DataLog (drive.py:44, plots.py:19-27) */
int b200mp_track_closed_loop_f64(int device, void *stream, const B200mpTrackArgs *args);

/* Conformal-lattice path generation: PathOptimizer.sample_spiral + thetaf (reference
 * libs/motionplanner/path_optimizer.py:109-174) followed by transform_paths (libs/motionplanner/local_planner.py:
 * 424-470) for P spirals at once.
 *   kappa1, kappa2, sf   dev [P]: the optimisation parameters p = [p1, p2, sf] of each spiral
 *   ego_x, ego_y, ego_yaw dev [P] (or [1] with ego_broadcast != 0), or all NULL for ego-frame output
 *   px, py               dev [P][n_samples-1] path points (cumulative trapezoid has no initial value)
 *   pyaw                 dev [P][n_samples-1] or NULL: heading of the sample BEFORE each point + ego_yaw (the
 *                        reference's 49/50 length quirk, kept because collision_check reads yaw_j with point_j)
 *   pcos, psin           dev [P][n_samples-1] or NULL: cos / sin of pyaw evaluated on the device, ready for
 *                        b200mp_collision_check_f64 (<= 1-2 ulp from numpy's: flags then agree with the host-trig
 *                        path except for obstacle points within ~1e-15 of a circle)
 *   end_xy               dev [2][P] or NULL: (x[-1], y[-1]) of every path for b200mp_select_best_f64
 * n_samples = 50 is the reference's (np.linspace default).  Parity: 1e-12 relative (CUDA vs numpy cos/sin/pow). */
int b200mp_sample_lattice_f64(int device, void *stream, int P, int n_samples, const double *kappa1,
                              const double *kappa2, const double *sf, const double *ego_x, const double *ego_y,
                              const double *ego_yaw, int ego_broadcast, double *px, double *py, double *pyaw,
                              double *pcos, double *psin, double *end_xy);

/* Batched cubic-spiral optimisation: for P goal states (xf, yf, tf) in the vehicle frame the
 * problem PathOptimizer.optimize_spiral poses (reference libs/motionplanner/path_optimizer.py:31-88, objective and
 * gradient :183-530): minimise fbe + 25 (fxf + fyf) + 30 ftf over p = [p1, p2, sf] from [0, 0, |goal|] within
 * p1, p2 in [-0.5, 0.5], sf >= |goal|.  The reference solves it with scipy's L-BFGS-B (not part of its sources); this
 * is a projected Levenberg-Marquardt iteration to a 1e-13 projected-gradient tolerance: the same minimiser, resolved
 * further than L-BFGS-B stops (parity: parameters within ~1e-6, objective never above the reference's result).
 *   xf, yf, tf  dev [P]
 *   p_out       dev [3][P]  (p1, p2, sf): the kappa1 / kappa2 / sf inputs of b200mp_sample_lattice_f64
 *   f_out       dev [P] objective at p_out, or NULL;  iters_out dev [P] int, or NULL
 *   valid_out   dev [P] bytes or NULL: the planner's acceptance test on the spiral sampled with n_samples points,
 *               norm([x_end - xf, y_end - yf, t_end - tf]) <= 0.1 (local_planner.py:317-323) */
int b200mp_optimize_spirals_f64(int device, void *stream, int P, int n_samples, const double *xf, const double *yf,
                                const double *tf, double *p_out, double *f_out, int *iters_out,
                                unsigned char *valid_out);

/* Measured pipe peak for the roofline denominator: runs a register-resident FMA chain kernel
 * (dtype_bits 64 or 32) `reps` times, synchronises, and returns the best TFLOP/s (FMA = 2 flop). */
int b200mp_fma_peak(int device, int dtype_bits, int reps, double *tflops_out);

/* Free library-owned device state (parameter tables, reduction scratch) on every device. */
int b200mp_shutdown(void);

#ifdef __cplusplus
}
#endif
#endif /* B200MP_H */
