/*
 * mppi_b200.h -- C ABI of the B200-native MPPI controller core (libmppi_b200.so).
 *
 * Drop-in boundary for the optimisation step of the three reference nodes.  The reference has no FFI: the
 * solve is four private member functions sharing member arrays, called once per control cycle from run().
 * Each entry point below names the reference code it replaces (paths relative to /root/reference):
 *   DD  = src/diff_drive_mppi.cpp            DDh = include/ccv_mppi_path_tracker/diff_drive_mppi.h
 *   SD  = src/steering_diff_drive_mppi.cpp   SDh = include/ccv_mppi_path_tracker/steering_diff_drive_mppi.h
 *   FB  = src/full_body_mppi.cpp             FBh = include/ccv_mppi_path_tracker/full_body_mppi.h
 *
 * Conventions: plain pointers and sizes only; every function returns MPPI_OK (0) or a negative mppi_status;
 * no C++ exception crosses this boundary; mppi_last_error() gives the text of the last failure on the handle.
 * A handle is bound to one CUDA device and is not thread-safe (the reference is single threaded, DD:336-368).
 * There is NO CPU fallback: every compute entry point fails with MPPI_ERR_CUDA when no device is usable.
 */
#ifndef MPPI_B200_H
#define MPPI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPPI_B200_ABI_VERSION 1

typedef enum {
  MPPI_OK = 0,
  MPPI_ERR_INVALID = -1,   /* bad argument */
  MPPI_ERR_CUDA = -2,      /* CUDA runtime failure / no device */
  MPPI_ERR_STATE = -3,     /* call order (e.g. solve before a path or window was set) */
  MPPI_ERR_NCCL = -4,      /* collective layer failure */
  MPPI_ERR_ALLOC = -5
} mppi_status;

typedef enum {
  MPPI_MODEL_DIFF_DRIVE = 0, /* class DiffDriveMPPI,          U = 2 (v, w)                              DDh:52 */
  MPPI_MODEL_STEERING = 1,   /* class SteeringDiffDriveMPPI,  U = 3 (v, w, steer)                        SDh:56 */
  MPPI_MODEL_FULL_BODY = 2   /* class FullBodyMPPI,           U = 5 (v, w, direction, roll_v, pitch_v)   FBh:68 */
} mppi_model;

/* The nodes' ROS-parameter surface that the solve reads (DD:17-34, SD:18-36, FB:8-46), as doubles like the
 * reference members (DDh:87-96, FBh:165-181).  horizon / num_samples are arguments of mppi_create. */
typedef struct {
  double control_noise; /* sigma of every control (DD:20) */
  double lambda;        /* DD:21 */
  double v_ref;         /* DD:28 */
  double resolution;    /* path resolution used by calc_RefPath (DD:29, DD:160) */
  double u_min[5];      /* v_min, w_min, steer_min, roll_v_min, pitch_v_min (DD:24-26, SD:26-28, FB:20-26) */
  double u_max[5];      /* v_max, w_max, steer_max, roll_v_max, pitch_v_max; u_min[k] <= u_max[k] is required */
  double path_weight;   /* DD:33 */
  double v_weight;      /* DD:34 (read from "control_weight"), FB:35 */
  double zmp_weight;    /* FB:36; caller applies roll_off (FB:43-46) */
  double roll_v_weight; /* FB:37 */
  double back_weight;   /* FB:38 */
  double yaw_weight;    /* FB:39 */
  int32_t steer_off;    /* FB:41, FB:517 */
  int32_t reserved;
} mppi_params;

typedef struct mppi_handle_s *mppi_handle;

/* Optional per-solve debug taps (all device->host copies happen only when requested). */
typedef enum {
  MPPI_DEBUG_NONE = 0,
  MPPI_DEBUG_NEAREST = 1, /* keep nearest window index per (sample, t): the implicit argmin of DD:186-190 */
  MPPI_DEBUG_STATES = 2   /* keep every predicted state: what publish_CandidatePath shows (DD:265-294) */
} mppi_debug_flags;

/* Implementation of the nearest-window-point scan inside the fused rollout+cost kernel. Both are exact and
 * bit-identical in every output (cost and, with MPPI_DEBUG_NEAREST, the argmin index: first minimum wins,
 * DD:186-190); LITERAL evaluates all T window points per state, PRUNED only the points that a per-solve candidate
 * grid cannot rule out.  AUTO = PRUNED when T >= 24, LITERAL below; LITERAL is forced while MPPI_DEBUG_STATES is on
 * (it is the kernel that records the predicted states). */
typedef enum {
  MPPI_SCAN_AUTO = 0,
  MPPI_SCAN_LITERAL = 1,
  MPPI_SCAN_PRUNED = 2
} mppi_scan_mode;

/* Where get_CurrentIndex + calc_RefPath (DD:126-181) run for robots that have a path: on the host inside
 * mppi_upload (FP64), or in a device kernel, one CTA per robot (FP64, same expressions; for many-robot handles).
 * AUTO = DEVICE when n_robots >= 8. */
typedef enum {
  MPPI_WINDOW_AUTO = 0,
  MPPI_WINDOW_HOST = 1,
  MPPI_WINDOW_DEVICE = 2
} mppi_window_builder;

/* Tuning and behaviour switches of one handle (mppi_set_option / mppi_get_option); values are doubles, integer
 * options take the integral value.  Every option has a working default: none is needed for a correct solve, and none
 * is read from the process environment. */
typedef enum {
  MPPI_OPT_GRID_MAX_CELLS = 1,      /* cells of the candidate grid per robot (pruned scan), 256 .. 262144; default
                                       clamp(num_samples / 2, 256, 65536).  Re-allocates the cell table. */
  MPPI_OPT_GRID_H_MIN = 2,          /* smallest grid cell side [m], > 0; default 0.05 */
  MPPI_OPT_GRID_MARGIN = 3,         /* margin around the window's bounding box [m]; 0 (default) = a quarter of the
                                       horizon reach v_ref * dt * (T-1) */
  MPPI_OPT_GRID_LANES = 4,          /* lanes per cell of the grid builder: 0 (default) = by the handle's cell count,
                                       or 1, 2, 4, 8, 16, 32 */
  MPPI_OPT_REDUCE_GROUPS = 5,       /* plane groups one block of the weighted-control reduction walks through;
                                       0 (default) = by the number of blocks */
  MPPI_OPT_FUSE_CONTROLS = 6,       /* weighted controls reduced per CTA inside the rollout kernel (1) or by the separate
                                       weight + reduction kernels (0); -1 (default) = by problem shape */
  MPPI_OPT_NOISE_PREFETCH = 7,      /* 1: the normals of solve n+1 are generated on a second stream while solve n runs
                                       (double-buffered tensor; same Philox stream, same results); 0: generated at the
                                       start of their own solve; -1 (default) = on when the pruned scan is used and
                                       the second tensor fits in device memory */
  MPPI_OPT_EXCHANGE_TIMEOUT_MS = 8, /* peer exchange (mppi_comm_connect): how long the merge waits for a peer's record
                                       before the solve fails with MPPI_ERR_NCCL; default 2000 */
  MPPI_OPT_FEEDBACK_WARM_START = 9, /* 1 (default): the new controls become the warm start of the next mppi_enqueue, as
                                       the reference's optimal_solution does (DD:89-90); 0: every mppi_enqueue starts
                                       from the warm start of the last mppi_upload (a host that shifts or resets the
                                       sequence itself; repeated identical solves) */
  MPPI_OPT_UPLOAD_WARM_START = 10,  /* 1 (default): mppi_solve reads u_nominal on entry (in/out, like the reference's
                                       optimal_solution member); 0: after the first solve the warm start is the
                                       device's own copy of the previous result and u_nominal is output only -- saves
                                       the conversion and the upload of n_robots x (T-1) x U values per cycle */
  MPPI_INFO_FUSED_CONTROLS = 100    /* read-only (mppi_get_option): 1 when the kernel sequence issued last reduced the
                                       weighted controls inside the rollout kernel (what MPPI_OPT_FUSE_CONTROLS = -1
                                       decided) */
} mppi_option;

/* ---- lifetime ------------------------------------------------------------------------------------------ */

/* Replaces the constructors' allocation of sample[K], optimal_solution, window and weights_ (DD:36-46,
 * SD:38-48, FB:72-84).  num_samples = samples owned by THIS handle (a shard when sharded), n_robots >= 1
 * independent controllers batched in one handle (each with its own state, path/window and warm start).
 * device = CUDA ordinal.  All device memory is allocated here; mppi_solve allocates nothing. */
int mppi_create(mppi_handle *out, int model, const mppi_params *params, int num_samples, int horizon,
                int n_robots, int device);
int mppi_destroy(mppi_handle h);
const char *mppi_last_error(mppi_handle h); /* h may be NULL: error of the last failed mppi_create */
int mppi_abi_version(void);

/* Re-read parameters that the reference reads once in the ctor (weights, limits, sigma, lambda). */
int mppi_set_params(mppi_handle h, const mppi_params *params);
int mppi_set_debug(mppi_handle h, int debug_flags);
int mppi_set_scan_mode(mppi_handle h, int scan_mode);
int mppi_set_window_builder(mppi_handle h, int mode);
/* MPPI_ERR_INVALID for an unknown option or a value outside the documented range (the handle is left unchanged). */
int mppi_set_option(mppi_handle h, int option, double value);
int mppi_get_option(mppi_handle h, int option, double *value);

/* ---- inputs -------------------------------------------------------------------------------------------- */

/* pathCallback (DD:48-52): full reference path of one robot, n_points x {x, y} doubles.  The window is
 * rebuilt from it on every solve by get_CurrentIndex + calc_RefPath (DD:126-181). */
int mppi_set_path(mppi_handle h, int robot, const double *path_xy, int n_points);

/* Alternative to mppi_set_path: give the T-point window x_ref_, y_ref_, yaw_ref_ (DD:42-44) directly,
 * T x {x, y, yaw} doubles; bypasses calc_RefPath for that robot until mppi_set_path is called again. */
int mppi_set_window(mppi_handle h, int robot, const double *window_xyyaw);

/* Noise source.  Default: internal Philox4x32-10 + Box-Muller stream (replaces std::mt19937 +
 * std::normal_distribution, DD:83-97); element (robot, t, u, global sample i) is a pure function of
 * (seed, solve counter, robot_offset + robot, t, u, sample_offset + i), so shards of one solve draw disjoint,
 * reproducible sub-streams.  sample_offset must be a multiple of 4. */
int mppi_set_seed(mppi_handle h, uint64_t seed, uint64_t first_solve_counter);
int mppi_set_shard(mppi_handle h, int64_t sample_offset, int64_t num_samples_global, int robot_offset);
/* Parity runs: use exactly this standard-normal tensor, logical order [robot][t][i][u] (the reference's draw
 * order, DD:86-100), float32, for every following solve; eps = NULL returns to the internal generator. */
int mppi_set_noise(mppi_handle h, const float *eps);

/* ---- the solve ----------------------------------------------------------------------------------------- */

/* One control cycle: sampling() -> predict_States() -> calc_Weights() -> determine_OptimalSolution()
 * (DD:352-358, SD:385-391, FB:638-644) for every robot of the handle.
 *   state      n_robots x S doubles: x, y, yaw (DD:115-117) [+ roll, pitch for FULL_BODY, FB:458-464]
 *   dt         dt_ of this cycle (DD:347)
 *   u_nominal  n_robots x (T-1) x U doubles, in: optimal_solution of the previous cycle (warm start, DD:89-90),
 *              out: the new optimal_solution (DD:230-235 with the t<T-1 bound)
 * Host buffers in, host buffers out; synchronous (returns after u_nominal is written).
 * MPPI_ERR_INVALID for dt <= 0 or a non-finite state (the reference would publish NaN commands, DD:117 -> DD:250). */
int mppi_solve(mppi_handle h, const double *state, double dt, double *u_nominal);

/* The same cycle split for pipelining and for device-resident timing:
 *   mppi_upload   host -> pinned -> device copy of state / window / warm start (async on the handle's stream)
 *   mppi_enqueue  the kernels (and the collective when sharded) on the handle's stream; the new controls stay
 *                 on the device and become the warm start of the next enqueue
 *   mppi_download device -> host copy of the current controls, then stream synchronise */
int mppi_upload(mppi_handle h, const double *state, double dt, const double *u_nominal);
int mppi_enqueue(mppi_handle h);
int mppi_download(mppi_handle h, double *u_nominal);
int mppi_synchronize(mppi_handle h);
/* Use a caller-owned cudaStream_t (e.g. torch's current stream) instead of the handle's own. NULL = default. */
int mppi_set_stream(mppi_handle h, void *cuda_stream);
/* Capture the solve into a CUDA graph and replay it on later calls: mppi_enqueue replays the kernels, mppi_solve
 * the H2D copy + kernels + D2H copy.  Works for unsharded handles and for the NVLink peer exchange
 * (mppi_comm_connect); handles on the NCCL transport (mppi_comm_init) keep plain stream launches. */
int mppi_use_graph(mppi_handle h, int enable);

/* ---- outputs beyond the controls (debug / parity) ------------------------------------------------------- */

int mppi_get_costs(mppi_handle h, int robot, float *cost /* [K] */);
int mppi_get_weights(mppi_handle h, int robot, float *weights /* [K], exp(-(c-c_min)/lambda), not normalised */);
int mppi_get_nearest(mppi_handle h, int robot, int32_t *nearest /* [K][T], needs MPPI_DEBUG_NEAREST */);
/* candidate trajectories sample[i].x_, y_, yaw_ [, roll_, pitch_] (DD:287-288), world frame; needs MPPI_DEBUG_STATES */
int mppi_get_states(mppi_handle h, int robot, double *states /* [K][T][S] */);
int mppi_get_noise(mppi_handle h, int robot, float *eps /* [T-1][K][U] */);
int mppi_get_window(mppi_handle h, int robot, double *window_xyyaw /* [T][3] */, int *current_index);
/* stats[0] = c_min, stats[1] = sum of shifted weights, stats[2] = effective sample size (global when sharded) */
int mppi_get_stats(mppi_handle h, int robot, double *stats);
int mppi_get_info(mppi_handle h, int *model, int *num_samples, int *horizon, int *num_controls, int *n_robots);
/* Per-kernel device time of the solve (CUDA events between the launches on the handle's stream), averaged over
 * n_iters solves after one warm-up: ms[0..5] = noise, rollout+cost, weights, weighted controls, finalize, merge
 * (collective time, when sharded, is included in ms[5]); ms[6] = whole enqueue; ms[7] = candidate grid of the
 * pruned scan (0 when the literal scan runs).  The kernels run one after the other here (no second stream), so the
 * sum exceeds the time of a pipelined mppi_enqueue.  Side effects: runs n_iters + 1 REAL solves -- the warm start
 * and the solve counter advance exactly as by n_iters + 1 calls of mppi_enqueue. */
int mppi_time_kernels(mppi_handle h, int n_iters, float *ms /* [8] */);
/* bytes the last mppi_solve / mppi_upload moved host -> device (header, poses, state records; the warm start when the
 * host supplies it; the windows when they are built on the host) and one solve moves device -> host (controls, stats) */
int mppi_get_io_bytes(mppi_handle h, size_t *h2d_bytes, size_t *d2h_bytes);
/* number of kernel launches issued by the last mppi_enqueue (bench.py's gpu_launches claim) */
int mppi_last_launch_count(mppi_handle h);

/* ---- multi-GPU: samples sharded over ranks, one exchange of (c_min, sum w, sum w*u) per solve ----------- */

#define MPPI_COMM_ID_BYTES 128
/* rank 0 creates the id, the host launcher broadcasts it (torch.distributed / MPI / a file) */
int mppi_comm_get_unique_id(void *id_out /* MPPI_COMM_ID_BYTES */);
int mppi_comm_init(mppi_handle h, const void *id, int rank, int n_ranks);
/* The same exchange without NCCL, over NVLink peer memory: every rank exports an exchange buffer (CUDA IPC handle,
 * MPPI_IPC_HANDLE_BYTES), the launcher all-gathers the handles, every rank connects.  Afterwards the last kernel
 * of a solve stores this rank's record straight into every peer's buffer -- 8-byte words {value, solve stamp}, so
 * data and arrival travel in one store -- polls its own buffer for the peers' words and merges.  One process per GPU; every rank must call mppi_solve / mppi_enqueue the same number
 * of times, and the ranks should pass a barrier between mppi_comm_connect and their first solve.  A peer whose
 * record does not arrive within MPPI_OPT_EXCHANGE_TIMEOUT_MS (default 2 s, device-side) fails the solve: the warm
 * start keeps its previous value, and mppi_solve / mppi_download / mppi_synchronize return MPPI_ERR_NCCL. */
#define MPPI_IPC_HANDLE_BYTES 64
int mppi_comm_export(mppi_handle h, int n_ranks, void *handle_out /* MPPI_IPC_HANDLE_BYTES */);
int mppi_comm_connect(mppi_handle h, const void *handles /* n_ranks x MPPI_IPC_HANDLE_BYTES */, int rank, int n_ranks);
/* This rank's partial record of the last solve for one robot: {c_min, sum w, sum w^2, 0, N[(T-1)*U]} with
 * w = exp(-(c - c_min)/lambda) over the local samples and N the weighted sum of the clamped samples -- what the
 * collective exchanges.  Lets a host that owns its own transport (MPI, shared memory) do the exchange itself. */
int mppi_get_record(mppi_handle h, int robot, float *record /* 4 + (T-1)*U */);
/* Host-side merge of per-rank partials, the same arithmetic as the device merge kernel; exported so the
 * sharded path can be tested without GPUs.  partials: n_ranks x (2 + n) floats = {c_min, sum_w, num[n]};
 * u_out[n] = merged weighted mean. */
int mppi_merge_partials(const float *partials, int n_ranks, int n, double lambda, float *u_out, double *stats);

/* ---- host-only helpers shared with the C++ classes (no GPU needed) -------------------------------------- */

/* get_CurrentIndex (DD:126-140) and calc_RefPath (DD:156-181) in double, as used inside mppi_solve. */
int mppi_calc_ref_path(const double *path_xy, int n_points, double px, double py, double v_ref, double dt,
                       double resolution, int horizon, double *window_xyyaw, int *current_index);
/* Philox4x32-10 block function (counter[4], key[2]) -> out[4]; the generator behind the noise kernel. */
void mppi_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* MPPI_B200_H */
