/*
 * mppi_oracle.h -- CPU oracle for the MPPI solve hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, IEEE-double restatement of the per-cycle solve of the three reference nodes
 * (all citations relative to /root/reference):
 *   DD = src/diff_drive_mppi.cpp, SD = src/steering_diff_drive_mppi.cpp, FB = src/full_body_mppi.cpp
 *
 * Nothing under ccv_mppi_path_tracker_b200/ (the product) may include, link or call this file.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Parity pin: the reference ships no tests or golden vectors.  This restatement is pinned
 * against the UNMODIFIED reference translation units compiled against stub ROS/tf/Eigen headers
 * (oracle/ref_shim -> oracle/_ref/libref_*.so, see oracle/Makefile) by tests/test_oracle_vs_ref.py when
 * /root/reference is present, and against the golden vectors generated from that build (tests/golden/).
 *
 * Frozen decisions for the reference's undefined behaviour:
 *   D1  calc_Cost (DD:199-207, SD:215-223): velocity term only for t < T-1 (v_[T-1] is out of bounds);
 *       path term for every t < T.
 *   D2  determine_OptimalSolution (DD:228-236, SD:244-254, FB:311-325): t < T-1 only.
 *   D3  weights: literal exp(-c/lambda) (DD:219) or shifted exp(-(c-c_min)/lambda); identical after normalisation.
 *   D4  yaw_ref[T-1] = 0; duplicate window points allowed; min_distance starts at 100.0 (acts as a cap);
 *       first minimum wins; warm start is not time shifted; dt is a per-solve input.
 *   D5  sample = clamp(u*_t + sigma * eps[t][i][u]) with eps a supplied standard-normal tensor in the
 *       reference's draw order [t][i][u] (DD:86-100); FB steer_off zeroes direction after clamping (FB:517).
 */
#ifndef MPPI_ORACLE_H
#define MPPI_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_DIFF_DRIVE = 0, ORACLE_STEERING = 1, ORACLE_FULL_BODY = 2 };

/* Parameter surface of the three nodes (DD:17-34, SD:18-36, FB:8-46).  Same field order as mppi_params
 * in include/mppi_b200.h so a test can fill both from one table. */
typedef struct {
  double control_noise;   /* sigma, one value for every control (DD:20, SD:21, FB:11) */
  double lambda;          /* DD:21 */
  double v_ref;           /* DD:28 */
  double resolution;      /* DD:29 */
  double u_min[5];        /* v, w, steer|direction, roll_v, pitch_v (DD:24-26, SD:26-28, FB:20-26) */
  double u_max[5];
  double path_weight;     /* DD:33 */
  double v_weight;        /* DD:34 ("control_weight"), FB:35 */
  double zmp_weight;      /* FB:36 (0 when roll_off, FB:43-46) */
  double roll_v_weight;   /* FB:37 (0 when roll_off) */
  double back_weight;     /* FB:38 */
  double yaw_weight;      /* FB:39 */
  int steer_off;          /* FB:41 */
  int reserved;
} oracle_params;

int oracle_num_controls(int model);        /* 2, 3, 5 */
int oracle_num_states(int model);          /* 3, 3, 5 */

/* a8: get_CurrentIndex (DD:126-140, SD:142-156, FB:335-349). path_xy = N x 2 doubles. */
int oracle_current_index(double px, double py, const double *path_xy, int n_path);

/* a9: calc_RefPath (DD:156-181). Writes x_ref,y_ref,yaw_ref [T]; returns current_index_. */
int oracle_calc_ref_path(double px, double py, const double *path_xy, int n_path, double v_ref, double dt,
                         double resolution, int T, double *x_ref, double *y_ref, double *yaw_ref);

/* a10: calc_MinDistance (DD:183-192) -- additionally returns the first-min index (-1 when nothing is < 100). */
double oracle_min_distance(double x, double y, const double *x_ref, const double *y_ref, int T, int *argmin);

/* a7: computeZMPfromModel (FB:597-603) with the constants of FBh:30, FBh:213-216, FB:86-91. out = zmp xyz. */
void oracle_zmp_from_model(const double CoM[3], const double accel[3], const double HGdot[3], double out[3]);

typedef struct {
  /* all optional (NULL = not wanted) */
  double *controls;   /* [K][T-1][U]  clamped samples (a3) */
  double *states;     /* [K][T][S]    predicted states (a6); S = 3 or 5 */
  double *zmp;        /* [K][T-2][2]  FB only (a6 second loop) */
  double *cost;       /* [K]          (a11) */
  int *nearest;       /* [K][T]       argmin of a10 per state (FB: t<T-2 filled, rest -1) */
  double *weights;    /* [K]          normalised (a12) */
  double *window;     /* [T][3]       x_ref,y_ref,yaw_ref (a9) */
  double *stats;      /* [3]          c_min, sum of shifted weights, effective sample size */
  int *current_index; /* [1] */
} oracle_outputs;

/* One full solve: a3(D5) -> a6 -> a9 -> a11 -> a12 -> a13.
 *   state      : x,y,yaw (DD,SD) or x,y,yaw,roll,pitch (FB)
 *   eps        : [T-1][K][U] float32 standard normals (reference draw order)
 *   u_nominal  : [T-1][U] in: previous optimal_solution, out: new one
 *   shifted    : 0 literal weights (DD:219), 1 min-shifted (D3)
 *   nthreads   : OpenMP threads over samples (1 = the reference's single thread)
 * Returns 0, or <0 on bad arguments. */
int oracle_solve(int model, const oracle_params *p, int K, int T, const double *state, double dt,
                 const double *path_xy, int n_path, const float *eps, double *u_nominal, int shifted,
                 int nthreads, const oracle_outputs *out);

/* Structure-faithful timing mode for the CPU baseline: std::mt19937-free, but keeps the reference's
 * per-call heap copies of the sample and window (DD:183, DD:194 pass by value) when literal_copies != 0.
 * Runs n_solves chained solves on internally generated noise (xorshift + Box-Muller, double) and returns
 * seconds of wall time; rollout-steps = n_solves * K * (T-1). */
double oracle_time_solves(int model, const oracle_params *p, int K, int T, const double *state, double dt,
                          const double *path_xy, int n_path, int n_solves, int literal_copies, int nthreads);

/* Synthetic path of src/reference_path_creator.cpp:38-46 (accumulated s += resolution loop).
 * Returns the number of points written (<= cap); xy = cap x 2. */
int oracle_make_sin_path(double course_length, double resolution, double A1, double omega1, double delta1,
                         double A2, double omega2, double delta2, double A3, double omega3, double delta3,
                         double init_x, double init_y, double *xy, int cap);

#ifdef __cplusplus
}
#endif
#endif
