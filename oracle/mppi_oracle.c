/*
 * mppi_oracle.c -- CPU oracle (IEEE double) for the MPPI solve hot path.  TEST INFRASTRUCTURE ONLY.
 * See mppi_oracle.h for scope, the parity pin and the frozen decisions D1-D5.
 * Citations: DD = /root/reference/src/diff_drive_mppi.cpp, SD = .../steering_diff_drive_mppi.cpp,
 *            FB = .../full_body_mppi.cpp, FBh = .../include/ccv_mppi_path_tracker/full_body_mppi.h
 */
#include "mppi_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_num_controls(int model) { return model == ORACLE_DIFF_DRIVE ? 2 : model == ORACLE_STEERING ? 3 : 5; }
int oracle_num_states(int model) { return model == ORACLE_FULL_BODY ? 5 : 3; }

/* clamp: DD:62-67, SD:78-83, FB:522-526 (NaN passes through both comparisons) */
static void clamp_ref(double *val, double lo, double hi) {
  if (*val < lo) *val = lo;
  else if (*val > hi) *val = hi;
}

/* get_CurrentIndex: DD:126-140 */
int oracle_current_index(double px, double py, const double *path_xy, int n_path) {
  int index = 0;
  double min_distance = 100.0;
  for (int i = 0; i < n_path; i++) {
    double distance = sqrt(pow(px - path_xy[2 * i], 2) + pow(py - path_xy[2 * i + 1], 2));
    if (distance < min_distance) {
      min_distance = distance;
      index = i;
    }
  }
  return index;
}

/* calc_RefPath: DD:156-181 (SD:172-197, FB:365-392 identical) */
int oracle_calc_ref_path(double px, double py, const double *path_xy, int n_path, double v_ref, double dt,
                         double resolution, int T, double *x_ref, double *y_ref, double *yaw_ref) {
  int current_index = oracle_current_index(px, py, path_xy, n_path);
  double step = v_ref * dt / resolution; /* DD:160 */
  for (int i = 0; i < T; i++) {
    int index = (int)(current_index + i * step); /* DD:163: int index = current_index_ + i * step; */
    if (n_path > 0 && index >= 0 && index < n_path) {
      x_ref[i] = path_xy[2 * index];
      y_ref[i] = path_xy[2 * index + 1];
    } else if (n_path > 0) { /* DD:169-173: the last pose */
      x_ref[i] = path_xy[2 * (n_path - 1)];
      y_ref[i] = path_xy[2 * (n_path - 1) + 1];
    } else {
      x_ref[i] = 0.0;
      y_ref[i] = 0.0;
    }
  }
  for (int i = 0; i < T - 1; i++) yaw_ref[i] = atan2(y_ref[i + 1] - y_ref[i], x_ref[i + 1] - x_ref[i]); /* DD:175-178 */
  if (T > 0) yaw_ref[T - 1] = 0.0; /* D4: never written by the reference, stays 0 from resize (DD:44) */
  return current_index;
}

/* calc_MinDistance: DD:183-192 */
double oracle_min_distance(double x, double y, const double *x_ref, const double *y_ref, int T, int *argmin) {
  double min_distance = 100.0;
  int arg = -1;
  for (int i = 0; i < T; i++) {
    double distance = sqrt(pow(x - x_ref[i], 2) + pow(y - y_ref[i], 2));
    if (distance < min_distance) {
      min_distance = distance;
      arg = i;
    }
  }
  if (argmin) *argmin = arg;
  return min_distance;
}

/* --- minimal 3-vector algebra standing in for Eigen::Vector3d (FB:475-483, FB:599-601) --- */
typedef struct { double x, y, z; } v3;
static v3 v3_make(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 v3_scale(double s, v3 a) { return v3_make(s * a.x, s * a.y, s * a.z); }
static v3 v3_div(v3 a, double s) { return v3_make(a.x / s, a.y / s, a.z / s); }
static double v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static v3 v3_cross(v3 a, v3 b) { return v3_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

/* constants: FBh:30 gravity_, FBh:213-216 body box + mass, FB:86 base2CoM = upper_body_height/2 */
static const double kMass = 60.0;
static const double kBodyH = 0.8075, kBodyD = 0.208, kBodyW = 0.208;
static const v3 kGravity = {0.0, 0.0, -9.8};

static double base2com(void) { return kBodyH / 2; }
static v3 inertia_diag(void) { /* FB:87-91 */
  double b = base2com();
  return v3_make((kMass * (kBodyW * kBodyW + kBodyH * kBodyH)) / 12 + kMass * b * b,
                 (kMass * (kBodyH * kBodyH + kBodyD * kBodyD)) / 12 + kMass * b * b,
                 (kMass * (kBodyD * kBodyD + kBodyW * kBodyW)) / 12);
}

/* computeZMPfromModel: FB:597-603 */
static v3 zmp_from_model(v3 CoM, v3 accel, v3 HGdot) {
  v3 z = v3_make(0.0, 0.0, 1.0);
  v3 M_O = v3_sub(v3_sub(v3_cross(CoM, v3_scale(kMass, kGravity)), v3_cross(CoM, v3_scale(kMass, accel))), HGdot);
  return v3_div(v3_cross(z, M_O), kMass * v3_dot(v3_sub(kGravity, accel), z));
}
void oracle_zmp_from_model(const double CoM[3], const double accel[3], const double HGdot[3], double out[3]) {
  v3 r = zmp_from_model(v3_make(CoM[0], CoM[1], CoM[2]), v3_make(accel[0], accel[1], accel[2]),
                        v3_make(HGdot[0], HGdot[1], HGdot[2]));
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

/* Per-sample scratch in the reference's RobotStates layout (DDh:20-50, SDh:21-54, FBh:34-65). */
typedef struct {
  double *x, *y, *yaw, *roll, *pitch;        /* [T] */
  double *u[5];                               /* [T-1] v, w, steer|direction, roll_v, pitch_v */
  double *zmp_x, *zmp_y;                      /* [T-2] */
} robot_states;

static void rs_alloc(robot_states *s, int T) {
  int n = T > 1 ? T : 1;
  s->x = (double *)calloc((size_t)n * 12, sizeof(double));
  s->y = s->x + n; s->yaw = s->y + n; s->roll = s->yaw + n; s->pitch = s->roll + n;
  for (int k = 0; k < 5; k++) s->u[k] = s->pitch + n + (size_t)k * n;
  s->zmp_x = s->u[4] + n; s->zmp_y = s->zmp_x + n;
}
static void rs_free(robot_states *s) { free(s->x); }

/* predict_NextState: DD:104-109, SD:120-125, FB:445-452 */
static void predict_next_state(int model, robot_states *s, int t, double dt) {
  double heading = s->yaw[t];
  if (model != ORACLE_DIFF_DRIVE) heading = s->yaw[t] + s->u[2][t]; /* SD:122 steer_, FB:447 direction_ */
  s->x[t + 1] = s->x[t] + s->u[0][t] * cos(heading) * dt;
  s->y[t + 1] = s->y[t] + s->u[0][t] * sin(heading) * dt;
  s->yaw[t + 1] = s->yaw[t] + s->u[1][t] * dt;
  if (model == ORACLE_FULL_BODY) {
    s->roll[t + 1] = s->roll[t] + s->u[3][t] * dt;   /* FB:450 */
    s->pitch[t + 1] = s->pitch[t] + s->u[4][t] * dt; /* FB:451 */
  }
}

/* predict_States body for one sample: DD:113-122, FB:456-487 */
static void predict_states(int model, robot_states *s, const double *state, int T, double dt) {
  s->x[0] = state[0];
  s->y[0] = state[1];
  s->yaw[0] = state[2];
  if (model == ORACLE_FULL_BODY) {
    s->roll[0] = state[3];  /* FB:463 */
    s->pitch[0] = state[4]; /* FB:464 */
  }
  for (int t = 0; t < T - 1; t++) predict_next_state(model, s, t, dt);
  if (model == ORACLE_FULL_BODY) {
    const double b = base2com();
    const v3 I = inertia_diag();
    for (int t = 0; t < T - 2; t++) { /* FB:468-486 */
      double drive_accel = (s->u[0][t + 1] - s->u[0][t]) / dt;
      double ac = s->u[0][t] * s->u[1][t];
      double drive_accel_x = drive_accel * cos(s->u[2][t]) - ac * sin(s->u[2][t]);
      double drive_accel_y = drive_accel * sin(s->u[2][t]) + ac * cos(s->u[2][t]);
      v3 accel = v3_make(drive_accel_x, drive_accel_y, 0.0);
      v3 next_omega = v3_make(s->u[3][t + 1], s->u[4][t + 1], s->u[1][t + 1]);
      v3 omega = v3_make(s->u[3][t], s->u[4][t], s->u[1][t]);
      v3 HG_next = v3_make(I.x * next_omega.x, I.y * next_omega.y, I.z * next_omega.z); /* I_O diagonal */
      v3 HG = v3_make(I.x * omega.x, I.y * omega.y, I.z * omega.z);
      v3 HG_dot = v3_div(v3_sub(HG_next, HG), dt);
      v3 CoM = v3_make(b * sin(s->pitch[t]), -b * sin(s->roll[t]), b * cos(s->pitch[t]) * cos(s->roll[t]));
      v3 ZMP = zmp_from_model(CoM, accel, HG_dot);
      s->zmp_x[t] = ZMP.x;
      s->zmp_y[t] = ZMP.y;
    }
  }
}

/* calc_Cost: DD:194-210 / SD:210-226 (with D1) and FB:404-424. nearest may be NULL. */
static double calc_cost(int model, const oracle_params *p, const robot_states *s, int T, const double *x_ref,
                        const double *y_ref, const double *yaw_ref, int *nearest, int literal_copies) {
  double cost = 0.0;
  double *xr = (double *)x_ref, *yr = (double *)y_ref;
  if (model != ORACLE_FULL_BODY) {
    for (int t = 0; t < T; t++) {
      int arg;
      if (literal_copies) { /* calc_MinDistance takes both vectors by value: DD:183 */
        xr = (double *)malloc(sizeof(double) * (size_t)T); memcpy(xr, x_ref, sizeof(double) * (size_t)T);
        yr = (double *)malloc(sizeof(double) * (size_t)T); memcpy(yr, y_ref, sizeof(double) * (size_t)T);
      }
      double distance = oracle_min_distance(s->x[t], s->y[t], xr, yr, T, &arg);
      if (literal_copies) { free(xr); free(yr); }
      if (nearest) nearest[t] = arg;
      double v_cost = 0.0;
      if (t < T - 1) v_cost = (s->u[0][t] - p->v_ref) * (s->u[0][t] - p->v_ref); /* D1 */
      cost += p->path_weight * distance * distance + p->v_weight * v_cost; /* DD:206 */
    }
  } else {
    cost += p->yaw_weight * (s->yaw[0] - yaw_ref[0]) * (s->yaw[0] - yaw_ref[0]); /* FB:408 */
    if (nearest) for (int t = 0; t < T; t++) nearest[t] = -1;
    for (int t = 0; t < T - 2; t++) {
      int arg;
      /* FB:411 evaluates calc_MinDistance twice; the value is the same */
      double d1 = oracle_min_distance(s->x[t], s->y[t], xr, yr, T, &arg);
      double d2 = literal_copies ? oracle_min_distance(s->x[t], s->y[t], xr, yr, T, NULL) : d1;
      if (nearest) nearest[t] = arg;
      cost += p->path_weight * d1 * d2;
      cost += p->v_weight * (s->u[0][t] - p->v_ref) * (s->u[0][t] - p->v_ref);               /* FB:413 */
      cost += p->zmp_weight * s->zmp_y[t] * s->zmp_y[t];                                       /* FB:416 */
      cost += p->roll_v_weight * (s->u[3][t + 1] - s->u[3][t]) * (s->u[3][t + 1] - s->u[3][t]); /* FB:418 */
      if (s->u[0][t] < 0.0) cost += p->back_weight * s->u[0][t] * s->u[0][t];                  /* FB:420 */
    }
  }
  return cost;
}

/* sampling with supplied noise (D5): DD:86-100, SD:102-117, FB:496-519 */
static void sample_controls(int model, const oracle_params *p, robot_states *s, int i, int K, int T,
                            const float *eps, const double *u_nominal) {
  const int U = oracle_num_controls(model);
  for (int t = 0; t < T - 1; t++) {
    for (int k = 0; k < U; k++) {
      /* std::normal_distribution(mean, sigma)(mt) = z * sigma + mean */
      double val = (double)eps[((size_t)t * K + i) * U + k] * p->control_noise + u_nominal[t * U + k];
      clamp_ref(&val, p->u_min[k], p->u_max[k]);
      s->u[k][t] = val;
    }
    if (model == ORACLE_FULL_BODY && p->steer_off) s->u[2][t] = 0.0; /* FB:517 */
  }
}

int oracle_solve(int model, const oracle_params *p, int K, int T, const double *state, double dt,
                 const double *path_xy, int n_path, const float *eps, double *u_nominal, int shifted,
                 int nthreads, const oracle_outputs *out) {
  if (model < 0 || model > 2 || K <= 0 || T < 2 || !p || !state || !eps || !u_nominal) return -1;
  const int U = oracle_num_controls(model), S = oracle_num_states(model);
  oracle_outputs none;
  memset(&none, 0, sizeof none);
  if (!out) out = &none;

  double *x_ref = (double *)malloc(sizeof(double) * 3 * (size_t)T);
  double *y_ref = x_ref + T, *yaw_ref = y_ref + T;
  int cur = oracle_calc_ref_path(state[0], state[1], path_xy, n_path, p->v_ref, dt, p->resolution, T, x_ref, y_ref, yaw_ref);
  if (out->current_index) *out->current_index = cur;
  if (out->window)
    for (int t = 0; t < T; t++) { out->window[3 * t] = x_ref[t]; out->window[3 * t + 1] = y_ref[t]; out->window[3 * t + 2] = yaw_ref[t]; }

  double *cost = (double *)malloc(sizeof(double) * (size_t)K);
  double *weights = (double *)malloc(sizeof(double) * (size_t)K);
  double *ctrl = (double *)malloc(sizeof(double) * (size_t)K * (T - 1) * U);
  if (nthreads < 1) nthreads = 1;

#pragma omp parallel num_threads(nthreads)
  {
    robot_states s;
    rs_alloc(&s, T);
#pragma omp for schedule(static)
    for (int i = 0; i < K; i++) {
      sample_controls(model, p, &s, i, K, T, eps, u_nominal);
      predict_states(model, &s, state, T, dt);
      cost[i] = calc_cost(model, p, &s, T, x_ref, y_ref, yaw_ref, out->nearest ? out->nearest + (size_t)i * T : NULL, 0);
      for (int t = 0; t < T - 1; t++)
        for (int k = 0; k < U; k++) ctrl[((size_t)i * (T - 1) + t) * U + k] = s.u[k][t];
      if (out->states)
        for (int t = 0; t < T; t++) {
          double *d = out->states + ((size_t)i * T + t) * S;
          d[0] = s.x[t]; d[1] = s.y[t]; d[2] = s.yaw[t];
          if (S == 5) { d[3] = s.roll[t]; d[4] = s.pitch[t]; }
        }
      if (out->zmp && model == ORACLE_FULL_BODY)
        for (int t = 0; t < T - 2; t++) {
          out->zmp[((size_t)i * (T - 2) + t) * 2] = s.zmp_x[t];
          out->zmp[((size_t)i * (T - 2) + t) * 2 + 1] = s.zmp_y[t];
        }
    }
    rs_free(&s);
  }

  /* calc_Weights: DD:215-222 -- serial, i ascending */
  double c_min = cost[0];
  for (int i = 1; i < K; i++) if (cost[i] < c_min) c_min = cost[i];
  double sum = 0.0;
  for (int i = 0; i < K; i++) {
    weights[i] = shifted ? exp(-(cost[i] - c_min) / p->lambda) : exp(-cost[i] / p->lambda);
    sum += weights[i];
  }
  for (int i = 0; i < K; i++) weights[i] /= sum;

  /* determine_OptimalSolution: DD:228-236 with D2 (t < T-1) */
  for (int t = 0; t < T - 1; t++)
    for (int k = 0; k < U; k++) {
      double acc = 0.0;
      for (int i = 0; i < K; i++) acc += weights[i] * ctrl[((size_t)i * (T - 1) + t) * U + k];
      u_nominal[t * U + k] = acc;
    }

  if (out->cost) memcpy(out->cost, cost, sizeof(double) * (size_t)K);
  if (out->weights) memcpy(out->weights, weights, sizeof(double) * (size_t)K);
  if (out->controls) memcpy(out->controls, ctrl, sizeof(double) * (size_t)K * (T - 1) * U);
  if (out->stats) {
    double s_shift = 0.0, ess = 0.0;
    for (int i = 0; i < K; i++) { s_shift += exp(-(cost[i] - c_min) / p->lambda); ess += weights[i] * weights[i]; }
    out->stats[0] = c_min; out->stats[1] = s_shift; out->stats[2] = 1.0 / ess;
  }
  free(ctrl); free(weights); free(cost); free(x_ref);
  return 0;
}

/* ---- CPU baseline timing ---- */
static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static uint64_t xs_next(uint64_t *s) { uint64_t x = *s; x ^= x << 13; x ^= x >> 7; x ^= x << 17; return *s = x; }
static double xs_normal(uint64_t *s) {
  double u1 = ((double)(xs_next(s) >> 11) + 1.0) * (1.0 / 9007199254740993.0);
  double u2 = (double)(xs_next(s) >> 11) * (1.0 / 9007199254740992.0);
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

double oracle_time_solves(int model, const oracle_params *p, int K, int T, const double *state, double dt,
                          const double *path_xy, int n_path, int n_solves, int literal_copies, int nthreads) {
  const int U = oracle_num_controls(model);
  double *u_nominal = (double *)calloc((size_t)(T - 1) * U, sizeof(double));
  double *x_ref = (double *)malloc(sizeof(double) * 3 * (size_t)T);
  double *y_ref = x_ref + T, *yaw_ref = y_ref + T;
  double *cost = (double *)malloc(sizeof(double) * (size_t)K);
  double *ctrl = (double *)malloc(sizeof(double) * (size_t)K * (T - 1) * U);
  if (nthreads < 1) nthreads = 1;
  double t0 = now_s();
  for (int it = 0; it < n_solves; it++) {
    oracle_calc_ref_path(state[0], state[1], path_xy, n_path, p->v_ref, dt, p->resolution, T, x_ref, y_ref, yaw_ref);
#pragma omp parallel num_threads(nthreads)
    {
      robot_states s;
      rs_alloc(&s, T);
#ifdef _OPENMP
      uint64_t rng = 0x9E3779B97F4A7C15ull * (uint64_t)(1 + it) + 0xD1B54A32D192ED03ull * (uint64_t)(1 + omp_get_thread_num());
#else
      uint64_t rng = 0x9E3779B97F4A7C15ull * (uint64_t)(1 + it);
#endif
#pragma omp for schedule(static)
      for (int i = 0; i < K; i++) {
        for (int t = 0; t < T - 1; t++) {
          for (int k = 0; k < U; k++) {
            double val = xs_normal(&rng) * p->control_noise + u_nominal[t * U + k];
            clamp_ref(&val, p->u_min[k], p->u_max[k]);
            s.u[k][t] = val;
          }
          if (model == ORACLE_FULL_BODY && p->steer_off) s.u[2][t] = 0.0;
        }
        predict_states(model, &s, state, T, dt);
        if (literal_copies) { /* calc_Cost(RobotStates sample) takes the sample by value: DD:194 */
          robot_states c;
          rs_alloc(&c, T);
          memcpy(c.x, s.x, sizeof(double) * 12 * (size_t)(T > 1 ? T : 1));
          cost[i] = calc_cost(model, p, &c, T, x_ref, y_ref, yaw_ref, NULL, 1);
          rs_free(&c);
        } else {
          cost[i] = calc_cost(model, p, &s, T, x_ref, y_ref, yaw_ref, NULL, 0);
        }
        for (int t = 0; t < T - 1; t++)
          for (int k = 0; k < U; k++) ctrl[((size_t)i * (T - 1) + t) * U + k] = s.u[k][t];
      }
      rs_free(&s);
    }
    double c_min = cost[0];
    for (int i = 1; i < K; i++) if (cost[i] < c_min) c_min = cost[i];
    double sum = 0.0;
    for (int i = 0; i < K; i++) { cost[i] = exp(-(cost[i] - c_min) / p->lambda); sum += cost[i]; }
    for (int i = 0; i < K; i++) cost[i] /= sum;
    for (int t = 0; t < T - 1; t++)
      for (int k = 0; k < U; k++) {
        double acc = 0.0;
        for (int i = 0; i < K; i++) acc += cost[i] * ctrl[((size_t)i * (T - 1) + t) * U + k];
        u_nominal[t * U + k] = acc;
      }
  }
  double t1 = now_s();
  free(ctrl); free(cost); free(x_ref); free(u_nominal);
  return t1 - t0;
}

/* reference_path_creator.cpp:38-46 */
int oracle_make_sin_path(double course_length, double resolution, double A1, double omega1, double delta1,
                         double A2, double omega2, double delta2, double A3, double omega3, double delta3,
                         double init_x, double init_y, double *xy, int cap) {
  int n = 0;
  for (double s = 0.0; s < course_length; s += resolution) {
    if (n >= cap) break;
    double x = init_x + s;
    double y = A1 * cos(2 * M_PI * omega1 * s + delta1) + A2 * cos(2 * M_PI * omega2 * s + delta2) +
               A3 * cos(2 * M_PI * omega3 * s + delta3) + init_y;
    y -= A1 + A2 + A3;
    xy[2 * n] = x;
    xy[2 * n + 1] = y;
    n++;
  }
  return n;
}
