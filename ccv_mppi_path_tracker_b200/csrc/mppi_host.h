// mppi_host.h -- host-side (double precision) pieces of the solve that run before the kernels:
// parameter conversion, get_CurrentIndex and calc_RefPath.  Header-only so that the library, the C++ controller
// classes and the test-only FP32 twin all convert inputs identically.
// Reference: /root/reference/src/diff_drive_mppi.cpp (DD); the SD and FB copies are identical
// (SD:142-197, FB:335-392).
#pragma once
#include <math.h>
#include <stddef.h>

#include "../../include/mppi_b200.h"
#include "mppi_math.h"

namespace mppi {

// ROS-parameter surface (doubles) -> FP32 per-solve constants.  Full-body constants: FBh:213-216 (body box,
// mass), FB:86-91 (base2CoM = height/2, I_O diagonal), FBh:30 (gravity_ z = -9.8).
inline SolveParams make_solve_params(int model, int horizon, const mppi_params &p, double dt) {
  SolveParams P;
  memset(&P, 0, sizeof P);
  P.model = model;
  P.T = horizon;
  P.U = num_controls(model);
  P.steer_off = p.steer_off;
  P.dt = (float)dt;
  P.inv_dt = 1.0f / P.dt;
  P.sigma = (float)p.control_noise;
  P.v_ref = (float)p.v_ref;
  for (int k = 0; k < kMaxControls; ++k) {
    P.u_min[k] = (float)p.u_min[k];
    P.u_max[k] = (float)p.u_max[k];
  }
  P.path_weight = (float)p.path_weight;
  P.v_weight = (float)p.v_weight;
  P.zmp_weight = (float)p.zmp_weight;
  P.roll_v_weight = (float)p.roll_v_weight;
  P.back_weight = (float)p.back_weight;
  P.yaw_weight = (float)p.yaw_weight;
  const double mass = 60.0, h = 0.8075, d = 0.208, w = 0.208, gz = -9.8;
  const double b = h / 2;
  P.base2com = (float)b;
  P.ixx = (float)((mass * (w * w + h * h)) / 12 + mass * b * b);
  P.iyy = (float)((mass * (h * h + d * d)) / 12 + mass * b * b);
  P.inv_gz = (float)(1.0 / gz);
  P.inv_mgz = (float)(1.0 / (mass * gz));
  return P;
}

// get_CurrentIndex: DD:126-140.  (dx*dx is the exact value of pow(dx, 2).)
inline int current_index(const double *path_xy, int n, double px, double py) {
  int index = 0;
  double min_distance = 100.0;
  for (int i = 0; i < n; ++i) {
    double dx = px - path_xy[2 * i], dy = py - path_xy[2 * i + 1];
    double distance = sqrt(dx * dx + dy * dy);
    if (distance < min_distance) {
      min_distance = distance;
      index = i;
    }
  }
  return index;
}

// calc_RefPath after get_CurrentIndex: DD:160-181.  window = T x {x_ref, y_ref, yaw_ref}; yaw_ref[T-1] = 0 (never
// written, DD:44).
// with_yaw = false leaves the yaw_ref column as it is (window_yaw fills it in later): the kernels take yaw_ref[0]
// from the FP32 window, so the T-1 atan2 calls are only needed when somebody asks for the window itself.
inline void window_yaw(int T, double *window) {  // DD:175-178
  for (int i = 0; i + 1 < T; ++i)
    window[3 * i + 2] = atan2(window[3 * (i + 1) + 1] - window[3 * i + 1], window[3 * (i + 1)] - window[3 * i]);
  if (T > 0) window[3 * (T - 1) + 2] = 0.0;
}
inline void window_from_index(const double *path_xy, int n, int cur, double v_ref, double dt, double resolution, int T,
                              double *window, bool with_yaw = true) {
  const double step = v_ref * dt / resolution;  // DD:160
  for (int i = 0; i < T; ++i) {
    int index = (int)(cur + i * step);  // DD:163 truncation of a double
    if (n <= 0) {
      window[3 * i] = window[3 * i + 1] = 0.0;
    } else {
      if (index < 0 || index >= n) index = n - 1;  // DD:169-173 last pose
      window[3 * i] = path_xy[2 * index];
      window[3 * i + 1] = path_xy[2 * index + 1];
    }
  }
  if (with_yaw) window_yaw(T, window);
}

// calc_RefPath: DD:156-181
inline int calc_ref_path(const double *path_xy, int n, double px, double py, double v_ref, double dt,
                         double resolution, int T, double *window, bool with_yaw = true) {
  const int cur = current_index(path_xy, n, px, py);
  window_from_index(path_xy, n, cur, v_ref, dt, resolution, T, window, with_yaw);
  return cur;
}

// Robot-centred FP32 window: subtracting the pose in double keeps FP32 resolution independent of where the
// robot is on the map.  out = T x {x, y}.
inline void window_to_robot_frame(const double *window, int T, double px, double py, float *out) {
  for (int j = 0; j < T; ++j) {
    out[2 * j] = (float)(window[3 * j] - px);
    out[2 * j + 1] = (float)(window[3 * j + 1] - py);
  }
}

// state record {yaw, roll, pitch, 0}: the kernels work in the robot-centred frame (x = y = 0); yaw_ref_[0] is derived
// from the FP32 window by the kernels
inline void state_to_robot_frame(int model, const double *state, float *out) {
  out[0] = (float)state[2];
  out[1] = model == kFullBody ? (float)state[3] : 0.f;
  out[2] = model == kFullBody ? (float)state[4] : 0.f;
  out[3] = 0.f;
}

}  // namespace mppi
