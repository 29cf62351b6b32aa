// mppi_twin.cpp -- FP32 "twin" of the rollout + cost kernel on the CPU.  TEST INFRASTRUCTURE ONLY.
//
// Compiles the SAME header as the device code (ccv_mppi_path_tracker_b200/csrc/mppi_math.h) for the host with
// -ffp-contract=off, so every FP32 operation (explicit fmaf, own sincos) rounds exactly as on the GPU
// (-fmad=false).  Contract checked by tests/test_gpu_parity.py:
//   GPU nearest indices == twin nearest indices, bit for bit; GPU costs == twin costs, bit for bit.
// The twin itself is checked against the FP64 oracle (mppi_oracle.c): indices equal except at near-ties,
// costs within 1e-5 relative (tests/test_oracle.py).  The twin is NOT a second opinion on the algorithm (it
// shares the product header) -- the FP64 oracle, which restates the reference independently, is.
#include <stdint.h>

#include <vector>

#include "../ccv_mppi_path_tracker_b200/csrc/mppi_host.h"

using namespace mppi;

namespace {
struct EpsHost {
  const float *eps;  // [T-1][K][U]
  int K, U, i;
  float get(int t, int u) const { return eps[((size_t)t * K + i) * U + u]; }
};
struct NomHost {
  const float *nom;
  int U;
  float get(int t, int u) const { return nom[t * U + u]; }
};
struct WinHost {
  const float *w;
  float x(int j) const { return w[2 * j]; }
  float y(int j) const { return w[2 * j + 1]; }
};
struct TapSink {
  int T, U;
  int *near_out;
  float *d2, *states, *zmp_out, *controls;
  void state(int t, float x, float y, float yaw, float roll, float pitch) {
    if (states) {
      float *s = states + (size_t)t * 5;
      s[0] = x; s[1] = y; s[2] = yaw; s[3] = roll; s[4] = pitch;
    }
  }
  void nearest(int t, int j, float v) {
    if (near_out) near_out[t] = j;
    if (d2) d2[t] = v;
  }
  void control(int t, int u, float v) {
    if (controls) controls[(size_t)t * U + u] = v;
  }
  void zmp(int t, float zx, float zy) {
    if (zmp_out) {
      zmp_out[2 * t] = zx;
      zmp_out[2 * t + 1] = zy;
    }
  }
};

template <int MODEL>
void run(const SolveParams &P, int K, const float *state0, float yaw_ref0, const float *win, const float *eps,
         const float *nom, float *cost, int *nearest, float *d2, float *states, float *zmp, float *controls) {
  const int T = P.T, U = P.U;
  for (int i = 0; i < K; ++i) {
    EpsHost e{eps, K, U, i};
    NomHost n{nom, U};
    WinHost w{win};
    TapSink sink{T, U, nearest ? nearest + (size_t)i * T : nullptr, d2 ? d2 + (size_t)i * T : nullptr,
                 states ? states + (size_t)i * T * 5 : nullptr, zmp ? zmp + (size_t)i * T * 2 : nullptr,
                 controls ? controls + (size_t)i * (T - 1) * U : nullptr};
    if (nearest) for (int t = 0; t < T; ++t) nearest[(size_t)i * T + t] = -1;
    cost[i] = rollout_cost_literal<MODEL>(P, state0, yaw_ref0, e, n, w, sink);
  }
}
}  // namespace

extern "C" {

// window: [T][3] doubles (absolute frame, as mppi_get_window returns); eps: [T-1][K][U]; u_nominal [T-1][U].
// Outputs (any may be NULL except cost): cost[K], nearest[K][T], d2[K][T], states[K][T][5] (robot-centred),
// zmp[K][T][2], controls[K][T-1][U].
int twin_rollout_cost(int model, const mppi_params *p, int K, int T, const double *state, double dt,
                      const double *window, const float *eps, const double *u_nominal, float *cost, int *nearest,
                      float *d2, float *states, float *zmp, float *controls) {
  if (model < 0 || model > 2 || !p || K < 1 || T < 2 || !state || !window || !eps || !u_nominal || !cost) return -1;
  SolveParams P = make_solve_params(model, T, *p, dt);
  std::vector<float> win(2 * (size_t)T), nom((size_t)(T - 1) * P.U);
  float rec[4];  // the device's state record {yaw, roll, pitch, -}; the rollout starts at the origin of the robot frame
  window_to_robot_frame(window, T, state[0], state[1], win.data());
  state_to_robot_frame(model, state, rec);
  const float st[5] = {0.f, 0.f, rec[0], rec[1], rec[2]};
  const float yaw_ref0 = yaw_ref0_f32(win[0], win[1], win[2], win[3]);  // what the kernels derive from the same FP32 window
  for (size_t k = 0; k < nom.size(); ++k) nom[k] = (float)u_nominal[k];
  switch (model) {
    case kDiffDrive: run<kDiffDrive>(P, K, st, yaw_ref0, win.data(), eps, nom.data(), cost, nearest, d2, states, zmp, controls); break;
    case kSteering: run<kSteering>(P, K, st, yaw_ref0, win.data(), eps, nom.data(), cost, nearest, d2, states, zmp, controls); break;
    default: run<kFullBody>(P, K, st, yaw_ref0, win.data(), eps, nom.data(), cost, nearest, d2, states, zmp, controls); break;
  }
  return 0;
}

void twin_atan2(const float *y, const float *x, int n, float *out) {
  for (int i = 0; i < n; ++i) out[i] = atan2_f32(y[i], x[i]);
}

void twin_sincos(const float *a, int n, float *s, float *c) {
  for (int i = 0; i < n; ++i) sincos_f32(a[i], s[i], c[i]);
}

// sin / cos of a per-step angle increment (short polynomial inside +-0.35 rad, sincos_f32 beyond)
void twin_sincos_increment(const float *a, int n, float *s, float *c) {
  for (int i = 0; i < n; ++i) sincos_increment(a[i], s[i], c[i]);
}

// the heading recurrence: (cos, sin) of angle0 advanced by n increments; also the directly accumulated angle
void twin_heading(float angle0, const float *increments, int n, float *c_out, float *s_out, float *angle_out) {
  float s, c, a = angle0;
  sincos_f32(a, s, c);
  for (int i = 0; i < n; ++i) {
    rotate_by(c, s, increments[i]);
    a += increments[i];
    c_out[i] = c;
    s_out[i] = s;
    angle_out[i] = a;
  }
}

// clamp of the contract: min.NaN(max.NaN(v, lo), hi) with the PTX ordering of signed zeros
void twin_clamp(const float *v, int n, float lo, float hi, float *out) {
  for (int i = 0; i < n; ++i) out[i] = clamp_ref(v[i], lo, hi);
}

}  // extern "C"
