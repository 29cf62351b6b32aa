// mppi_math.h -- FP32 arithmetic contract of the B200 MPPI core, shared by device kernels and host code.
//
// Every floating-point operation of the rollout + cost path is written here ONCE, as a __host__ __device__
// inline function with explicit fmaf() where a fused multiply-add is wanted.  The CUDA side is compiled with
// -fmad=false and the host side with -ffp-contract=off, so `a*b+c` is two roundings on both and fmaf() is one:
// the device kernels and a host build of this header produce bit-identical states, squared distances, nearest
// indices and per-sample costs.  (expf in the weight kernel is outside that contract.)
//
// What is restated (reference = /root/reference, DD = src/diff_drive_mppi.cpp, SD = src/steering_diff_drive_mppi.cpp,
// FB = src/full_body_mppi.cpp):
//   clamp               DD:62-67
//   sampling            DD:86-100 / SD:102-117 / FB:496-519   sample = clamp(mean + sigma*eps)
//   predict_NextState   DD:104-109 / SD:120-125 / FB:445-452
//   ZMP model           FB:468-486 + FB:597-603 (only zmp_y enters the cost, FB:416)
//   calc_MinDistance    DD:183-192 in the squared domain (min(d,100)^2 == min(d^2,1e4))
//   calc_Cost           DD:194-210 / SD:210-226 / FB:404-424 with decisions D1
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define MPPI_HD __host__ __device__ __forceinline__
#else
#define MPPI_HD inline
#endif

namespace mppi {

enum Model : int { kDiffDrive = 0, kSteering = 1, kFullBody = 2 };

constexpr int kMaxControls = 5;
constexpr float kDist2Cap = 10000.0f;  // min_distance = 100.0 (DD:185) squared

MPPI_HD int num_controls(int model) { return model == kDiffDrive ? 2 : (model == kSteering ? 3 : 5); }
MPPI_HD int num_states(int model) { return model == kFullBody ? 5 : 3; }
// number of states that enter the path cost: t < T (DD:199) or t < T-2 (FB:409)
MPPI_HD int num_cost_states(int model, int T) { return model == kFullBody ? (T - 2 > 0 ? T - 2 : 0) : T; }

// Per-solve constants in FP32.  Built by make_solve_params() on the host, identical for library and twin.
struct SolveParams {
  int model;
  int T;  // horizon (states); T-1 control steps
  int U;
  int steer_off;
  float dt, inv_dt;
  float sigma;
  float v_ref;
  float u_min[kMaxControls], u_max[kMaxControls];
  float path_weight, v_weight, zmp_weight, roll_v_weight, back_weight, yaw_weight;
  // full-body constants (FBh:213-216, FB:86-91, FBh:30)
  float base2com;       // upper_body_height / 2
  float ixx, iyy;       // I_O diagonal (x: roll axis, y: pitch axis)
  float inv_gz;         // 1 / g_z, g_z = -9.8
  float inv_mgz;        // 1 / (mass * g_z)
};

MPPI_HD float bits_to_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
MPPI_HD uint32_t float_to_bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}

// clamp: DD:62-67, `if (v < lo) v = lo; else if (v > hi) v = hi;` -- NaN falls through both tests.
// Evaluated as min.NaN(max.NaN(v, lo), hi): two FMNMX instead of two compare/select pairs, NaN still passes
// through.  Identical to the reference's result for lo <= hi (mppi_create rejects lo > hi) except for the SIGN of a
// zero result when v and the bound are zeros of opposite sign (PTX orders -0 < +0) -- no effect on any distance,
// cost or control value.  The host version reproduces the PTX semantics bit for bit.
MPPI_HD float max_nan(float a, float b) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
#else
  if (a != a || b != b) return bits_to_float(0x7fffffffu);
  if (a == b) return bits_to_float(float_to_bits(a) & float_to_bits(b));  // +0 wins over -0
  return a > b ? a : b;
#endif
}
MPPI_HD float min_nan(float a, float b) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
#else
  if (a != a || b != b) return bits_to_float(0x7fffffffu);
  if (a == b) return bits_to_float(float_to_bits(a) | float_to_bits(b));  // -0 wins over +0
  return a < b ? a : b;
#endif
}
MPPI_HD float clamp_ref(float v, float lo, float hi) { return min_nan(max_nan(v, lo), hi); }

// sampling (D5): std::normal_distribution(mean, sigma) returns z*sigma + mean; then clamp (DD:96-99)
MPPI_HD float sample_control(float eps, float sigma, float mean, float lo, float hi) {
  return clamp_ref(fmaf(eps, sigma, mean), lo, hi);
}

// sin and cos on [-pi/4, pi/4]: the classic single-precision minimax polynomials.
MPPI_HD void sincos_poly(float r, float &sn, float &cs) {
  float z = r * r;
  float sp = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
  sp = fmaf(sp, z, -1.6666654611e-1f);
  sn = fmaf(sp, z * r, r);
  float cp = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
  cp = fmaf(cp, z, 4.166664568298827e-2f);
  cs = fmaf(cp, z * z, fmaf(z, -0.5f, 1.0f));
}

// sin and cos of one angle, same bits on host and device.
// Cody-Waite 3-term reduction by pi/2 (round-to-nearest via the 1.5*2^23 magic constant) and sincos_poly on the
// reduced argument; |a| < 2^22*pi/2 keeps the quadrant exact, the reduced argument is accurate to ~1 ulp for
// |a| < ~1e4 (yaw after a 10 s horizon is < 40 rad).
MPPI_HD void sincos_reduce(float a, float &s, float &c) {
  const float kMagic = 12582912.0f;           // 1.5 * 2^23
  const float kTwoOverPi = 0.636619772367581f;
  const float kPio2Hi = 1.5707963705062866f;   // (float)(pi/2)
  const float kPio2Mid = -4.371139000186243e-08f;  // (float)(pi/2 - hi)
  const float kPio2Lo = -1.7151245100059e-15f;     // (float)(pi/2 - hi - mid)
  float t = fmaf(a, kTwoOverPi, kMagic);
  uint32_t q = float_to_bits(t);  // low mantissa bits hold round(a*2/pi) mod 4 (two's complement safe)
  float k = t - kMagic;
  float r = fmaf(k, -kPio2Hi, a);
  r = fmaf(k, -kPio2Mid, r);
  r = fmaf(k, -kPio2Lo, r);
  float sn, cs;
  sincos_poly(r, sn, cs);
  float so = (q & 1u) ? cs : sn;
  float co = (q & 1u) ? sn : cs;
  s = (q & 2u) ? -so : so;
  c = ((q + 1u) & 2u) ? -co : co;
}
// Same function, same bits: for |a| <= 0.78 (< pi/4) the reduction is the identity (k = 0, r = a, quadrant 0), so
// the polynomial is evaluated directly -- the common case for steering / direction angles (|a| <= 30 deg).
// SMALL = the caller has proved |a| <= kSmallAngle (bounded controls): same bits, no test.
constexpr float kSmallAngle = 0.78f;
template <bool SMALL = false>
MPPI_HD void sincos_f32(float a, float &s, float &c) {
  if (SMALL || fabsf(a) <= kSmallAngle) {
    sincos_poly(a, s, c);
  } else {
    sincos_reduce(a, s, c);
  }
}

// sin and cos of a per-step angle INCREMENT (w*dt, roll_v*dt, pitch_v*dt).  For |a| <= 0.35 rad a degree-5 / degree-6
// minimax pair fitted on that interval (|sin error| < 1.5e-8 relative, |cos error| < 1.2e-10, both below half an
// FP32 ulp); beyond that
// sincos_f32.  Part of the FP32 contract (host twin and device take the same branch on the same bits).
constexpr float kSmallIncrement = 0.35f;
template <bool SMALL = false>
MPPI_HD void sincos_increment(float a, float &s, float &c) {
  if (SMALL || fabsf(a) <= kSmallIncrement) {
    float z = a * a;
    s = fmaf(fmaf(z, 8.299172942064573e-3f, -1.6666535808051738e-1f), z * a, a);
    c = fmaf(fmaf(fmaf(z, -1.3841075014937032e-3f, 4.166644728107707e-2f), z, -0.5f), z, 1.0f);
  } else {
    sincos_f32(a, s, c);
  }
}

// (c, s) <- (c, s) rotated by the angle whose cosine / sine are (cd, sd): the heading recurrence.
// predict_NextState (DD:106-108) evaluates cos(yaw), sin(yaw) with yaw_{t+1} = yaw_t + w_t*dt; the contract carries
// (cos yaw, sin yaw) through the horizon instead of re-deriving them from the accumulated angle every step:
// 4 operations + a short polynomial per step instead of a full range-reduced sincos.  The error of the pair grows
// like that of the FP32 yaw accumulation it replaces (~1e-7 per step; measured against the FP64 oracle in
// tests/test_oracle.py).
MPPI_HD void rotate(float &c, float &s, float cd, float sd) {
  float c2 = fmaf(c, cd, -(s * sd));
  float s2 = fmaf(s, cd, c * sd);
  c = c2;
  s = s2;
}
template <bool SMALL = false>
MPPI_HD void rotate_by(float &c, float &s, float angle) {
  float sd, cd;
  sincos_increment<SMALL>(angle, sd, cd);
  rotate(c, s, cd, sd);
}

// True when every per-step angle of this solve is provably inside the polynomial ranges: the controls are clamped
// to [u_min, u_max], and |clamp(v) * dt| <= max(|u_min|, |u_max|) * dt holds in FP32 as well (rounding is
// monotonic).  The kernels then run the instantiation without the range tests -- same bits as the tested path.
MPPI_HD bool angles_are_small(const SolveParams &P) {
  bool ok = true;
  auto bound = [&](int u) { return fmaxf(fabsf(P.u_min[u]), fabsf(P.u_max[u])); };
  ok = ok && (bound(1) * fabsf(P.dt) <= kSmallIncrement);
  if (P.model != kDiffDrive) ok = ok && (bound(2) <= kSmallAngle);
  if (P.model == kFullBody) ok = ok && (bound(3) * fabsf(P.dt) <= kSmallIncrement) && (bound(4) * fabsf(P.dt) <= kSmallIncrement);
  return ok;  // any NaN parameter compares false
}

// atan2 in the FP32 contract (same bits on host and device): octant reduction + the classic single-precision
// arctangent polynomial on [0, tan(pi/8)]; |error| < 3e-7 rad.  atan2_f32(0, 0) = 0 like atan2 (duplicate window
// points, DD:175-178).  Used for yaw_ref_[0] of the full-body yaw term (FB:408), evaluated on the robot-centred
// FP32 window; non-finite inputs are not specified.
MPPI_HD float atan2_f32(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = ax > ay ? ax : ay, mn = ax > ay ? ay : ax;
  float a = mx > 0.f ? mn / mx : 0.f;  // in [0, 1]
  float base = 0.f;
  if (a > 0.4142135623730950f) {  // tan(pi/8)
    base = 0.7853981633974483f;
    a = (a - 1.0f) / (a + 1.0f);
  }
  const float z = a * a;
  float p = fmaf(z, 8.05374449538e-2f, -1.38776856032e-1f);
  p = fmaf(p, z, 1.99777106478e-1f);
  p = fmaf(p, z, -3.33329491539e-1f);
  float r = base + fmaf(p * z, a, a);
  if (ay > ax) r = 1.5707963267948966f - r;
  if (x < 0.f) r = 3.141592653589793f - r;
  return y < 0.f ? -r : r;
}

// yaw_ref_[0] = atan2(y_ref[1] - y_ref[0], x_ref[1] - x_ref[0]) (DD:175-178) from the FP32 window
MPPI_HD float yaw_ref0_f32(float x0, float y0, float x1, float y1) { return atan2_f32(y1 - y0, x1 - x0); }

// squared distance of calc_MinDistance's inner expression (DD:188) without the sqrt
MPPI_HD float dist2(float x, float y, float xr, float yr) {
  float dx = x - xr;
  float dy = y - yr;
  return fmaf(dy, dy, dx * dx);
}

// Literal first-minimum scan over the T window points (DD:185-190). win = {x_ref, y_ref} pairs.
// Returns min(d^2, 1e4); *arg = first index attaining it, or -1 when no point is closer than 100 m.
template <typename Win>
MPPI_HD float min_dist2_literal(float x, float y, const Win &win, int T, int *arg) {
  float best = kDist2Cap;
  int bi = -1;
  for (int j = 0; j < T; ++j) {
    float d2 = dist2(x, y, win.x(j), win.y(j));
    if (d2 < best) {
      best = d2;
      bi = j;
    }
  }
  if (arg) *arg = bi;
  return best;
}

// Orientation of one sample as the kernels carry it: the angles themselves (debug taps only -- dead code in the
// production kernel) and the cos / sin pairs advanced by rotate_by().
template <int MODEL>
struct Attitude {
  float yaw, cy, sy;
  float roll, cr, sr;    // full body only
  float pitch, cp, sp;   // full body only
  MPPI_HD void init(const float *state0) {
    yaw = state0[2];
    sincos_f32(yaw, sy, cy);
    roll = pitch = 0.f;
    cr = cp = 1.f;
    sr = sp = 0.f;
    if (MODEL == kFullBody) {
      roll = state0[3];
      pitch = state0[4];
      sincos_f32(roll, sr, cr);
      sincos_f32(pitch, sp, cp);
    }
  }
};

// predict_NextState (DD:104-109 / SD:120-125 / FB:445-452) for one sample; u = controls of this step.
// Heading = yaw (DD) or yaw + steer/direction (SD:122, FB:447): cos/sin(yaw + d) by the addition theorem from
// the carried (cos yaw, sin yaw) and (sd, cd) = sincos_f32(d), which the caller shares with the ZMP model
// (FB:473-474); (sd, cd) is ignored for DD.
// The three pieces are separate so that a kernel can evaluate the position update with packed instructions
// (same IEEE operations per element).
template <int MODEL>
MPPI_HD void step_heading(const Attitude<MODEL> &a, float sd, float cd, float &ch, float &sh) {
  ch = a.cy;
  sh = a.sy;
  if (MODEL != kDiffDrive) rotate(ch, sh, cd, sd);
}
MPPI_HD void step_position(float &x, float &y, float v, float ch, float sh, float dt) {
  x = fmaf(v * ch, dt, x);
  y = fmaf(v * sh, dt, y);
}
template <int MODEL, bool SMALL = false>
MPPI_HD void step_attitude(Attitude<MODEL> &a, const float *u, float dt) {
  a.yaw = fmaf(u[1], dt, a.yaw);
  rotate_by<SMALL>(a.cy, a.sy, u[1] * dt);
  if (MODEL == kFullBody) {
    a.roll = fmaf(u[3], dt, a.roll);    // FB:450
    a.pitch = fmaf(u[4], dt, a.pitch);  // FB:451
    rotate_by<SMALL>(a.cr, a.sr, u[3] * dt);
    rotate_by<SMALL>(a.cp, a.sp, u[4] * dt);
  }
}
template <int MODEL, bool SMALL = false>
MPPI_HD void step_state(float &x, float &y, Attitude<MODEL> &a, const float *u, float dt, float sd, float cd) {
  float ch, sh;
  step_heading<MODEL>(a, sd, cd, ch, sh);
  step_position(x, y, u[0], ch, sh, dt);
  step_attitude<MODEL, SMALL>(a, u, dt);
}

// zmp_y of FB:468-486 + FB:597-603 for one (sample, t):
//   CoM = (b sin pitch, -b sin roll, b cos pitch cos roll);  a = (ax, ay, 0);  HGdot = I (omega+ - omega)/dt
//   zmp_y = CoM_y + CoM_z * ay / g_z - HGdot_x / (m g_z)
// zmp_x (unused by the cost) is returned for tests:  zmp_x = CoM_x + CoM_z * ax / g_z + HGdot_y / (m g_z)
// (sd, cd) = sin/cos of direction_[t]; (sr, cr), (sp, cp) = sin/cos of roll_[t], pitch_[t].
MPPI_HD void zmp_model(const SolveParams &P, float v0, float v1, float w0, float sd, float cd, float rv0, float rv1,
                       float pv0, float pv1, float sr, float cr, float sp, float cp, float &zmp_x, float &zmp_y) {
  float drive_accel = (v1 - v0) * P.inv_dt;
  float ac = v0 * w0;
  float ax = drive_accel * cd - ac * sd;
  float ay = fmaf(drive_accel, sd, ac * cd);
  float hgd_x = P.ixx * ((rv1 - rv0) * P.inv_dt);
  float hgd_y = P.iyy * ((pv1 - pv0) * P.inv_dt);
  float com_x = P.base2com * sp;
  float com_y = -(P.base2com * sr);
  float com_z = P.base2com * cp * cr;
  zmp_x = fmaf(com_z * ax, P.inv_gz, fmaf(hgd_y, P.inv_mgz, com_x));
  zmp_y = fmaf(com_z * ay, P.inv_gz, fmaf(-hgd_x, P.inv_mgz, com_y));
}

// Per-sample cost accumulators.  Each term is summed sequentially in t in its own accumulator and the
// weights are applied once at the end, so every kernel variant (literal / transposed-chunk) and the host twin
// can reproduce the same bits regardless of how they interleave the terms.
struct CostAcc {
  float path = 0.f;    // sum_t min d^2                        (DD:201,206 / FB:411)
  float vel = 0.f;     // sum_t (v_t - v_ref)^2                (DD:204 with D1 / FB:413)
  float zmp = 0.f;     // sum_t zmp_y^2                        (FB:416)
  float droll = 0.f;   // sum_t (roll_v[t+1] - roll_v[t])^2    (FB:418)
  float back = 0.f;    // sum_t [v_t < 0] v_t^2                (FB:420)
};

MPPI_HD float combine_cost(const SolveParams &P, const CostAcc &a, float yaw0_err) {
  float c = P.path_weight * a.path;
  c = fmaf(P.v_weight, a.vel, c);
  if (P.model == kFullBody) {
    c = fmaf(P.zmp_weight, a.zmp, c);
    c = fmaf(P.roll_v_weight, a.droll, c);
    c = fmaf(P.back_weight, a.back, c);
    c = fmaf(P.yaw_weight, yaw0_err * yaw0_err, c);  // FB:408 (same for every sample)
  }
  return c;
}

// The whole per-sample path in its literal loop structure: sampling (D5) -> predict_States -> calc_Cost.
//   Eps::get(t, u)  -> standard normal of this sample at control step t
//   Nom::get(t, u)  -> previous optimal_solution (warm start, not time shifted)
//   Win::x(j)/y(j)  -> window point j, robot-centred frame
//   Sink::state(t, x, y, yaw, roll, pitch), Sink::nearest(t, j, d2), Sink::control(t, u, value),
//   Sink::zmp(t, zx, zy)  -> debug taps, no-ops in the production kernel
// state0 = {0, 0, yaw, roll, pitch} in the robot-centred frame; yaw_ref0 = yaw_ref0_f32 of the window (T >= 2).
template <int MODEL, typename Eps, typename Nom, typename Win, typename Sink>
MPPI_HD float rollout_cost_literal(const SolveParams &P, const float *state0, float yaw_ref0, const Eps &eps,
                                   const Nom &nom, const Win &win, Sink &sink) {
  constexpr int U = MODEL == kDiffDrive ? 2 : (MODEL == kSteering ? 3 : 5);
  const int T = P.T;
  const int Tc = num_cost_states(MODEL, T);
  float x = state0[0], y = state0[1];
  Attitude<MODEL> att;
  att.init(state0);
  CostAcc acc;
  float cur[U], nxt[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    cur[u] = 0.f;
    nxt[u] = 0.f;
  }
  auto draw = [&](int t, float *dst) {
#pragma unroll
    for (int u = 0; u < U; ++u) dst[u] = sample_control(eps.get(t, u), P.sigma, nom.get(t, u), P.u_min[u], P.u_max[u]);
    if (MODEL == kFullBody && P.steer_off) dst[2] = 0.f;  // FB:517
#pragma unroll
    for (int u = 0; u < U; ++u) sink.control(t, u, dst[u]);
  };
  if (T > 1) draw(0, cur);
  for (int t = 0; t < T; ++t) {
    sink.state(t, x, y, att.yaw, att.roll, att.pitch);
    const bool has_step = t < T - 1;
    if (MODEL == kFullBody && t + 1 < T - 1) draw(t + 1, nxt);
    if (t < Tc) {
      int arg;
      float d2 = min_dist2_literal(x, y, win, T, &arg);
      sink.nearest(t, arg, d2);
      acc.path += d2;
    }
    float sd = 0.f, cd = 1.f;  // sin / cos of steer_[t] / direction_[t]
    if (MODEL != kDiffDrive && has_step) sincos_f32(cur[2], sd, cd);
    if (MODEL != kFullBody) {
      if (has_step) {
        float dv = cur[0] - P.v_ref;
        acc.vel = fmaf(dv, dv, acc.vel);
      }
    } else if (t < Tc) {
      float dv = cur[0] - P.v_ref;
      acc.vel = fmaf(dv, dv, acc.vel);
      float zx, zy;
      zmp_model(P, cur[0], nxt[0], cur[1], sd, cd, cur[3], nxt[3], cur[4], nxt[4], att.sr, att.cr, att.sp, att.cp, zx,
                zy);
      sink.zmp(t, zx, zy);
      acc.zmp = fmaf(zy, zy, acc.zmp);
      float dr = nxt[3] - cur[3];
      acc.droll = fmaf(dr, dr, acc.droll);
      if (cur[0] < 0.f) acc.back = fmaf(cur[0], cur[0], acc.back);
    }
    if (has_step) {
      step_state<MODEL>(x, y, att, cur, P.dt, sd, cd);
      if (MODEL == kFullBody) {
#pragma unroll
        for (int u = 0; u < U; ++u) cur[u] = nxt[u];
      } else if (t + 1 < T - 1) {
        draw(t + 1, cur);
      }
    }
  }
  return combine_cost(P, acc, MODEL == kFullBody ? state0[2] - yaw_ref0 : 0.f);
}

struct NullSink {
  MPPI_HD void state(int, float, float, float, float, float) {}
  MPPI_HD void nearest(int, int, float) {}
  MPPI_HD void control(int, int, float) {}
  MPPI_HD void zmp(int, float, float) {}
};

}  // namespace mppi
