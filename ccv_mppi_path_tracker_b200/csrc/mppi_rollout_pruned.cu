// mppi_rollout_pruned.cu -- K2, production variant: fused sampling-clamp + rollout + cost with an EXACT pruned
// nearest-window-point scan.  One thread per sample; nothing per-step leaves the chip.
//
// What it replaces: clamp (DD:98-99) + predict_States (DD:111-122) + calc_Cost / calc_MinDistance (DD:183-210)
// [+ the ZMP loop FB:468-486 and cost FB:404-424], reference = /root/reference/src/{diff_drive,steering_diff_drive,
// full_body}_mppi.cpp.  Per-sample arithmetic is the FP32 contract of mppi_math.h; the result is bit-identical to
// rollout_cost_literal_kernel (and to the host twin) for every input -- the scan only SKIPS window points that
// provably cannot be the minimum:
//
//   * the T window points are grouped into leaves of 4 consecutive points (padded with copies of the last point);
//     each leaf, each prefix [0,b) and each suffix [b,NB) of leaves has a bounding circle (centre c, radius rho),
//     built once per CTA in shared memory;
//   * per state, a thread scans a window of kWin consecutive leaves around the leaf that held its previous
//     minimum (temporal coherence: a rollout moves <= v_max*dt per step), giving a candidate minimum `best`;
//   * every point outside that window lies in one of kMid leaves either side or in the prefix / suffix beyond;
//     a node is discarded when |p - c| > sqrt(best) + rho (triangle inequality), evaluated in the squared domain
//     with a 2^-17 relative safety margin (>> the few-ulp rounding of the test itself), so a discarded point
//     always has fl(d^2) > best;
//   * if any node cannot be discarded the thread falls back to the full scan for that state.
// min() is exact and order-independent, so the accumulated path cost has the same bits as the literal scan.
// Cost per state is independent of T (about 12 exact distances + 6 circle tests instead of T distances).
//
// The distance evaluation uses Blackwell's packed FP32 pipe (sub/mul/fma .f32x2 -> FADD2/FMUL2/FFMA2, two window
// points per instruction, same IEEE roundings per element) and the 3-input FMNMX3.
#include "mppi_device.cuh"

namespace mppi {

namespace {

constexpr int kLeaf = 4;   // window points per leaf
constexpr int kWin = 3;    // leaves scanned around the tracked minimum
constexpr int kMid = 2;    // leaves tested individually either side of the scanned window
constexpr float kInflate = 1.0f + 1.0f / 262144.0f;  // 1 + 2^-18 on sqrt(best) and on every radius
constexpr float kFar = 1.0e18f;                      // centre of an empty node: never within reach

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// {d2(p, r0), d2(p, r1)} with the roundings of mppi::dist2: dx = x - xr; dy = y - yr; fma(dy, dy, dx*dx)
__device__ __forceinline__ u64 dist2_pair(u64 xx, u64 yy, u64 xr, u64 yr) {
  u64 dx, dy, m, d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(xx), "l"(xr));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(yy), "l"(yr));
  asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(m) : "l"(dx));
  asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(d) : "l"(dy), "l"(m));
  return d;
}
__device__ __forceinline__ float min3f(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float sqrt_approx(float a) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}

// minimum squared distance to the 4 points of one leaf; leaf = {x0,x1,y0,y1}, {x2,x3,y2,y3}
__device__ __forceinline__ float leaf_min(const float4 *leaf, u64 xx, u64 yy) {
  const float4 a = leaf[0], b = leaf[1];
  float a0, a1, b0, b1;
  unpack2(dist2_pair(xx, yy, pack2(a.x, a.y), pack2(a.z, a.w)), a0, a1);
  unpack2(dist2_pair(xx, yy, pack2(b.x, b.y), pack2(b.z, b.w)), b0, b1);
  return min3f(a0, a1, fminf(b0, b1));
}

// true when no point inside the node can be closer than sqrt(best): |p - c| > s + rho, squared domain.
// node = {cx, cy, rho * kInflate, -}; s_up = sqrt(best) * kInflate (rounded-up estimate)
__device__ __forceinline__ bool node_excluded(const float4 node, float x, float y, float s_up) {
  const float dx = x - node.x, dy = y - node.y;
  const float D2 = fmaf(dy, dy, dx * dx);
  const float t = s_up + node.z;
  return D2 > t * t;
}

struct ScanTables {
  const float4 *pts;   // [2*NB]  leaf b = pts[2b], pts[2b+1]
  const float4 *leaf;  // [NB + 2*kMid]  circle of leaf b at leaf[b + kMid]; empties either side
  const float4 *pre;   // [NB + 1]  circle of leaves [0, b)
  const float4 *suf;   // [NB + 1]  circle of leaves [b, NB)
  int NB;
};

// Exact min_j min(d2(p, r_j), 1e4) over the whole window.  wl = first leaf of this thread's scan window (in/out).
__device__ __forceinline__ float min_dist2_pruned(const ScanTables &tb, float x, float y, int &wl) {
  const u64 xx = pack2(x, x), yy = pack2(y, y);
  const float4 *w = tb.pts + 2 * wl;
  float m[kWin];
#pragma unroll
  for (int k = 0; k < kWin; ++k) m[k] = leaf_min(w + 2 * k, xx, yy);
  float best = kDist2Cap;
#pragma unroll
  for (int k = 0; k < kWin; ++k) best = fminf(best, m[k]);
  const float s_up = sqrt_approx(best) * kInflate;
  bool ok = node_excluded(tb.pre[max(wl - kMid, 0)], x, y, s_up);
  ok = ok && node_excluded(tb.suf[min(wl + kWin + kMid, tb.NB)], x, y, s_up);
#pragma unroll
  for (int k = 0; k < kMid; ++k) {
    ok = ok && node_excluded(tb.leaf[wl + k], x, y, s_up);                       // leaves wl-kMid .. wl-1
    ok = ok && node_excluded(tb.leaf[wl + kMid + kWin + k], x, y, s_up);         // leaves wl+kWin .. wl+kWin+kMid-1
  }
  if (ok) {
    // keep the minimum in the middle of the window
    if (m[0] == best && wl > 0) --wl;
    else if (m[kWin - 1] == best && m[kWin / 2] != best && wl + kWin < tb.NB) ++wl;
    return best;
  }
  // fallback: every leaf (still exact), re-centre on the leaf that holds the minimum
  best = kDist2Cap;
  int bl = wl + kWin / 2;
  for (int b = 0; b < tb.NB; ++b) {
    const float mb = leaf_min(tb.pts + 2 * b, xx, yy);
    if (mb < best) {
      best = mb;
      bl = b;
    }
  }
  wl = min(max(bl - kWin / 2, 0), tb.NB - kWin);
  return best;
}

// bounding circle of window points [j0, j1) (indices into the padded point list): bbox centre, max distance
__device__ float4 bounding_circle(const float2 *pt, int j0, int j1) {
  if (j1 <= j0) return make_float4(kFar, kFar, 0.f, 0.f);
  float xmin = pt[j0].x, xmax = xmin, ymin = pt[j0].y, ymax = ymin;
  for (int j = j0 + 1; j < j1; ++j) {
    xmin = fminf(xmin, pt[j].x);
    xmax = fmaxf(xmax, pt[j].x);
    ymin = fminf(ymin, pt[j].y);
    ymax = fmaxf(ymax, pt[j].y);
  }
  const float cx = 0.5f * (xmin + xmax), cy = 0.5f * (ymin + ymax);
  float r2 = 0.f;
  for (int j = j0; j < j1; ++j) r2 = fmaxf(r2, dist2(cx, cy, pt[j].x, pt[j].y));
  // a NaN / inf window coordinate makes the node unusable: force the fallback scan
  if (!(r2 < 1.0e30f)) return make_float4(0.f, 0.f, 1.0e18f, 0.f);
  return make_float4(cx, cy, sqrtf(r2) * kInflate + 1.0e-30f, 0.f);
}

}  // namespace

size_t pruned_smem_bytes(int T, int planes) {
  const int NB = (T + kLeaf - 1) / kLeaf;
  // raw points (float2 x NB*4) | packed leaves (float4 x 2NB) | leaf circles | prefix | suffix | nominal
  return sizeof(float2) * (size_t)NB * kLeaf + sizeof(float4) * ((size_t)2 * NB + (NB + 2 * kMid) + 2 * (NB + 1)) +
         sizeof(float) * (size_t)planes;
}

template <int MODEL>
__global__ void __launch_bounds__(128)
    rollout_cost_pruned_kernel(const SolveHeader *__restrict__ hdr, const float *__restrict__ eps,
                               const float *__restrict__ nominal, const float *__restrict__ window,
                               const float *__restrict__ state, float *__restrict__ cost,
                               unsigned int *__restrict__ cmin, int K, int Kp, int planes, int win_stride, int T) {
  constexpr int U = MODEL == kDiffDrive ? 2 : (MODEL == kSteering ? 3 : 5);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ SolveParams sP;
  __shared__ float s_red[32];
  const int robot = blockIdx.y;
  const int NB = (T + kLeaf - 1) / kLeaf;
  float4 *s_pts = reinterpret_cast<float4 *>(smem_raw);
  float4 *s_leaf = s_pts + 2 * NB;
  float4 *s_pre = s_leaf + NB + 2 * kMid;
  float4 *s_suf = s_pre + NB + 1;
  float2 *s_raw = reinterpret_cast<float2 *>(s_suf + NB + 1);
  float *s_nom = reinterpret_cast<float *>(s_raw + NB * kLeaf);

  // ---- prologue: stage the window, build the bounding circles -------------------------------------------
  const float *g_win = window + (size_t)robot * win_stride;
  load_params_to_shared(&sP, hdr);
  for (int j = threadIdx.x; j < NB * kLeaf; j += blockDim.x) {
    const int js = min(j, T - 1);  // pad with copies of the last point (does not change any minimum)
    s_raw[j] = make_float2(g_win[2 * js], g_win[2 * js + 1]);
  }
  for (int j = threadIdx.x; j < planes; j += blockDim.x) s_nom[j] = nominal[(size_t)robot * planes + j];
  __syncthreads();
  for (int b = threadIdx.x; b < 2 * NB; b += blockDim.x) {
    const float2 p0 = s_raw[2 * b], p1 = s_raw[2 * b + 1];
    s_pts[b] = make_float4(p0.x, p1.x, p0.y, p1.y);
  }
  for (int k = threadIdx.x; k < NB + 2 * kMid; k += blockDim.x) {
    const int b = k - kMid;
    s_leaf[k] = (b >= 0 && b < NB) ? bounding_circle(s_raw, b * kLeaf, (b + 1) * kLeaf) : make_float4(kFar, kFar, 0.f, 0.f);
  }
  // prefix / suffix circles: thread pairs from the top so the long ones do not all land in warp 0
  for (int k = blockDim.x - 1 - threadIdx.x; k < 2 * (NB + 1); k += blockDim.x) {
    const int b = k >> 1;
    if (k & 1) s_suf[b] = bounding_circle(s_raw, b * kLeaf, NB * kLeaf);
    else s_pre[b] = bounding_circle(s_raw, 0, b * kLeaf);
  }
  __syncthreads();

  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float c = 0.f;
  if (i < K) {
    const ScanTables tb{s_pts, s_leaf, s_pre, s_suf, NB};
    const float *st = state + (size_t)robot * 8;
    const float *e = eps + (size_t)robot * planes * Kp + i;
    const int steps = T - 1;
    const int Tc = num_cost_states(MODEL, T);
    float x = st[0], y = st[1], yaw = st[2];
    float roll = MODEL == kFullBody ? st[3] : 0.f;
    float pitch = MODEL == kFullBody ? st[4] : 0.f;
    const float sigma = sP.sigma, dt = sP.dt, v_ref = sP.v_ref;
    const bool steer_off = MODEL == kFullBody && sP.steer_off;
    CostAcc acc;
    float cur[U], nxt[U], raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) cur[u] = nxt[u] = raw[u] = 0.f;
    auto load_raw = [&](int t) {
      if (t < steps) {
#pragma unroll
        for (int u = 0; u < U; ++u) raw[u] = __ldcs(e + (size_t)(t * U + u) * Kp);
      }
    };
    auto make = [&](int t, float *dst) {  // sampling (D5) from the prefetched normals of step t
      if (t < steps) {
#pragma unroll
        for (int u = 0; u < U; ++u) dst[u] = sample_control(raw[u], sigma, s_nom[t * U + u], sP.u_min[u], sP.u_max[u]);
        if (steer_off) dst[2] = 0.f;  // FB:517
      }
    };
    load_raw(0);
    make(0, cur);
    load_raw(1);
    make(1, nxt);
    load_raw(2);
    int wl = 0;
    for (int t = 0; t < T; ++t) {
      const bool has_step = t < steps;
      if (t < Tc) acc.path += min_dist2_pruned(tb, x, y, wl);
      if (MODEL != kFullBody) {
        if (has_step) {
          const float dv = cur[0] - v_ref;
          acc.vel = fmaf(dv, dv, acc.vel);
        }
      } else if (t < Tc) {
        const float dv = cur[0] - v_ref;
        acc.vel = fmaf(dv, dv, acc.vel);
        float zx, zy;
        zmp_model(sP, cur[0], nxt[0], cur[1], cur[2], cur[3], nxt[3], cur[4], nxt[4], roll, pitch, zx, zy);
        acc.zmp = fmaf(zy, zy, acc.zmp);
        const float dr = nxt[3] - cur[3];
        acc.droll = fmaf(dr, dr, acc.droll);
        if (cur[0] < 0.f) acc.back = fmaf(cur[0], cur[0], acc.back);
      }
      if (has_step) {
        const float heading = MODEL == kDiffDrive ? yaw : yaw + cur[2];
        step_pose(x, y, yaw, cur[0], cur[1], heading, dt);
        if (MODEL == kFullBody) {
          roll = fmaf(cur[3], dt, roll);
          pitch = fmaf(cur[4], dt, pitch);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) cur[u] = nxt[u];
      make(t + 2, nxt);
      load_raw(t + 3);
    }
    c = combine_cost(sP, acc, MODEL == kFullBody ? st[2] - st[5] : 0.f);
    cost[(size_t)robot * K + i] = c;
  }
  block_min_to_global(c, i < K, cmin + robot, s_red);
}

cudaError_t launch_rollout_cost_pruned(const DeviceState &d, cudaStream_t s) {
  dim3 grid((d.K + 127) / 128, d.R);
  const size_t smem = pruned_smem_bytes(d.T, d.planes);
#define MPPI_LAUNCH_PRUNED(M)                                                                                    \
  do {                                                                                                           \
    if (smem > 48 * 1024) {                                                                                      \
      cudaError_t e = cudaFuncSetAttribute(rollout_cost_pruned_kernel<M>,                                        \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
      if (e != cudaSuccess) return e;                                                                            \
    }                                                                                                            \
    rollout_cost_pruned_kernel<M><<<grid, 128, smem, s>>>(d.hdr, d.eps, d.nominal, d.window, d.state, d.cost,    \
                                                          d.cmin, d.K, d.Kp, d.planes, d.win_stride, d.T);       \
  } while (0)
  switch (d.model) {
    case kDiffDrive: MPPI_LAUNCH_PRUNED(kDiffDrive); break;
    case kSteering: MPPI_LAUNCH_PRUNED(kSteering); break;
    default: MPPI_LAUNCH_PRUNED(kFullBody); break;
  }
#undef MPPI_LAUNCH_PRUNED
  return cudaGetLastError();
}

// worth it (and the tables fit) only for windows of more than a few leaves
bool pruned_scan_supported(int T, int planes) {
  return (T + kLeaf - 1) / kLeaf >= kWin + 2 * kMid && pruned_smem_bytes(T, planes) <= 200 * 1024;
}

}  // namespace mppi
