// mppi_kernels.cu -- hand-written sm_100a kernels of the MPPI solve.
//   K1 noise_kernel              Philox4x32-10 + Box-Muller, float4 stores      (replaces sampling() RNG, DD:83-97)
//   K2 rollout_cost_*_kernel     sampling clamp + predict_States + calc_Cost    (DD:98-122, DD:183-210, FB:404-489)
//   K3 weights_kernel            exp(-(c - c_min)/lambda) + partial sums        (DD:216-222, min-shifted: D3)
//   K4 weighted_controls_kernel  sum_i w_i * clamp(u* + sigma*eps)              (DD:228-236, D2)
//   K5 finalize_kernel / K6 merge_kernel   fixed-order final sums, cross-rank merge, u_new = N / S
// Reference citations are relative to /root/reference/src (DD = diff_drive_mppi.cpp, FB = full_body_mppi.cpp).
// Compiled with -fmad=false: all FMAs are explicit (see mppi_math.h).
#include "mppi_kernels.h"

#include "mppi_device.cuh"
#include "philox.h"

namespace mppi {

// ---------------------------------------------------------------------------------------------------------
// K1: noise
// ---------------------------------------------------------------------------------------------------------
// grid = (quad blocks, planes, robots); one thread = one Philox block = 4 consecutive samples of one plane.
__global__ void __launch_bounds__(256)
    noise_kernel(const SolveHeader *__restrict__ hdr, const uint32_t *__restrict__ counter, float *__restrict__ eps,
                 unsigned int *__restrict__ cmin, int Kq, int Kp, int planes) {
  const int plane = blockIdx.y;
  const int robot = blockIdx.z;
  if (blockIdx.x == 0 && plane == 0 && threadIdx.x == 0) cmin[robot] = 0xFFFFFFFFu;
  const uint32_t key0 = hdr->key0, key1 = hdr->key1;
  const uint32_t c2 = hdr->robot_offset + (uint32_t)robot, c3 = *counter, q_off = hdr->q_offset;
  float4 *out = reinterpret_cast<float4 *>(eps + ((size_t)robot * planes + plane) * Kp);
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < Kq; q += gridDim.x * blockDim.x) {
    Philox4 r = philox4x32_10(q_off + (uint32_t)q, (uint32_t)plane, c2, c3, key0, key1);
    // Box-Muller: u1 in (0,1), angle in [-pi, pi)
    const float kInv24 = 5.9604644775390625e-08f;       // 2^-24
    const float kHalfUlp = 2.98023223876953125e-08f;    // 2^-25
    const float kPiOver2p31 = 1.4629180792671596e-09f;  // pi * 2^-31
    float u1a = fmaf((float)(r.v[0] >> 8), kInv24, kHalfUlp);
    float u1b = fmaf((float)(r.v[2] >> 8), kInv24, kHalfUlp);
    float ra = sqrtf(-2.0f * __logf(u1a));
    float rb = sqrtf(-2.0f * __logf(u1b));
    float sa, ca, sb, cb;
    __sincosf((float)(int32_t)r.v[1] * kPiOver2p31, &sa, &ca);
    __sincosf((float)(int32_t)r.v[3] * kPiOver2p31, &sb, &cb);
    out[q] = make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
  }
}

__global__ void reset_cmin_kernel(unsigned int *cmin, int R) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R) cmin[i] = 0xFFFFFFFFu;
}

cudaError_t launch_noise(const DeviceState &d, cudaStream_t s) {
  const int Kq = d.Kp / 4;
  int bx = (Kq + 255) / 256;
  if (bx > 2048) bx = 2048;
  dim3 grid(bx, d.planes, d.R);
  noise_kernel<<<grid, 256, 0, s>>>(d.hdr, d.counter, d.eps, d.cmin, Kq, d.Kp, d.planes);
  return cudaGetLastError();
}

cudaError_t launch_reset_cmin(const DeviceState &d, cudaStream_t s) {
  reset_cmin_kernel<<<(d.R + 255) / 256, 256, 0, s>>>(d.cmin, d.R);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// K2 (literal): one thread per sample, window + warm start staged in shared memory
// ---------------------------------------------------------------------------------------------------------
struct EpsGlobal {
  const float *base;  // eps of this robot, offset by the sample index
  size_t plane_stride;
  int U;
  __device__ __forceinline__ float get(int t, int u) const { return __ldg(base + (size_t)(t * U + u) * plane_stride); }
};
struct NomShared {
  const float *s;
  int U;
  __device__ __forceinline__ float get(int t, int u) const { return s[t * U + u]; }
};
struct WinShared {
  const float2 *w;
  __device__ __forceinline__ float x(int j) const { return w[j].x; }
  __device__ __forceinline__ float y(int j) const { return w[j].y; }
};
struct NearestSink {
  int *row;  // [T] of this sample or nullptr
  __device__ __forceinline__ void state(int, float, float, float, float, float) {}
  __device__ __forceinline__ void nearest(int t, int j, float) {
    if (row) row[t] = j;
  }
  __device__ __forceinline__ void control(int, int, float) {}
  __device__ __forceinline__ void zmp(int, float, float) {}
};

template <int MODEL>
__global__ void __launch_bounds__(128)
    rollout_cost_literal_kernel(const SolveHeader *__restrict__ hdr, const float *__restrict__ eps,
                                const float *__restrict__ nominal, const float *__restrict__ window,
                                const float *__restrict__ state, float *__restrict__ cost,
                                unsigned int *__restrict__ cmin, int *__restrict__ nearest, int K, int Kp, int planes,
                                int win_stride, int T) {
  extern __shared__ __align__(16) float smem[];
  __shared__ SolveParams sP;
  __shared__ float s_red[32];
  const int robot = blockIdx.y;
  float2 *s_win = reinterpret_cast<float2 *>(smem);
  float *s_nom = smem + 2 * T;
  const float *g_win = window + (size_t)robot * win_stride;
  load_params_to_shared(&sP, hdr);
  for (int j = threadIdx.x; j < T; j += blockDim.x) s_win[j] = make_float2(g_win[2 * j], g_win[2 * j + 1]);
  for (int j = threadIdx.x; j < planes; j += blockDim.x) s_nom[j] = nominal[(size_t)robot * planes + j];
  __syncthreads();

  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float c = 0.f;
  if (i < K) {
    const float *st = state + (size_t)robot * 8;
    float state0[5] = {st[0], st[1], st[2], st[3], st[4]};
    EpsGlobal e{eps + (size_t)robot * planes * Kp + i, (size_t)Kp, sP.U};
    NomShared n{s_nom, sP.U};
    WinShared w{s_win};
    NearestSink sink{nearest ? nearest + ((size_t)robot * K + i) * T : nullptr};
    c = rollout_cost_literal<MODEL>(sP, state0, st[5], e, n, w, sink);
    cost[(size_t)robot * K + i] = c;
  }
  block_min_to_global(c, i < K, cmin + robot, s_red);
}

cudaError_t launch_rollout_cost(const DeviceState &d, int scan_mode, bool write_nearest, cudaStream_t s) {
  // scan_mode 2 = exact pruned scan (production); 1 = literal (every window point; the one that records argmin)
  if (scan_mode == 2 && !write_nearest && pruned_scan_supported(d.T, d.planes)) return launch_rollout_cost_pruned(d, s);
  dim3 grid((d.K + 127) / 128, d.R);
  size_t smem = sizeof(float) * (2 * (size_t)d.T + d.planes);
  int *nearest = write_nearest ? d.nearest : nullptr;
#define MPPI_LAUNCH_LITERAL(M)                                                                                   \
  do {                                                                                                           \
    if (smem > 48 * 1024) {                                                                                      \
      cudaError_t e = cudaFuncSetAttribute(rollout_cost_literal_kernel<M>,                                       \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
      if (e != cudaSuccess) return e;                                                                            \
    }                                                                                                            \
    rollout_cost_literal_kernel<M><<<grid, 128, smem, s>>>(d.hdr, d.eps, d.nominal, d.window, d.state, d.cost,   \
                                                           d.cmin, nearest, d.K, d.Kp, d.planes, d.win_stride,   \
                                                           d.T);                                                 \
  } while (0)
  switch (d.model) {
    case kDiffDrive: MPPI_LAUNCH_LITERAL(kDiffDrive); break;
    case kSteering: MPPI_LAUNCH_LITERAL(kSteering); break;
    default: MPPI_LAUNCH_LITERAL(kFullBody); break;
  }
#undef MPPI_LAUNCH_LITERAL
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// K3: weights
// ---------------------------------------------------------------------------------------------------------
// grid = (nb3, robots); 4 samples per thread.  Partial (sum w, sum w^2) per block -> wpart, summed in block
// order by the finalize kernel (deterministic).
__global__ void __launch_bounds__(kWeightBlock)
    weights_kernel(const SolveHeader *__restrict__ hdr, const float *__restrict__ cost,
                   const unsigned int *__restrict__ cmin, float *__restrict__ weight, float *__restrict__ wpart, int K,
                   int nb3) {
  const int robot = blockIdx.y;
  const float inv_lambda = hdr->inv_lambda;
  const float c_min = ordered_to_float(cmin[robot]);
  const float *c = cost + (size_t)robot * K;
  float *w = weight + (size_t)robot * K;
  float sw = 0.f, sw2 = 0.f;
  const int base = (blockIdx.x * kWeightBlock + threadIdx.x) * 4;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int i = base + k;
    if (i < K) {
      float wi = expf(-(c[i] - c_min) * inv_lambda);
      w[i] = wi;
      sw += wi;
      sw2 = fmaf(wi, wi, sw2);
    }
  }
  __shared__ float s_a[kWeightBlock / 32], s_b[kWeightBlock / 32];
  sw = warp_sum(sw);
  sw2 = warp_sum(sw2);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    s_a[wid] = sw;
    s_b[wid] = sw2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < kWeightBlock / 32; ++k) {
      a += s_a[k];
      b += s_b[k];
    }
    wpart[((size_t)robot * nb3 + blockIdx.x) * 2] = a;
    wpart[((size_t)robot * nb3 + blockIdx.x) * 2 + 1] = b;
  }
}

cudaError_t launch_weights(const DeviceState &d, cudaStream_t s) {
  dim3 grid(d.nb3, d.R);
  weights_kernel<<<grid, kWeightBlock, 0, s>>>(d.hdr, d.cost, d.cmin, d.weight, d.wpart, d.K, d.nb3);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// K4: weighted control reduction (a GEMV with M = 1: HBM-bound, CUDA cores)
// ---------------------------------------------------------------------------------------------------------
// grid = (nchunk, planes, robots).  One block reduces kReduceChunk samples of one (t,u) plane:
//   npart[robot][plane][chunk] = sum_i w_i * clamp(u*[plane] + sigma * eps[plane][i])
__global__ void __launch_bounds__(kReduceBlock)
    weighted_controls_kernel(const SolveHeader *__restrict__ hdr, const float *__restrict__ eps,
                             const float *__restrict__ weight, const float *__restrict__ nominal,
                             float *__restrict__ npart, int K, int Kp, int planes, int nchunk) {
  const int chunk = blockIdx.x, plane = blockIdx.y, robot = blockIdx.z;
  const SolveParams &P = hdr->P;
  const int u = plane % P.U;
  const float mean = nominal[(size_t)robot * planes + plane];
  const float lo = P.u_min[u], hi = P.u_max[u], sigma = P.sigma;
  const bool zeroed = (P.model == kFullBody) && P.steer_off && u == 2;  // FB:517
  const float4 *e4 = reinterpret_cast<const float4 *>(eps + ((size_t)robot * planes + plane) * Kp);
  const float *w = weight + (size_t)robot * K;
  const bool w_vec = (K % 4) == 0;
  float acc = 0.f;
  const int q0 = chunk * (kReduceChunk / 4);
  const int q1 = min(q0 + kReduceChunk / 4, Kp / 4);
  for (int q = q0 + threadIdx.x; q < q1; q += kReduceBlock) {
    float4 e = __ldcs(e4 + q);
    float wv[4];
    const int i = q * 4;
    if (w_vec) {
      float4 t = __ldg(reinterpret_cast<const float4 *>(w) + q);
      wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) wv[k] = (i + k < K) ? __ldg(w + i + k) : 0.f;
    }
    float ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float uk = zeroed ? 0.f : sample_control(ev[k], sigma, mean, lo, hi);
      if (i + k < K) acc = fmaf(wv[k], uk, acc);
    }
  }
  __shared__ float s_a[kReduceBlock / 32];
  acc = warp_sum(acc);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) s_a[wid] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < kReduceBlock / 32; ++k) a += s_a[k];
    npart[((size_t)robot * planes + plane) * nchunk + chunk] = a;
  }
}

cudaError_t launch_weighted_controls(const DeviceState &d, cudaStream_t s) {
  dim3 grid(d.nchunk, d.planes, d.R);
  weighted_controls_kernel<<<grid, kReduceBlock, 0, s>>>(d.hdr, d.eps, d.weight, d.nominal, d.npart, d.K, d.Kp,
                                                          d.planes, d.nchunk);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// K5: finalize -- fixed-order sums of the partials into the per-robot record {c_min, S, Q, -, N[P]}
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    finalize_kernel(const unsigned int *__restrict__ cmin, const float *__restrict__ wpart,
                    const float *__restrict__ npart, float *__restrict__ record, int planes, int nb3, int nchunk,
                    int rec_stride) {
  const int robot = blockIdx.x;
  float *rec = record + (size_t)robot * rec_stride;
  if (threadIdx.x < 32) {
    // lanes stride over the block partials, then a fixed shuffle tree -> deterministic
    float a = 0.f, b = 0.f;
    for (int k = threadIdx.x; k < nb3; k += 32) {
      a += wpart[((size_t)robot * nb3 + k) * 2];
      b += wpart[((size_t)robot * nb3 + k) * 2 + 1];
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (threadIdx.x == 0) {
      rec[0] = ordered_to_float(cmin[robot]);
      rec[1] = a;
      rec[2] = b;
      rec[3] = 0.f;
    }
  }
  for (int p = threadIdx.x; p < planes; p += blockDim.x) {
    const float *np = npart + ((size_t)robot * planes + p) * nchunk;
    float a = 0.f;
    for (int k = 0; k < nchunk; ++k) a += np[k];
    rec[4 + p] = a;
  }
}

cudaError_t launch_finalize(const DeviceState &d, cudaStream_t s) {
  finalize_kernel<<<d.R, 256, 0, s>>>(d.cmin, d.wpart, d.npart, d.record, d.planes, d.nb3, d.nchunk, d.rec_stride);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// K6: merge of G rank records -> u_new (out), nominal (warm start of the next solve), stats, counter++
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    merge_kernel(const SolveHeader *__restrict__ hdr, const float *__restrict__ gathered, float *__restrict__ u_new,
                 float *__restrict__ nominal, float *__restrict__ stats, uint32_t *__restrict__ counter, int planes,
                 int rec_stride, int R, int n_ranks) {
  const int robot = blockIdx.x;
  const float inv_lambda = hdr->inv_lambda;
  const size_t rank_stride = (size_t)R * rec_stride;
  const float *recs = gathered + (size_t)robot * rec_stride;
  const float m = merge_min(recs, n_ranks, rank_stride);
  float S, Q;
  merge_sums(recs, n_ranks, rank_stride, m, inv_lambda, S, Q);
  for (int p = threadIdx.x; p < planes; p += blockDim.x) {
    float u = merge_numerator(recs, n_ranks, rank_stride, m, inv_lambda, p) / S;
    u_new[(size_t)robot * planes + p] = u;
    nominal[(size_t)robot * planes + p] = u;  // un-shifted warm start, as the reference (DD:89-90)
  }
  if (threadIdx.x == 0) {
    stats[robot * 4 + 0] = m;
    stats[robot * 4 + 1] = S;
    stats[robot * 4 + 2] = S * S / Q;
    stats[robot * 4 + 3] = 0.f;
    if (robot == 0) *counter = *counter + 1u;
  }
}

cudaError_t launch_merge(const DeviceState &d, cudaStream_t s) {
  merge_kernel<<<d.R, 256, 0, s>>>(d.hdr, d.gathered, d.u_new, d.nominal, d.stats, d.counter, d.planes, d.rec_stride,
                                   d.R, d.n_ranks);
  return cudaGetLastError();
}

}  // namespace mppi
