// mppi_rollout_pruned.cu -- K2, production variant: fused sampling-clamp + rollout + cost with an EXACT pruned
// nearest-window-point scan.  One thread per sample; nothing per-step leaves the chip.
//
// What it replaces: clamp (DD:98-99) + predict_States (DD:111-122) + calc_Cost / calc_MinDistance (DD:183-210)
// [+ the ZMP loop FB:468-486 and cost FB:404-424], reference = /root/reference/src/{diff_drive,steering_diff_drive,
// full_body}_mppi.cpp.  Per-sample arithmetic is the FP32 contract of mppi_math.h; the per-sample cost is
// bit-identical to rollout_cost_literal_kernel (and to the host twin) for every input -- the scan only SKIPS
// window points that provably cannot be the minimum.
//
// How.  All K samples of a robot query the same T window points.  K0 (candidate_grid_kernel, once per robot and
// solve) lays a uniform grid of square cells (side h) over the window's bounding box plus a margin and stores,
// per cell, the smallest contiguous index range [lo, lo+n) that contains every window point that can be the
// nearest one for ANY position inside the cell:
//     with c the cell centre, r its half diagonal (inflated) and D_j = |c - r_j|:  for p in the cell
//     | |p - r_j| - D_j | <= r, so a point with D_j > min_k D_k + 2r is strictly farther from p than the point
//     that attains the minimum at c.  Candidates = { j : D_j <= D_min + 2r } (+ 1e-4 relative slack, four orders
//     of magnitude above the FP32 rounding of the distances involved), thinned by a corner test against the point
//     k0 nearest to c: |p - r_j|^2 - |p - r_k0|^2 is affine in p, positive at the four cell corners => r_j is
//     strictly farther than r_k0 everywhere in the cell (this keeps far-from-path cells at 3-5 candidates).
// K2 looks the cell of each predicted state up (one packed FMA + float->int per axis, one 32-bit load) and evaluates
// exact squared distances only for that range -- about 4-8 points instead of T, independent of T -- with
// Blackwell's packed FP32 pipe (sub/mul/fma .f32x2 -> FADD2/FMUL2/FFMA2, two window points per instruction,
// the same IEEE roundings per element as mppi::dist2) and the 3-input FMNMX3.  States outside the grid scan the
// whole window.  min() is exact and order independent, so the accumulated path cost has the literal scan's bits.
//
// Kernels in this file:
//   candidate_grid_kernel<LANES>     K0
//   rollout_cost_tma_kernel<MODEL,TAP> K2, production: the normals arrive through a per-warp TMA ring; optionally the
//                                    CTA also reduces its weighted controls (cta_weighted_controls).  TAP = the
//                                    instantiation that also records the argmin index (MPPI_DEBUG_NEAREST)
//   rescale_tail_kernel              K3' (per-CTA records -> partial sums at the robot's global minimum) + K5 + K6
//                                    (+ the NVLink record exchange) in one launch
#include <cuda.h>

#include <type_traits>

#include "mppi_device.cuh"

namespace mppi {

namespace {

typedef unsigned long long u64;

constexpr float kCellInflate = 1.002f;   // on the half diagonal: absorbs the rounding of the cell lookup
constexpr float kCandSlack = 1.0001f;    // relative slack on the candidate threshold (squared domain)

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
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float min3f(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// shared-memory accesses by 32-bit shared-window address: no generic->shared conversion in the inner loop
__device__ __forceinline__ float4 lds128(unsigned addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32(unsigned addr) {
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32_volatile(unsigned addr) {  // ring slots are rewritten by cp.async
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// Cell entry (K0 -> K2): low half = byte offset of the first window pair of the candidate range inside the shared
// pair array, high half = byte length of the range, a non-zero multiple of 32 (two pairs = four window points per
// scan iteration).  T <= 4096 keeps both below 2^16.
__host__ __device__ __forceinline__ uint32_t cell_entry(int first_pair, int iterations) {
  return (uint32_t)first_pair * 16u | ((uint32_t)iterations * 32u) << 16;
}

// min over the window pairs of one cell entry of min(d2, 1e4); pairs = {x0, x1, y0, y1}.  Two pairs (four window
// points) per iteration, at least one iteration; scanning a few points more than the candidate range is always
// safe (they are window points too), so the range is rounded up instead of predicated.
// (Round 2 tried to hand the rare long ranges -- ~0.7 % of the states, but they set the iteration count of ~18 % of
// the warp-steps -- to the whole warp: owner's position and range broadcast by shuffles, one pair per lane, shuffle-tree
// minimum.  Bit-identical, but a loss wherever several lanes of a warp are long at once (coarse grids, states outside
// the grid): K2 20 -> 32 us at K = 4096 / T = 50, 240 -> 505 us for 1024 robots, +-2 % at K = 2^17 .. 2^20.  Dropped;
// numbers in DESIGN.md section 4.)
__device__ __forceinline__ float pairs_min(float best, unsigned p, u64 xx, u64 yy) {  // two pairs at address p
  const float4 a = lds128(p), b = lds128(p + 16u);
  float a0, a1, b0, b1;
  unpack2(dist2_pair(xx, yy, pack2(a.x, a.y), pack2(a.z, a.w)), a0, a1);
  unpack2(dist2_pair(xx, yy, pack2(b.x, b.y), pack2(b.z, b.w)), b0, b1);
  return min3f(min3f(best, a0, a1), b0, b1);
}
// first iteration of the scan: straight-line code (every range has at least one), so that the first iterations of
// several states can be interleaved by the scheduler
__device__ __forceinline__ float scan_first(unsigned pairs_s, uint32_t e, u64 xy) {
  float x, y;
  unpack2(xy, x, y);
  return pairs_min(kDist2Cap, pairs_s + (e & 0xFFFFu), pack2(x, x), pack2(y, y));
}
// the rest of the range (~2 % of the states have one): the divergent part
__device__ __forceinline__ float scan_rest(float best, unsigned pairs_s, uint32_t e, u64 xy) {
  unsigned p = pairs_s + (e & 0xFFFFu) + 32u;
  const unsigned p_end = p + (e >> 16) - 32u;
  if (p != p_end) {
    float x, y;
    unpack2(xy, x, y);
    const u64 xx = pack2(x, x), yy = pack2(y, y);
#pragma unroll 1
    do {
      best = pairs_min(best, p, xx, yy);
      p += 32u;
    } while (p != p_end);
  }
  return best;
}

// The same scan, also returning the FIRST index that attains the minimum (-1 when no point is closer than the cap):
// what the literal loop `if (d2 < best) { best = d2; arg = j; }` (DD:186-190) records.  Every point that can be
// nearest -- ties included -- lies inside the candidate range (points outside are strictly farther), the range is
// scanned in index order, and the padding pairs repeat point T-1 AFTER it, so a strict `<` finds the same index.
__device__ __forceinline__ float scan_pairs_arg(unsigned pairs_s, uint32_t e, float x, float y, int *arg) {
  const u64 xx = pack2(x, x), yy = pack2(y, y);
  unsigned p = pairs_s + (e & 0xFFFFu);
  const unsigned p_end = p + (e >> 16);
  float best = kDist2Cap;
  int bi = -1;
#pragma unroll 1
  do {
    const float4 a = lds128(p), b = lds128(p + 16u);
    float d[4];
    unpack2(dist2_pair(xx, yy, pack2(a.x, a.y), pack2(a.z, a.w)), d[0], d[1]);
    unpack2(dist2_pair(xx, yy, pack2(b.x, b.y), pack2(b.z, b.w)), d[2], d[3]);
    const int j0 = (int)((p - pairs_s) >> 3);  // 16 bytes per pair of points
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (d[k] < best) {
        best = d[k];
        bi = j0 + k;
      }
    p += 32u;
  } while (p != p_end);
  *arg = bi;
  return best;
}

struct GridView {
  const uint32_t *cells;  // [ny + 2][nx + 2] cell entries; the one-cell border ring says "scan the whole window"
                          // (positions outside the grid are clamped to it)
  float inv_h, cx, cy;    // table index = floor(fma(x, inv_h, cx)), floor(fma(y, inv_h, cy)) (border included)
  int nx2, ixmax, iymax;  // nx + 2, nx + 1, ny + 1
};

// clamp(v, 0, hi) in one instruction: max(min(v, hi), 0)
__device__ __forceinline__ int clamp0(int v, int hi) {
  int r;
  asm("min.relu.s32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(hi));
  return r;
}

// Candidate range of the cell that contains the packed position {x, y}.  Branch-free lookup: one FFMA2 for both axes
// (the same fmaf per element), the float->int conversion saturates, the clamp sends everything outside the grid to the
// border ring.
__device__ __forceinline__ uint32_t grid_cell(const GridView &g, u64 xy) {
  float fx, fy;
  unpack2(fma2(xy, pack2(g.inv_h, g.inv_h), pack2(g.cx, g.cy)), fx, fy);
  const int ix = clamp0(__float2int_rd(fx), g.ixmax);
  const int iy = clamp0(__float2int_rd(fy), g.iymax);
  return __ldg(g.cells + (iy * g.nx2 + ix));
}
}  // namespace

// ---------------------------------------------------------------------------------------------------------
// K0: candidate grid (once per robot and solve)
// ---------------------------------------------------------------------------------------------------------
// grid = (cell blocks, robots), 128 threads, LANES adjacent lanes per cell (each takes every LANES-th window point;
// the partial results are combined by shuffles).  LANES > 1 shortens the serial chain per cell, which is what the
// latency configurations wait for; many-robot handles, whose grids together are large, use one lane per cell (least
// total work).  Every CTA re-derives the grid geometry from the window (same arithmetic, same result); CTA 0
// publishes it in ghdr[robot] for K2.
// The grid covers the window's bounding box plus a margin for the lateral spread of the rollouts: a quarter of
// the distance v_ref * dt * (T-1) a sample travels over the horizon (3 m at T = 100, 1.5 m at T = 50 with the launch
// parameters: measured optima), or margin_abs when that is positive.  Positions outside scan the whole window.
template <int LANES>
__global__ void __launch_bounds__(128)
    candidate_grid_kernel(const SolveHeader *__restrict__ hdr, const float *__restrict__ window,
                          GridHeader *__restrict__ ghdr, uint32_t *__restrict__ cells, int T, int win_stride,
                          int max_cells, float h_min, float margin_abs) {
  constexpr int kCellLanes = LANES;
  constexpr int kCellsPerBlock = 128 / LANES;
  extern __shared__ __align__(16) float2 s_win[];
  __shared__ GridHeader s_h;
  pdl_trigger();  // K2 may start its prologue (inputs -> shared memory) now; it waits for this grid before the lookup
  const int robot = blockIdx.y;
  float margin = margin_abs;
  if (!(margin > 0.f)) margin = fminf(fmaxf(0.25f * fabsf(hdr->P.v_ref) * hdr->P.dt * (float)(T - 1), 0.5f), 6.0f);
  const float *g_win = window + (size_t)robot * win_stride;
  for (int j = threadIdx.x; j < T; j += blockDim.x) s_win[j] = make_float2(g_win[2 * j], g_win[2 * j + 1]);
  __syncthreads();
  if (threadIdx.x < 32) {
    float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
    bool finite = true;
    for (int j = threadIdx.x; j < T; j += 32) {
      const float2 p = s_win[j];
      finite = finite && (fabsf(p.x) < 1.0e15f) && (fabsf(p.y) < 1.0e15f);
      xmin = fminf(xmin, p.x);
      xmax = fmaxf(xmax, p.x);
      ymin = fminf(ymin, p.y);
      ymax = fmaxf(ymax, p.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
      xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
      ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
      ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    finite = __all_sync(0xffffffffu, finite);
    if (threadIdx.x == 0) {
      GridHeader gh;
      gh.nx = gh.ny = 0;
      gh.inv_h = gh.cx = gh.cy = 0.f;
      gh.h = 0.f;
      gh.x0 = gh.y0 = 0.f;
      if (finite) {
        const float w = (xmax - xmin) + 2.f * margin, hgt = (ymax - ymin) + 2.f * margin;
        float h = fmaxf(h_min, sqrtf(w * hgt / (float)max_cells));
        int nx = 0, ny = 0;
        for (int it = 0; it < 64; ++it) {
          nx = (int)ceilf(w / h);
          ny = (int)ceilf(hgt / h);
          if ((long long)(nx + 2) * (ny + 2) <= max_cells && nx < 32768 && ny < 32768) break;
          h *= 1.05f;
        }
        if ((long long)(nx + 2) * (ny + 2) <= max_cells) {
          gh.nx = nx;
          gh.ny = ny;
          gh.h = h;
          gh.inv_h = 1.0f / h;
          gh.x0 = xmin - margin;
          gh.y0 = ymin - margin;
          gh.cx = 1.0f - gh.x0 * gh.inv_h;  // + 1: the border ring occupies table column / row 0
          gh.cy = 1.0f - gh.y0 * gh.inv_h;
        }
      }
      s_h = gh;
      if (blockIdx.x == 0) ghdr[robot] = gh;
    }
  }
  __syncthreads();
  const GridHeader gh = s_h;
  const int nx2 = gh.nx + 2, ny2 = gh.ny + 2;  // a failed geometry (nx = ny = 0) leaves a 2 x 2 table of border cells
  const int sub = threadIdx.x % kCellLanes;
  const int cell = blockIdx.x * kCellsPerBlock + threadIdx.x / kCellLanes;
  // the kCellLanes lanes of a cell take the same branches below (they are in one warp: shuffles stay converged)
  if (cell >= nx2 * ny2) return;
  const int tx = cell % nx2, ty = cell / nx2;
  if (tx == 0 || ty == 0 || tx == nx2 - 1 || ty == ny2 - 1) {  // border ring: the whole (padded) window
    if (sub == 0) cells[(size_t)robot * max_cells + cell] = cell_entry(0, ((T + 1) / 2 + 1) / 2);
    return;
  }
  const unsigned quad = (kCellLanes >= 32 ? 0xffffffffu : ((1u << (kCellLanes & 31)) - 1u))
                        << ((threadIdx.x & 31) / kCellLanes * kCellLanes);
  const int ix = tx - 1, iy = ty - 1;
  const float ccx = gh.x0 + ((float)ix + 0.5f) * gh.h, ccy = gh.y0 + ((float)iy + 0.5f) * gh.h;
  const float r = 0.70710678f * gh.h * kCellInflate + 1.0e-6f * (fabsf(ccx) + fabsf(ccy));
  // nearest window point to the cell centre: first minimum, as the serial scan (ties -> lowest index)
  float m = INFINITY;
  int k0 = 0x7FFFFFFF;
  for (int j = sub; j < T; j += kCellLanes) {
    const float dj = dist2(ccx, ccy, s_win[j].x, s_win[j].y);
    if (dj < m) {
      m = dj;
      k0 = j;
    }
  }
#pragma unroll
  for (int o = 1; o < kCellLanes; o <<= 1) {
    const float om = __shfl_xor_sync(quad, m, o);
    const int ok = __shfl_xor_sync(quad, k0, o);
    if (om < m || (om == m && ok < k0)) {
      m = om;
      k0 = ok;
    }
  }
  if (k0 == 0x7FFFFFFF) k0 = 0;  // every distance NaN
  const float reach = sqrtf(m) + 2.f * r;
  const float thr = reach * reach * kCandSlack;
  // Second, sharper filter for the survivors of the circle test: |p - r_j|^2 - |p - r_k0|^2 is affine in p, so if
  // it is positive at the four corners of the (inflated) cell it is positive everywhere in it -- r_j is then
  // strictly farther than r_k0 from every position in the cell and can never be the nearest point.  The margin
  // (1e-5 of the largest squared distance involved) is an order above the FP32 rounding of K2's own distances.
  const float hs = 0.5f * gh.h * kCellInflate + 1.0e-6f * (fabsf(ccx) + fabsf(ccy));
  const float kx = s_win[k0].x, ky = s_win[k0].y;
  float dk[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) dk[c] = dist2(ccx + ((c & 1) ? hs : -hs), ccy + ((c & 2) ? hs : -hs), kx, ky);
  int lo = T, hi = -1;
  for (int j = sub; j < T; j += kCellLanes) {
    const float jx = s_win[j].x, jy = s_win[j].y;
    if (dist2(ccx, ccy, jx, jy) <= thr) {
      float gap = INFINITY, scale = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float dj = dist2(ccx + ((c & 1) ? hs : -hs), ccy + ((c & 2) ? hs : -hs), jx, jy);
        gap = fminf(gap, dj - dk[c]);
        scale = fmaxf(scale, dj + dk[c]);
      }
      if (!(gap > 1.0e-5f * scale + 1.0e-30f)) {  // not provably farther than r_k0 everywhere (or NaN): candidate
        lo = min(lo, j);
        hi = max(hi, j);
      }
    }
  }
#pragma unroll
  for (int o = 1; o < kCellLanes; o <<= 1) {
    lo = min(lo, __shfl_xor_sync(quad, lo, o));
    hi = max(hi, __shfl_xor_sync(quad, hi, o));
  }
  // pairs of points, two pairs per scan iteration; an empty candidate set can only arise from NaNs: scan everything
  int q0 = 0, n2 = ((T + 1) / 2 + 1) / 2;
  if (hi >= lo) {
    q0 = lo >> 1;
    n2 = ((hi >> 1) - q0 + 2) >> 1;
  }
  if (sub == 0) cells[(size_t)robot * max_cells + cell] = cell_entry(q0, n2);
}

cudaError_t launch_candidate_grid(const DeviceState &d, cudaStream_t s) {
  const size_t smem = sizeof(float2) * (size_t)d.T;
  // lanes per cell: by the total number of cells of the handle (see the kernel's header comment)
  const long long total = (long long)d.R * d.grid_max_cells;
  int lanes = total <= 8192 ? 8 : (total <= 131072 ? 4 : 1);
  if (d.grid_lanes > 0) lanes = d.grid_lanes;
#define MPPI_LAUNCH_K0(L)                                                                                        \
  do {                                                                                                           \
    cudaError_t ce = set_carveout((const void *)candidate_grid_kernel<L>, d.side_carveout);                      \
    if (ce != cudaSuccess) return ce;                                                                            \
    if (smem > 48 * 1024) {                                                                                      \
      cudaError_t e = cudaFuncSetAttribute(candidate_grid_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           (int)smem);                                                           \
      if (e != cudaSuccess) return e;                                                                            \
    }                                                                                                            \
    dim3 grid((d.grid_max_cells + 128 / L - 1) / (128 / L), d.R);                                                \
    candidate_grid_kernel<L><<<grid, 128, smem, s>>>(d.hdr, d.window, d.grid_hdr, d.grid_cells, d.T, d.win_stride, \
                                                     d.grid_max_cells, d.grid_h_min, d.grid_margin);             \
  } while (0)
  switch (lanes) {
    case 32: MPPI_LAUNCH_K0(32); break;
    case 16: MPPI_LAUNCH_K0(16); break;
    case 8: MPPI_LAUNCH_K0(8); break;
    case 4: MPPI_LAUNCH_K0(4); break;
    case 2: MPPI_LAUNCH_K0(2); break;
    default: MPPI_LAUNCH_K0(1); break;
  }
#undef MPPI_LAUNCH_K0
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// K2 (pruned)
// ---------------------------------------------------------------------------------------------------------
constexpr int kScanChunk = 4;  // steps whose nearest-point scans are issued together (= one TMA stage of controls)

// What one thread needs besides its sample index (all CTA-uniform)
struct ThreadCtx {
  const SolveParams *sP;    // shared copy of the per-solve constants
  const ControlBounds *kb;  // clamp bounds: kernel parameter (constant bank operands, no registers)
  GridView gv;
  unsigned pairs_s, nom_s;  // shared-window addresses: window pairs, warm start
  float state0[5];          // {0, 0, yaw, roll, pitch}: the robot's state in its own frame
  int *tap_row;             // TAP only: nearest[sample][0..T) of this thread's sample (nullptr past K)
  int T;
};

// Registers of one rollout: pose, carried cos/sin pairs, cost accumulators, the candidate range of the current state.
// TAP: the debug instantiation that also records the argmin index of every cost state (MPPI_DEBUG_NEAREST).
template <int MODEL, bool SMALL, bool TAP>
struct Rollout {
  static constexpr int U = MODEL == kDiffDrive ? 2 : (MODEL == kSteering ? 3 : 5);
  const SolveParams &sP;
  const ControlBounds &kb;
  const GridView &gv;
  const unsigned pairs_s, nom_s;
  u64 xy;  // position {x, y} as one f32x2 register pair
  Attitude<MODEL> att;  // cos / sin of yaw (roll, pitch) carried through the horizon (mppi_math.h)
  CostAcc acc;
  uint32_t cell;
  float sigma, dt, v_ref;
  bool steer_off;
  int *tap_row;
  int tap_t;
  // the states of the current chunk whose nearest-point scans are still to do: position and cell entry
  u64 pend_xy[kScanChunk];
  uint32_t pend_cell[kScanChunk];

  __device__ __forceinline__ Rollout(const ThreadCtx &cx)
      : sP(*cx.sP), kb(*cx.kb), gv(cx.gv), pairs_s(cx.pairs_s), nom_s(cx.nom_s) {
    xy = pack2(cx.state0[0], cx.state0[1]);
    att.init(cx.state0);
    sigma = sP.sigma;
    dt = sP.dt;
    v_ref = sP.v_ref;
    steer_off = MODEL == kFullBody && sP.steer_off;
    tap_row = cx.tap_row;
    tap_t = 0;
    cell = grid_cell(gv, xy);
  }
  // sampling (D5) of step t: normals at shared address src + u * row_bytes; past the last step the result is never
  // used (s_nom is padded)
  __device__ __forceinline__ void make(int t, unsigned src, unsigned row_bytes, float *dst) const {
    const unsigned nom = nom_s + (unsigned)(t * U) * 4u;
    // mean + sigma * eps (D5) two controls at a time (FFMA2: the same fmaf per element), then the clamp
    float raw[U];
#pragma unroll
    for (int u = 0; u + 1 < U; u += 2) {
      const u64 e = pack2(lds32_volatile(src + u * row_bytes), lds32_volatile(src + (u + 1) * row_bytes));
      const u64 m = pack2(lds32(nom + u * 4u), lds32(nom + (u + 1) * 4u));
      unpack2(fma2(e, pack2(sigma, sigma), m), raw[u], raw[u + 1]);
    }
    if (U & 1) raw[U - 1] = fmaf(lds32_volatile(src + (U - 1) * row_bytes), sigma, lds32(nom + (U - 1) * 4u));
#pragma unroll
    for (int u = 0; u < U; ++u) dst[u] = clamp_ref(raw[u], kb.lo[u], kb.hi[u]);
    if (steer_off) dst[2] = 0.f;  // FB:517
  }
  // One control step, dynamics first: the Euler step with the controls `cur`, the candidate-range load of the NEXT
  // state, and the cost terms that do not need the nearest point.  The state the step starts from is parked in slot
  // SLOT of the chunk; its nearest-point scan happens in scans() -- a step's scan does not feed the dynamics, so the
  // scans of kScanChunk consecutive steps are independent of each other and of the Euler steps in between, and are
  // issued back to back: their shared-memory loads and FP chains overlap instead of each waiting behind its own
  // data-dependent loop.  (Per step in sequence, a lone warp spent 80 % of its cycles stalled: profiles/r02_ncu_latency.txt.)
  template <int SLOT>
  __device__ __forceinline__ void step(const float *cur, const float *nxt) {
    pend_xy[SLOT] = xy;
    pend_cell[SLOT] = cell;
    const float sr0 = att.sr, cr0 = att.cr, sp0 = att.sp, cp0 = att.cp;
    float sd = 0.f, cd = 1.f, ch, sh;
    if (MODEL != kDiffDrive) sincos_f32<SMALL>(cur[2], sd, cd);
    // step_state (mppi_math.h) with the position update packed: {x, y} = fma({v*ch, v*sh}, dt, {x, y})
    step_heading<MODEL>(att, sd, cd, ch, sh);
    xy = fma2(mul2(pack2(cur[0], cur[0]), pack2(ch, sh)), pack2(dt, dt), xy);
    step_attitude<MODEL, SMALL>(att, cur, dt);
    cell = grid_cell(gv, xy);
    const float dv = cur[0] - v_ref;
    acc.vel = fmaf(dv, dv, acc.vel);
    if (MODEL == kFullBody) {
      float zx, zy;
      zmp_model(sP, cur[0], nxt[0], cur[1], sd, cd, cur[3], nxt[3], cur[4], nxt[4], sr0, cr0, sp0, cp0, zx, zy);
      acc.zmp = fmaf(zy, zy, acc.zmp);
      const float dr = nxt[3] - cur[3];
      acc.droll = fmaf(dr, dr, acc.droll);
      if (cur[0] < 0.f) acc.back = fmaf(cur[0], cur[0], acc.back);
    }
  }
  // min_j min(d2, 1e4) of the N parked states over their cells' candidate ranges, added to the path cost in step
  // order (the accumulation order of the contract); the debug instantiation records the argmin of each.
  template <int N>
  __device__ __forceinline__ void scans() {
    float best[N];
    if (TAP) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        float x, y;
        unpack2(pend_xy[k], x, y);
        int arg;
        best[k] = scan_pairs_arg(pairs_s, pend_cell[k], x, y, &arg);
        if (tap_row) tap_row[tap_t] = arg;
        ++tap_t;
      }
    } else {
#pragma unroll
      for (int k = 0; k < N; ++k) best[k] = scan_first(pairs_s, pend_cell[k], pend_xy[k]);
#pragma unroll
      for (int k = 0; k < N; ++k) best[k] = scan_rest(best[k], pairs_s, pend_cell[k], pend_xy[k]);
    }
#pragma unroll
    for (int k = 0; k < N; ++k) acc.path += best[k];
  }
  // one step and its scan right away (the ragged end of the horizon)
  __device__ __forceinline__ void advance(const float *cur, const float *nxt) {
    step<0>(cur, nxt);
    scans<1>();
  }
  __device__ __forceinline__ float finish(const float *st) {
    if (MODEL != kFullBody) {  // state T-1: path term only (D1)
      pend_xy[0] = xy;
      pend_cell[0] = cell;
      scans<1>();
    }
    float yaw0_err = 0.f;
    if (MODEL == kFullBody) {  // FB:408; window points 0 and 1 are the first pair {x0, x1, y0, y1}
      const float4 w01 = lds128(pairs_s);
      yaw0_err = st[2] - yaw_ref0_f32(w01.x, w01.z, w01.y, w01.w);
    }
    return combine_cost(sP, acc, yaw0_err);
  }
};

// ---- TMA ring -------------------------------------------------------------------------------------------------
// One ring PER WARP: a stage is the tile {32 samples of the warp} x {kStageSteps control steps x U planes} of the
// plane-major noise tensor, fetched by ONE cp.async.bulk.tensor issued by an elected lane and landing as [row][32]
// floats (128 B rows, conflict-free column reads).  Completion is signalled on the stage's mbarrier (expect-tx
// bytes); a slot is refilled after __syncwarp(), once every lane has consumed its column in an Euler step.  Warps
// never wait for each other.  Per control step this costs ~2 issue slots (a per-thread LDGSTS ring -- address
// arithmetic, predicates, commit / wait groups -- cost ~14 and was dropped).
constexpr int kStageSteps = 4;
constexpr int kStages = 3;
static_assert(kStageSteps == kScanChunk, "the rollout loop parks one TMA stage of steps per scan chunk");

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap *map, int c0, int c1, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

struct TmaCtx {
  const CUtensorMap *map;
  unsigned ring_s;  // this warp's ring: kStages tiles of kStageSteps * U rows x 128 B
  unsigned bar_s;   // this warp's kStages mbarriers
  int c0;           // first sample of the warp
  int row0;         // first row of this robot's planes in the (double-buffered) tensor
};

template <int MODEL, bool SMALL, bool TAP>
__device__ __forceinline__ float rollout_thread_tma(const ThreadCtx &cx, const TmaCtx &tc) {
  constexpr int U = MODEL == kDiffDrive ? 2 : (MODEL == kSteering ? 3 : 5);
  constexpr int kRows = kStageSteps * U;
  constexpr unsigned kTileBytes = kRows * 128u;
  constexpr unsigned kStepBytes = U * 128u;
  Rollout<MODEL, SMALL, TAP> r(cx);
  const int T = cx.T;
  const int n_iter = MODEL == kFullBody ? T - 2 : T - 1;
  // tiles 0 .. n_tiles-1 cover control steps 0 .. n_iter (the last one may lie past the end: sampled, never used;
  // rows beyond the tensor are zero-filled by the TMA unit), so every make() below has a tile to wait for -- and
  // every tile that is fetched is also waited for before the thread returns (no copy in flight at CTA exit)
  const int n_tiles = n_iter / kStageSteps + 1;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned col = tc.ring_s + lane * 4u;
  auto fetch = [&](int tile, unsigned slot) {
    if (tile < n_tiles) {  // warp-uniform
      if (elect_one()) {
        const unsigned bar = tc.bar_s + slot * 8u;
        mbar_expect_tx(bar, kTileBytes);
        tma_load_2d(tc.ring_s + slot * kTileBytes, tc.map, tc.c0, tc.row0 + tile * kRows, bar);
      }
    }
  };
#pragma unroll
  for (int k = 0; k < kStages; ++k) fetch(k, (unsigned)k);
  float ca[U], cb[U];
  unsigned slot = 0, parity = 0;
  mbar_wait(tc.bar_s, 0);
  r.make(0, col, 128u, ca);
  int t = 0;
  for (int tile = 0; t + kStageSteps <= n_iter; t += kStageSteps, ++tile) {
    const unsigned base = col + slot * kTileBytes;
    r.make(t + 1, base + kStepBytes, 128u, cb);
    r.template step<0>(ca, cb);
    r.make(t + 2, base + 2 * kStepBytes, 128u, ca);
    r.template step<1>(cb, ca);
    r.make(t + 3, base + 3 * kStepBytes, 128u, cb);
    r.template step<2>(ca, cb);
    const unsigned done = slot;
    if (++slot == kStages) {
      slot = 0;
      parity ^= 1u;
    }
    mbar_wait(tc.bar_s + slot * 8u, parity);
    r.make(t + 4, col + slot * kTileBytes, 128u, ca);
    r.template step<3>(cb, ca);
    r.template scans<kScanChunk>();
    // Refill the finished slot with the tile kStages ahead.  Every value the lanes loaded from it has been consumed
    // by an Euler step above (cb, its last row, by the advance just before this line), so no shared-memory read of
    // the slot is outstanding when the bulk copy is issued; __syncwarp() orders the lanes.
    __syncwarp();
    fetch(tile + kStages, done);
  }
  {  // at most kStageSteps - 1 iterations, all inside the current tile; ca = step t on entry
    const unsigned base = col + slot * kTileBytes;
    for (int k = 1; t < n_iter; ++t, ++k) {
      r.make(t + 1, base + k * kStepBytes, 128u, cb);
      r.advance(ca, cb);
#pragma unroll
      for (int u = 0; u < U; ++u) ca[u] = cb[u];
    }
  }
  return r.finish(cx.state0);
}

// ---- fused weighted-control partials of one CTA's 128 samples (replaces K3 + K4) -------------------------------
// calc_Weights (DD:216-222) + determine_OptimalSolution (DD:228-236) against the CTA's own minimum m_cta:
//   w_i = exp(-(c_i - m_cta)/lambda),  S = sum w,  Q = sum w^2,  N[p] = sum_i w_i * clamp(u*_p + sigma*eps[p][i])
// written as the record {m_cta, S, Q, -, N[planes]}; the rescale brings the records of all CTAs to the robot's
// global minimum (the same log-sum-exp merge as between GPUs, section 6 of DESIGN.md).  The CTA re-reads its
// 128 x planes tile of the normals right after streaming it through the TMA ring -- mostly from L2, and in the
// shadow of the other CTAs' rollouts (K2 uses a quarter of the HBM bandwidth) -- instead of a separate pass of
// the whole tensor through HBM.  Not inlined (called once per thread, keeps the rollout loop's code compact).
template <int MODEL>
__device__ __noinline__ void cta_weighted_controls(const SolveParams &sP, float4 bounds01, float inv_lambda, float c,
                                                   bool valid, float m_cta, const float *s_nom,
                                                   const float *eps_robot, int Kp, int planes, float *rec) {
  constexpr int U = MODEL == kDiffDrive ? 2 : (MODEL == kSteering ? 3 : 5);
  __shared__ __align__(16) float s_w[128];
  __shared__ float s_sq[2][4];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  {
    const float w = valid ? expf(-(c - m_cta) * inv_lambda) : 0.f;
    s_w[threadIdx.x] = w;
    const float sw = warp_sum(w), sq = warp_sum(w * w);
    if (lane == 0) {
      s_sq[0][wid] = sw;
      s_sq[1][wid] = sq;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    rec[0] = m_cta;
    rec[1] = (s_sq[0][0] + s_sq[0][1]) + (s_sq[0][2] + s_sq[0][3]);
    rec[2] = (s_sq[1][0] + s_sq[1][1]) + (s_sq[1][2] + s_sq[1][3]);
    rec[3] = 0.f;
  }
  const float4 wq = *reinterpret_cast<const float4 *>(&s_w[4 * lane]);  // weights of samples 4*lane .. 4*lane+3
  // first of this lane's four samples; lanes past the (padded) sample count read the last quad instead -- their four
  // weights are 0 (samples >= K), so what they load does not matter and the loads need no predicate (Kp is a multiple of 4)
  const int s0 = min(blockIdx.x * 128 + 4 * lane, Kp - 4);
  const float *e_base = eps_robot + s0;
  const float sigma = sP.sigma;
  const bool steer_off = MODEL == kFullBody && sP.steer_off;
  const float4 *s_nom4 = reinterpret_cast<const float4 *>(s_nom);  // 16-byte aligned, padded by 2 U >= 4 entries
  const u64 sig2 = pack2(sigma, sigma), w01 = pack2(wq.x, wq.y), w23 = pack2(wq.z, wq.w);
  // a warp takes 4 consecutive planes at a time: 4 independent 16-byte loads per lane, then one transposing butterfly
  // reduces the 4 partial sums over the 32 lanes (lanes 0, 8, 16, 24 end up with one plane each); fixed order,
  // deterministic.  (Issuing the next four planes' loads ahead of the arithmetic, or 8 planes at a time, costs 12-18
  // registers -- one resident CTA per SM -- and gained nothing at K = 2^17 .. 2^20: measured, dropped.)
  // Full groups of four planes walk a running pointer (no per-load index arithmetic); the one ragged group at the end
  // of the record (planes is not a multiple of 4 for every model / horizon) clamps its plane indices instead.
  const size_t plane_stride = (size_t)Kp;
  const bool up = lane & 16, up8 = lane & 8;
  const int idx = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);  // value index held by this lane
  auto group = [&](int p0, const float *ptr, auto full_tag) {
    constexpr bool kFull = decltype(full_tag)::value;
    float4 e[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float *q = kFull ? ptr + k * plane_stride : e_base + (size_t)min(p0 + k, planes - 1) * plane_stride;
      e[k] = __ldcs(reinterpret_cast<const float4 *>(q));
    }
    const float4 mean4 = s_nom4[p0 >> 2];  // warm start of the four planes (p0 is a multiple of 4): one LDS.128
    const float means[4] = {mean4.x, mean4.y, mean4.z, mean4.w};
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // two controls: p0 is a multiple of 4, so the control index is k & 1 at compile time and the clamp bounds sit in
      // registers (bounds01 = {lo0, hi0, lo1, hi1}); otherwise they come from the shared parameter block
      const int u = U == 2 ? (k & 1) : (p0 + k) % U;  // (a plane past the end is never stored)
      const float lo = U == 2 ? ((k & 1) ? bounds01.z : bounds01.x) : sP.u_min[u];
      const float hi = U == 2 ? ((k & 1) ? bounds01.w : bounds01.y) : sP.u_max[u];
      const float mean = means[k];  // planes past the end read the zero padding
      // sample_control of the four samples, two at a time (FFMA2: the same fmaf per element, then the scalar clamp),
      // and the weighted sum as two packed partial sums {w0 u0 + w2 u2, w1 u1 + w3 u3}
      float r0, r1, r2, r3;
      unpack2(fma2(pack2(e[k].x, e[k].y), sig2, pack2(mean, mean)), r0, r1);
      unpack2(fma2(pack2(e[k].z, e[k].w), sig2, pack2(mean, mean)), r2, r3);
      const u64 c01 = pack2(clamp_ref(r0, lo, hi), clamp_ref(r1, lo, hi));
      const u64 c23 = pack2(clamp_ref(r2, lo, hi), clamp_ref(r3, lo, hi));
      float a_lo, a_hi;
      unpack2(fma2(c23, w23, mul2(c01, w01)), a_lo, a_hi);
      const float a = a_lo + a_hi;
      v[k] = (steer_off && u == 2) ? 0.f : a;  // FB:517: every sample of that control is 0
    }
    // 4 values x 32 lanes -> lane l (l % 8 == 0) holds plane p0 + (l >> 3 with the two bits swapped: see idx)
    float r[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float send = up ? v[k] : v[k + 2];
      const float keep = up ? v[k + 2] : v[k];
      r[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    float r1 = (up8 ? r[1] : r[0]) + __shfl_xor_sync(0xffffffffu, up8 ? r[0] : r[1], 8);
    r1 += __shfl_xor_sync(0xffffffffu, r1, 4);
    r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
    r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
    if ((lane & 7) == 0 && (kFull || p0 + idx < planes)) rec[4 + p0 + idx] = r1;
  };
  int p0 = wid * 4;
  const float *ptr = e_base + (size_t)p0 * plane_stride;
  for (; p0 + 4 <= planes; p0 += 16, ptr += 16 * plane_stride) group(p0, ptr, std::true_type{});
  if (p0 < planes) group(p0, ptr, std::false_type{});  // warp-uniform: p0 depends on the warp index only
}

// K2 with the TMA ring.  Dynamic shared memory = ring [4 warps][kStages][rows][32] (1 KB aligned) | window pairs |
// warm start.  Whole warps run the rollout (the refill needs all 32 lanes at the __syncwarp); lanes past K compute
// on zero / padding normals and are masked at the store.
size_t pruned_tma_smem_bytes(int T, int planes, int U) {
  return (size_t)4 * kStages * kStageSteps * U * 128 + sizeof(float4) * (size_t)((T + 1) / 2 + 1) +
         sizeof(float) * ((size_t)planes + 2 * U);
}

// resident CTAs per SM the register allocation must allow: 8 (64 registers, no spills) for the two- and
// three-control models, 5 for the full-body model.  Measured with the chunked scans (K = 2^20 / 2^17 / 4096):
// 9 CTAs (56 registers, 12 bytes spilled) 0.519 / 0.098 / 0.0279 ms, 8 CTAs 0.518 / 0.095 / 0.0272, 7 CTAs
// (72 registers) 0.523 / 0.094 / 0.0272
#ifndef MPPI_K2_MINBLOCKS
#define MPPI_K2_MINBLOCKS 8
#endif
#ifndef MPPI_K2_MINBLOCKS_FB
#define MPPI_K2_MINBLOCKS_FB 5
#endif
template <int MODEL, bool TAP>
__global__ void __launch_bounds__(128, MODEL == kFullBody ? MPPI_K2_MINBLOCKS_FB : MPPI_K2_MINBLOCKS)
    rollout_cost_tma_kernel(const __grid_constant__ CUtensorMap eps_map, const __grid_constant__ ControlBounds kb,
                            const SolveHeader *__restrict__ hdr,
                            const float *__restrict__ nominal, const float *__restrict__ window,
                            const float *__restrict__ state, const GridHeader *__restrict__ ghdr,
                            const uint32_t *__restrict__ cells, float *__restrict__ cost,
                            unsigned int *__restrict__ cmin, int K, int planes, int win_stride, int T, int max_cells,
                            const float *__restrict__ eps, int Kp, float *__restrict__ cta_part, int part_stride,
                            const uint32_t *__restrict__ counter, uint32_t pmask, size_t buf_elems,
                            int *__restrict__ nearest) {
  constexpr int U = MODEL == kDiffDrive ? 2 : (MODEL == kSteering ? 3 : 5);
  constexpr int kWarpRingBytes = kStages * kStageSteps * U * 128;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ SolveParams sP;
  __shared__ float s_red[32];
  __shared__ GridHeader s_gh;
  __shared__ __align__(8) unsigned long long s_bar[4 * kStages];
  const int robot = blockIdx.y;
  const int NP = (T + 1) / 2;
  float4 *s_pairs = reinterpret_cast<float4 *>(smem_raw + 4 * kWarpRingBytes);
  float *s_nom = reinterpret_cast<float *>(s_pairs + NP + 1);

  const float *g_win = window + (size_t)robot * win_stride;
  pdl_trigger();  // the reduction kernel behind may be launched; it waits for this grid before it reads a cost
  // Inputs of the solve that were complete before the kernel in front of this one was launched (header, window): the
  // kernel in front is the candidate grid (graph) or the previous solve's tail (stream launches).
  load_params_to_shared(&sP, hdr);
  if (threadIdx.x < 4 * kStages) mbar_init((unsigned)__cvta_generic_to_shared(&s_bar[threadIdx.x]), 1u);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  for (int q = threadIdx.x; q < NP + 1; q += blockDim.x) {
    // odd T and the pad pair: the last point again (no effect on a minimum)
    const int j0 = min(2 * q, T - 1), j1 = min(2 * q + 1, T - 1);
    s_pairs[q] = make_float4(g_win[2 * j0], g_win[2 * j1], g_win[2 * j0 + 1], g_win[2 * j1 + 1]);
  }
  pdl_wait();  // the candidate grid (K0) / the previous solve's warm start, counter and minimum slot are complete
  for (int j = threadIdx.x; j < planes + 2 * U; j += blockDim.x)
    s_nom[j] = j < planes ? nominal[(size_t)robot * planes + j] : 0.f;
  if (threadIdx.x == 0) s_gh = ghdr[robot];
  __syncthreads();

  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  // warp index through a shuffle: the compiler then knows it is warp-uniform (TMA operands live in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int warp_first = blockIdx.x * blockDim.x + warp * 32;
  const uint32_t par = eps_parity(counter, pmask);  // which buffer of the noise tensor belongs to this solve
  float c = 0.f;
  if (warp_first < K) {  // warp-uniform
    const uint32_t *cp = cells + (size_t)robot * max_cells;
    asm volatile("" : "+l"(cp));  // keep the robot's table base in a register pair (no per-step recomputation)
    ThreadCtx cx;
    cx.sP = &sP;
    cx.kb = &kb;
    cx.gv = GridView{cp, s_gh.inv_h, s_gh.cx, s_gh.cy, s_gh.nx + 2, s_gh.nx + 1, s_gh.ny + 1};
    cx.pairs_s = (unsigned)__cvta_generic_to_shared(s_pairs);
    cx.nom_s = (unsigned)__cvta_generic_to_shared(s_nom);
    {
      const float *st = state + (size_t)robot * kStateStride;  // {yaw, roll, pitch, -}
      cx.state0[0] = cx.state0[1] = 0.f;
      cx.state0[2] = st[0];
      cx.state0[3] = st[1];
      cx.state0[4] = st[2];
    }
    cx.tap_row = (TAP && i < K) ? nearest + ((size_t)robot * K + i) * T : nullptr;
    cx.T = T;
    TmaCtx tc;
    tc.map = &eps_map;
    tc.ring_s = (unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)(warp * kWarpRingBytes);
    tc.bar_s = (unsigned)__cvta_generic_to_shared(&s_bar[warp * kStages]);
    tc.c0 = warp_first;
    tc.row0 = ((int)par * (int)gridDim.y + robot) * planes;
    asm volatile("" : "+r"(cx.pairs_s), "+r"(cx.nom_s), "+r"(tc.ring_s), "+r"(tc.bar_s));
    c = angles_are_small(sP) ? rollout_thread_tma<MODEL, true, TAP>(cx, tc) : rollout_thread_tma<MODEL, false, TAP>(cx, tc);
    if (i < K) cost[(size_t)robot * K + i] = c;
  }
  const float m_cta = block_min_to_global(c, i < K, cmin + robot, s_red);
  if (cta_part == nullptr) return;  // kernel argument: uniform
  cta_weighted_controls<MODEL>(sP, make_float4(kb.lo[0], kb.hi[0], kb.lo[1], kb.hi[1]), hdr->inv_lambda, c, i < K, m_cta, s_nom,
                               eps + (size_t)par * buf_elems + (size_t)robot * planes * Kp, Kp, planes,
                               cta_part + ((size_t)robot * gridDim.x + blockIdx.x) * part_stride);
}

// ---- the tail of the fused-controls path: K3' + K5 + K6 (+ the record exchange) in ONE launch ---------------------
// K2 left one record {m_c, S_c, Q_c, -, N_c[P]} per CTA, exponentiated against the CTA's own minimum m_c.  With
//   a_c = exp(-(m_c - c_min)/lambda):   S = sum_c a_c S_c,   Q = sum_c a_c^2 Q_c,   N[p] = sum_c a_c N_c[p]
// the robot's record {c_min, S, Q, -, N[P]} follows column by column: one thread per record column, the same loop for
// every column (weight a_c, or a_c^2 for the Q column), coalesced over the records' rows.  The tail is a chain of L2
// round trips, not bandwidth, so it is kept to two: every block folds its group of CTA records (all loads of a thread
// -- the CTA minima, broadcast, and its column -- are independent: one batch), the block that finishes last for a robot
// adds the groups' partial records (a second batch) and, MODE 0 (unsharded), merges: u = N / S -> u_new, warm start,
// stats.  The block that finishes last of the whole grid advances the solve counter and, MODE 1 (peer exchange),
// first runs exchange_and_merge for all robots.  MODE 2: records only (NCCL all-gather + merge kernel follow).
// Sums use kTailMlp interleaved accumulators combined in a fixed order: deterministic for a given launch geometry.
constexpr int kTailMlp = 32;
template <int B>
__device__ __forceinline__ float sum_fixed(const float (&a)[B]) {
  float s[B / 2];
#pragma unroll
  for (int k = 0; k < B / 2; ++k) s[k] = a[2 * k] + a[2 * k + 1];
#pragma unroll
  for (int w = B / 4; w > 0; w >>= 1)
#pragma unroll
    for (int k = 0; k < w; ++k) s[k] = s[2 * k] + s[2 * k + 1];
  return s[0];
}

// One batch of a record column: B loads in flight per thread (rows c .. c+B-1 of the block's group; the rows past the
// group repeat its last record and get the weight 0 below)
template <int B>
__device__ __forceinline__ void load_batch(float (&v)[B], const float *__restrict__ part, int part_stride, int nc,
                                           int c, int col) {
#pragma unroll
  for (int k = 0; k < B; ++k) v[k] = part[(size_t)min(c + k, nc - 1) * part_stride + col];
}
// acc[k] += a_c * v[k] with the records' rescale factors a_c from shared memory (a_c^2 for the Q column)
template <int B>
__device__ __forceinline__ void fma_batch(float (&acc)[B], const float (&v)[B], const float *s_a, int nc, int c,
                                          bool squared) {
#pragma unroll
  for (int k = 0; k < B; ++k) {
    float a = c + k < nc ? s_a[c + k] : 0.f;
    if (squared) a *= a;
    acc[k] = fmaf(a, v[k], acc[k]);
  }
}

// Every column of the rescaled sum over the nc CTA records of the block's group (col 1: S, col 2: Q, col >= 4:
// N[col - 4]; columns 0 and 3 are 0), one thread per column.  The factors a_c = exp(-(m_c - c_min)/lambda) are
// computed ONCE per record (thread c) into shared memory -- not once per (record, column) -- while the first batch of
// every thread's column is already in flight: the block's first L2 round trip carries both.  All threads of the block
// must call it (it contains a barrier).  Results: column threadIdx.x in out0, threadIdx.x + blockDim.x in out1,
// further ones only through gp (when non-null every column is stored there).
template <int B>
__device__ __forceinline__ void fold_group(const float *__restrict__ part, int part_stride, int nc, int ncol,
                                           const unsigned int *__restrict__ cmin_slot, float inv_lambda, float *s_a,
                                           float *__restrict__ gp, float &c_min_out, float &out0, float &out1) {
  const int col0 = threadIdx.x;
  const bool live0 = col0 < ncol && col0 != 0 && col0 != 3;
  float v[B];
  if (live0) load_batch<B>(v, part, part_stride, nc, 0, col0);
  const float c_min = ordered_to_float(*cmin_slot);
  for (int c = threadIdx.x; c < nc; c += blockDim.x) s_a[c] = expf(-(part[(size_t)c * part_stride] - c_min) * inv_lambda);
  __syncthreads();
  c_min_out = c_min;
  int slot = 0;
  for (int col = threadIdx.x; col < ncol; col += blockDim.x, ++slot) {
    float r = 0.f;
    if (col != 0 && col != 3) {
      float acc[B];
#pragma unroll
      for (int k = 0; k < B; ++k) acc[k] = 0.f;
      for (int c = 0; c < nc; c += B) {
        if (slot != 0 || c != 0) load_batch<B>(v, part, part_stride, nc, c, col);
        fma_batch<B>(acc, v, s_a, nc, c, col == 2);
      }
      r = sum_fixed<B>(acc);
    }
    if (gp) gp[col] = r;
    if (slot == 0) out0 = r;
    else if (slot == 1) out1 = r;
  }
}

template <int MODE>
__global__ void __launch_bounds__(256)
    rescale_tail_kernel(const SolveHeader *__restrict__ hdr, const float *__restrict__ cta_part,
                        unsigned int *__restrict__ cmin, float *__restrict__ gpart, float *__restrict__ record,
                        float *__restrict__ u_new, float *__restrict__ nominal, float *__restrict__ stats,
                        uint32_t *__restrict__ counter, unsigned int *__restrict__ ticket, int planes, int n_cta,
                        int part_stride, int groups, int per, int R, ExchangeArgs x) {
  __shared__ float s_S;
  __shared__ float s_a[kRescaleMaxCtas];  // rescale factors of the block's CTA records
  const int robot = blockIdx.y, g = blockIdx.x;
  const int ncol = 4 + planes;
  const float inv_lambda = hdr->inv_lambda;  // an input of the solve
  const int c0 = g * per, nc = min(per, n_cta - c0);
  const float *part = cta_part + ((size_t)robot * n_cta + c0) * part_stride;
  float *rec = record + (size_t)robot * part_stride;
  // many-robot handles: one block owns the robot and (record columns <= 2 per thread) nothing leaves its registers
  const bool in_regs = ncol <= 2 * (int)blockDim.x;
  const bool single_regs = groups == 1 && in_regs;
  float *gp = single_regs ? nullptr : gpart + ((size_t)robot * groups + g) * part_stride;
  float mine0 = 0.f, mine1 = 0.f;  // this thread's (up to two) columns
  float c_min;
  int slot = 0;
  pdl_wait();     // the CTA records and the minimum come from K2, the kernel in front
  pdl_trigger();  // the next solve's K2 may bring its CTAs onto the SMs (header, window -> shared memory) meanwhile
  if (nc <= 8) fold_group<8>(part, part_stride, nc, ncol, cmin + robot, inv_lambda, s_a, gp, c_min, mine0, mine1);
  else fold_group<kTailMlp>(part, part_stride, nc, ncol, cmin + robot, inv_lambda, s_a, gp, c_min, mine0, mine1);
  if (!single_regs) {
    if (!last_block_of_grid(ticket + 1 + robot, gridDim.x)) return;
  } else {
    __syncthreads();  // every thread has read cmin before it is re-armed below
  }
  // this block finishes the robot: add the groups' partial records, column by column
  slot = 0;
  for (int col = threadIdx.x; col < ncol; col += blockDim.x, ++slot) {
    float tot;
    if (single_regs) {
      tot = slot == 0 ? mine0 : mine1;
    } else {
      const float *q = gpart + (size_t)robot * groups * part_stride + col;
      float acc[kTailMlp];
#pragma unroll
      for (int k = 0; k < kTailMlp; ++k) acc[k] = 0.f;
      for (int gg = 0; gg < groups; gg += kTailMlp) {
#pragma unroll
        for (int k = 0; k < kTailMlp; ++k)
          acc[k] += gg + k < groups ? __ldcg(q + (size_t)(gg + k) * part_stride) : 0.f;  // written by other blocks
      }
      tot = sum_fixed<kTailMlp>(acc);
    }
    if (col == 0) tot = c_min;
    rec[col] = tot;
    if (col == 1) s_S = fmaf(1.f, tot, 0.f);  // merge_sums with one rank
    if (slot == 0) mine0 = tot;
    else if (slot == 1) mine1 = tot;
  }
  __syncthreads();
  if (threadIdx.x == 0) cmin[robot] = 0xFFFFFFFFu;  // every reader of this solve is done: armed for the next atomicMin
  if (MODE == 0) {
    const float S = s_S;
    slot = 0;
    for (int col = threadIdx.x; col < ncol; col += blockDim.x, ++slot) {
      const float tot = in_regs ? (slot == 0 ? mine0 : mine1) : rec[col];
      if (col >= 4) {
        const float u = fmaf(1.f, tot, 0.f) / S;  // merge_numerator with one rank
        u_new[(size_t)robot * planes + col - 4] = u;
        if (nominal) nominal[(size_t)robot * planes + col - 4] = u;  // un-shifted warm start, as the reference (DD:89-90)
      } else if (col == 2) {
        const float Q = fmaf(1.f, tot, 0.f);
        stats[robot * 4 + 0] = c_min;
        stats[robot * 4 + 1] = S;
        stats[robot * 4 + 2] = S * S / Q;
        stats[robot * 4 + 3] = 0.f;
      }
    }
    // nothing reads the counter before the next solve's kernels: any one block may advance it
    if (robot == 0 && threadIdx.x == 0) *counter = *counter + 1u;
    return;
  }
  if (MODE == 1) {
    if (R > 1) {
      if (!last_block_of_grid(ticket, (unsigned)R)) return;
      exchange_and_merge(hdr, record, x, u_new, nominal, stats, counter, planes, part_stride, R);
    } else if (in_regs) {  // one robot: this block holds the whole record in registers -- push it from there
      exchange_and_merge(hdr, record, x, u_new, nominal, stats, counter, planes, part_stride, R, true,
                         threadIdx.x < ncol ? mine0 : 0.f, (int)(threadIdx.x + blockDim.x) < ncol ? mine1 : 0.f);
    } else {
      __threadfence();
      __syncthreads();
      exchange_and_merge(hdr, record, x, u_new, nominal, stats, counter, planes, part_stride, R);
    }
  }
}

// CTA records per block of the one-kernel tail: enough blocks to spread the read of the records, few enough groups for
// the last block's final sums
static int tail_groups(int n_cta, int R) {
  int per = kRescaleMaxCtas;
  if (R < 64) {
    per = 32;
    while (per < kRescaleMaxCtas && (n_cta + per - 1) / per > 64) per *= 2;
  }
  return (n_cta + per - 1) / per;
}

cudaError_t launch_rescale_tail(const DeviceState &d, int mode, cudaStream_t s) {
  const int n_cta = (d.K + 127) / 128;
  const int groups = tail_groups(n_cta, d.R);
  const int per = (n_cta + groups - 1) / groups;
  dim3 grid(groups, d.R);
  float *nominal = d.feedback ? d.nominal : nullptr;
  ExchangeArgs x = exchange_args(d);
#define MPPI_LAUNCH_TAIL(M)                                                                                          \
  le = launch_kernel(d.pdl, rescale_tail_kernel<M>, grid, dim3(256), 0, s, d.hdr, d.cta_part, d.cmin, d.npart,       \
                     d.record, d.u_new, nominal, d.stats, d.counter, d.tail_ticket, d.planes, n_cta, d.rec_stride,   \
                     groups, per, d.R, x)
  cudaError_t le;
  if (mode == 0) {
    MPPI_LAUNCH_TAIL(0);
  } else if (mode == 1) {
    MPPI_LAUNCH_TAIL(1);
  } else {
    MPPI_LAUNCH_TAIL(2);
  }
  if (le != cudaSuccess) return le;
#undef MPPI_LAUNCH_TAIL
  return cudaGetLastError();
}

// Tensor map of the noise tensor for the TMA ring: 2-D {Kp samples, B * R * planes rows} of f32, box {32, kStageSteps*U}.
// cuTensorMapEncodeTiled is taken from the driver through the runtime (no link-time dependency on libcuda).
cudaError_t make_eps_tensor_map(DeviceState &d) {
  d.eps_map_valid = false;
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                               const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess) return e;
  if (!fn || qres != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
  const cuuint64_t gdim[2] = {(cuuint64_t)d.Kp, (cuuint64_t)d.eps_buffers * (cuuint64_t)d.R * (cuuint64_t)d.planes};
  const cuuint64_t gstride[1] = {(cuuint64_t)d.Kp * sizeof(float)};
  const cuuint32_t box[2] = {32u, (cuuint32_t)(kStageSteps * d.U)};
  const cuuint32_t estride[2] = {1u, 1u};
  CUresult r = ((EncodeFn)fn)(&d.eps_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d.eps, gdim, gstride, box, estride,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
  d.eps_map_valid = true;
  return cudaSuccess;
}

cudaError_t launch_rollout_cost_pruned(const DeviceState &d, bool fused, bool write_nearest, cudaStream_t s) {
  if (!d.eps_map_valid) return cudaErrorNotSupported;
  dim3 grid((d.K + 127) / 128, d.R);
  const size_t smem = pruned_tma_smem_bytes(d.T, d.planes, d.U);
  int *nearest = write_nearest ? d.nearest : nullptr;
#define MPPI_LAUNCH_TMA(M, TAP)                                                                                  \
  do {                                                                                                           \
    if (smem > 48 * 1024) {                                                                                      \
      cudaError_t e = cudaFuncSetAttribute(rollout_cost_tma_kernel<M, TAP>,                                      \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
      if (e != cudaSuccess) return e;                                                                            \
    }                                                                                                            \
    cudaError_t le = launch_kernel(d.pdl, rollout_cost_tma_kernel<M, TAP>, grid, dim3(128), smem, s, d.eps_map,  \
                                   d.bounds, d.hdr, d.nominal, d.window, d.state, d.grid_hdr, d.grid_cells, d.cost, \
                                   d.cmin, d.K, d.planes, d.win_stride, d.T, d.grid_max_cells, d.eps, d.Kp,      \
                                   fused ? d.cta_part : nullptr, d.rec_stride, d.counter,                        \
                                   (uint32_t)(d.eps_buffers - 1), d.eps_buf_elems, nearest);                     \
    if (le != cudaSuccess) return le;                                                                            \
  } while (0)
#define MPPI_LAUNCH_TMA_MODEL(TAP)                                                                               \
  switch (d.model) {                                                                                             \
    case kDiffDrive: MPPI_LAUNCH_TMA(kDiffDrive, TAP); break;                                                    \
    case kSteering: MPPI_LAUNCH_TMA(kSteering, TAP); break;                                                      \
    default: MPPI_LAUNCH_TMA(kFullBody, TAP); break;                                                             \
  }
  if (nearest) { MPPI_LAUNCH_TMA_MODEL(true) } else { MPPI_LAUNCH_TMA_MODEL(false) }
#undef MPPI_LAUNCH_TMA_MODEL
#undef MPPI_LAUNCH_TMA
  return cudaGetLastError();
}

int rollout_carveout_percent(const DeviceState &d) {
  const size_t dyn = pruned_tma_smem_bytes(d.T, d.planes, d.U);
  const void *fn = d.model == kDiffDrive ? (const void *)rollout_cost_tma_kernel<kDiffDrive, false>
                   : d.model == kSteering ? (const void *)rollout_cost_tma_kernel<kSteering, false>
                                          : (const void *)rollout_cost_tma_kernel<kFullBody, false>;
  if (dyn > 48 * 1024 &&
      cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess)
    return -1;
  cudaFuncAttributes attr;
  int ctas = 0, dev = 0, max_smem = 0;
  if (cudaFuncGetAttributes(&attr, fn) != cudaSuccess) return -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev) != cudaSuccess) return -1;
  // what the resident CTAs of K2 need -- assuming the split is not the limiter (the driver sizes it the same way)
  const int by_regs = attr.numRegs > 0 ? 65536 / (attr.numRegs * 128) : 16;
  const size_t per_cta = dyn + attr.sharedSizeBytes + 1024;  // + the per-CTA reservation
  ctas = by_regs < 16 ? by_regs : 16;
  while (ctas > 1 && per_cta * ctas > (size_t)max_smem) --ctas;
  int pct = (int)((per_cta * ctas * 100 + max_smem - 1) / max_smem);
  return pct > 100 ? 100 : pct;
}

// worth it only for windows of more than a couple of dozen points; the 16-bit pair fields bound T
bool pruned_scan_supported(int T, int planes) {
  return T >= 24 && T <= 4096 && pruned_tma_smem_bytes(T, planes, kMaxControls) <= 200 * 1024;
}

}  // namespace mppi
