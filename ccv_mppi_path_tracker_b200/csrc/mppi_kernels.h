// mppi_kernels.h -- device-side data layout and kernel launchers of the MPPI core (internal, not the ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mppi_math.h"

namespace mppi {

// Everything that can change from one solve to the next lives in device memory (not in kernel arguments), so
// a captured CUDA graph of the solve can be replayed unchanged: the host rewrites the pinned copy of this
// header and the graph's H2D node carries it over.
struct SolveHeader {
  SolveParams P;
  float inv_lambda;
  uint32_t key0, key1;    // Philox key = 64-bit seed (informational; K1 takes the key as a kernel argument)
  uint32_t robot_offset;  // global index of robot 0 of this handle
  uint32_t q_offset;      // global sample offset of this shard / 4
  uint32_t pad[2];
  // FP64 inputs of the device-side window builder (get_CurrentIndex + calc_RefPath, DD:126-181)
  double v_ref64, dt64, resolution64;
};

// clamp bounds of the controls (DD:24-26, :98-99) as a kernel parameter of K2: constant-bank operands instead of
// registers.  Same FP32 values as SolveParams::u_min / u_max (make_solve_params).
struct ControlBounds {
  float lo[kMaxControls], hi[kMaxControls];
};

// geometry of one robot's candidate grid (written by K0, read by K2): cell (ix, iy) = floor(fma(x, inv_h, cx)), ...
struct GridHeader {
  float x0, y0, h, inv_h, cx, cy;
  int nx, ny;
};

// HBM layout owned by one handle.  R robots, K samples (this shard), T horizon, U controls, P = (T-1)*U planes.
//   inbuf    one allocation; a solve copies the prefix it needs (host -> device), in this order:
//     hdr      SolveHeader (padded to 256 B)
//     state64  [R][2]    pose x, y in FP64 (device-side window builder)
//     state    [R][4]    {yaw, roll, pitch, -} (the kernels work in the robot-centred frame: x = y = 0)
//     nominal  [R][P]    warm start u*; the merge kernel overwrites it in place with the new controls.  Copied only
//                        when the host supplies the warm start (mppi_upload with u_nominal, MPPI_OPT_UPLOAD_WARM_START)
//     window   [R][WS]   WS floats: T x {x_ref - x0, y_ref - y0} (robot-centred frame, FP32), padded to 16 B.  Copied
//                        only when windows are built on the host; the device builder (K-1) writes it in place
//   eps   [B][R][P][Kp]  standard normals, plane-major: one warp reads 32 consecutive samples of one (t,u)
//                        plane (128 B coalesced); Kp = K rounded up to 4 (float4 stores of the generator).
//                        B = eps_buffers: 2 when the next solve's normals are generated beside the current solve
//                        (noise prefetch); solve n uses buffer n & (B-1), n = the device-side solve counter
//   cost     [R][K]      per-sample cost;  weight [R][K] exp(-(c - c_min)/lambda)
//   cmin     [R]         ordered-uint encoding of the minimum cost (atomicMin target of K2); re-armed (all ones) by
//                        the last kernel of a solve that reads it, i.e. by the tail
//   wpart    [R][NB3][2] per-block partial (sum w, sum w^2) of the weight kernel (fixed-order final sum)
//   npart    [R][P][NC]  per-chunk partial numerators of the weighted control reduction
//   record   [R][REC]    REC = 4 + P floats: {c_min, sum w, sum w^2, -, N[P]} -- the exchanged partial
//   gathered [G][R][REC] all ranks' records (G = 1 when not sharded: aliased to record)
//   outbuf   one allocation, one D2H copy per solve:  u_new [R][P], stats [R][4] = {c_min, sum w, ESS, -}
//   nearest  [R][K][T]   int32, debug only
//   counter  [1]         solve counter of the Philox stream, advanced on the device by the merge kernel
//   grid_hdr [R], grid_cells [R][grid_max_cells]  candidate grid of the pruned scan (K0 -> K2): per cell the
//                        range of window-point pairs that can hold the nearest point, lo | n << 16
struct DeviceState {
  int model = 0, T = 0, U = 0;
  int K = 0, Kp = 0, R = 0, planes = 0, win_stride = 0, rec_stride = 0;
  int nb3 = 0, nchunk = 0, n_ranks = 1;
  uint32_t key0 = 0, key1 = 0;  // Philox key = 64-bit seed (kernel argument of K1)
  SolveHeader *hdr = nullptr;
  float *window = nullptr, *state = nullptr, *nominal = nullptr;
  float *eps = nullptr, *cost = nullptr, *weight = nullptr, *wpart = nullptr, *npart = nullptr;
  int eps_buffers = 1;        // 1 or 2 (noise prefetch); buffer of solve n = n & (eps_buffers - 1)
  size_t eps_buf_elems = 0;   // R * planes * Kp
  bool feedback = true;       // the merge writes the new controls into `nominal` (warm start of the next solve)
  long long xchg_timeout_cycles = 4000000000ll;  // peer exchange: device-side wait limit of the merge
  float *record = nullptr, *gathered = nullptr;
  float *u_new = nullptr, *stats = nullptr;
  unsigned int *cmin = nullptr;
  uint32_t *counter = nullptr;
  int *nearest = nullptr;
  float *states_dbg = nullptr;  // [R][K][T][5] predicted states, debug only (MPPI_DEBUG_STATES)
  // device-side window builder (K-1): all robots' paths concatenated, FP64
  const double *path_xy = nullptr;   // [sum n_r][2]
  const int *path_off = nullptr;     // [R + 1]
  const double *state64 = nullptr;   // [R][2] pose x, y (in inbuf)
  const unsigned char *win_fixed = nullptr;  // [R] 1 = window given by mppi_set_window (host-built), skip
  int *cur_index = nullptr;          // [R] current_index_ of the last solve
  // NVLink record exchange (C1 without NCCL): this rank's buffer, the peers' buffers (IPC-mapped), sequence
  void *xchg_buf = nullptr;
  void *const *xchg_peers = nullptr;  // [G] device array of exchange-buffer base pointers (own one included)
  unsigned int *xchg_seq = nullptr, *xchg_ticket = nullptr;  // solve sequence; two "last block" tickets
  int xchg_rank = 0;
  GridHeader *grid_hdr = nullptr;
  uint32_t *grid_cells = nullptr;
  int grid_max_cells = 0;
  float grid_h_min = 0.05f;   // K = 2^20: flat below 0.05 (0.025 .. 0.10 tried); small handles are bound by max_cells
  float grid_margin = 0.f;    // > 0: absolute margin around the window's bounding box; else a quarter of the horizon reach
  int grid_lanes = 0;         // > 0: lanes per cell of K0 (tuning); else chosen by the handle's total cell count
  bool fuse_controls = false;  // K2 also produces the weighted-control records of its CTAs; false: K3 + K4
  float *cta_part = nullptr;  // [R][ceil(K/128)][rec_stride] per-CTA records {m, S, Q, -, N[P]} of the fused K2
  int k4_groups = 0;  // > 0: plane groups per K4 block (tuning); else by the number of blocks
  unsigned int *tail_ticket = nullptr;  // [R + 1] "last block" tickets of the one-kernel tail (fused-controls path)
  bool pdl = false;           // launch K2 / K3 / K4 as programmatic dependents of the kernel in front of them
  int sm_count = 148;         // SMs of the device (cudaDevAttrMultiProcessorCount)
  int side_carveout = -1;     // shared-memory carve-out (per cent) given to the side-stream kernels: K2's own
  ControlBounds bounds = {};  // kernel parameter of K2 (kept equal to hdr->P.u_min / u_max by the host)
  bool eps_map_valid = false;
  CUtensorMap eps_map;  // 2-D {Kp, eps_buffers * R * planes} f32, box {32, 4 * U}
};

// Preferred L1 / shared-memory split of a kernel, in per cent of the maximum shared memory (< 0: leave the default).
// Two kernels with different carve-outs are never resident on one SM at the same time, so the side stream's kernels
// (candidate grid, noise prefetch) are given K2's split (DeviceState::side_carveout) -- otherwise they run before or
// after K2 instead of beside it.  Remembers the last value per device and function.
cudaError_t set_carveout(const void *kernel, int percent);
// the split the driver picks for the K2 instantiation of this handle (mppi_rollout_pruned.cu)
int rollout_carveout_percent(const DeviceState &d);

// kernel launch with or without the programmatic-dependent-launch attribute (see pdl_wait() in mppi_device.cuh)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                 Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// K2 (128-thread CTAs, 8 resident per SM by registers) fits the machine in a single wave
inline bool rollout_is_one_wave(const DeviceState &d) {
  return (long long)d.R * ((d.K + 127) / 128) <= (long long)8 * d.sm_count;
}

constexpr int kHeaderBytes = 256;
constexpr int kStateStride = 4;     // floats per robot of the state record {yaw, roll, pitch, -} (x = y = 0: robot frame)
constexpr int kExchangeHeaderBytes = 256;  // reserved head of an exchange buffer (keeps the slots 256-byte aligned)
constexpr int kWeightBlock = 256;   // threads of the weight kernel, 4 samples per thread
constexpr int kReduceBlock = 256;   // threads of the weighted-control reduction
constexpr int kReduceChunk = 4096;  // samples per (plane, chunk) block of the reduction
constexpr int kRescaleMaxCtas = 128;  // most CTA records folded by one block of rescale_tail_kernel
static_assert(sizeof(SolveHeader) <= kHeaderBytes, "SolveHeader must fit its slot");
static_assert(kRescaleMaxCtas % 32 == 0, "the tail reads the records thirty-two at a time");

// K1  Philox4x32-10 + Box-Muller -> eps (float4 stores) of solve (counter + ahead), into that solve's buffer.
cudaError_t launch_noise(const DeviceState &d, int ahead, cudaStream_t s);
// K2  fused rollout + cost (+ block min -> atomicMin on cmin).  scan_mode: 1 literal, 2 pruned.
cudaError_t launch_rollout_cost(const DeviceState &d, int scan_mode, bool write_nearest, bool write_states, bool fused,
                                cudaStream_t s);
// K-1 get_CurrentIndex + calc_RefPath on the device, one CTA per robot (many-robot handles; FP64 like the host path)
cudaError_t launch_window_builder(const DeviceState &d, cudaStream_t s);
// K0  candidate grid of the pruned scan, once per robot and solve (mppi_rollout_pruned.cu)
cudaError_t launch_candidate_grid(const DeviceState &d, cudaStream_t s);
// K2 production variant (mppi_rollout_pruned.cu): exact pruned nearest-point scan, bit-identical costs
// fused: the CTAs also write their weighted-control records into d.cta_part; launch_rescale_tail then replaces
// K3 + K4 + K5 + K6.  write_nearest: the instantiation that records the argmin index (MPPI_DEBUG_NEAREST)
cudaError_t launch_rollout_cost_pruned(const DeviceState &d, bool fused, bool write_nearest, cudaStream_t s);
// The whole tail of the fused-controls path in ONE launch: rescale of the per-CTA records, fixed-order final sums,
// then (last block) the record exchange with the peers over NVLink (p2p) or the local merge -> u_new, warm start,
// stats, counter++.  mode: 0 = unsharded, 1 = peer exchange, 2 = record only (NCCL all-gather + merge follow).
cudaError_t launch_rescale_tail(const DeviceState &d, int mode, cudaStream_t s);
bool pruned_scan_supported(int T, int planes);
// tensor map of d.eps for K2's TMA ring (after d.eps, Kp, R, planes, U are final)
cudaError_t make_eps_tensor_map(DeviceState &d);
// K3  weights w = exp(-(c - c_min)/lambda), per-block partial sums.  after_solve: the on-demand debug tap (c_min from
// the record; the tail of the solve has re-armed cmin by then)
cudaError_t launch_weights(const DeviceState &d, bool after_solve, cudaStream_t s);
// planes per block of K4 (1, 2 or 4) and the matching number of sample chunks: nchunk = ceil(Kp / (4096 / ppb))
int reduce_planes_per_block(int Kp);
// K4  weighted control reduction partials; fuse_weights: K3 folded in (small K; then nb3 must equal nchunk)
cudaError_t launch_weighted_controls(const DeviceState &d, bool fuse_weights, cudaStream_t s);
// K5  fixed-order final sums -> record;  K6 merge of G records -> u_new, nominal, stats, counter++
cudaError_t launch_finalize(const DeviceState &d, cudaStream_t s);
cudaError_t launch_merge(const DeviceState &d, cudaStream_t s);
// C1 over NVLink peer memory instead of NCCL: K5 whose last block pushes the record into every peer's exchange
// buffer, waits for the peers' flags and merges (one process per GPU, buffers shared through CUDA IPC)
cudaError_t launch_finalize_exchange(const DeviceState &d, cudaStream_t s);
// what the exchanging block needs (kernel argument)
struct ExchangeArgs {
  void *const *peers;     // [G] exchange-buffer base pointers of all ranks (own one included)
  void *xbuf;             // this rank's exchange buffer
  unsigned int *seq;      // solve sequence number of the exchange (parity selects the slot set)
  unsigned int *ticket;   // "last block" ticket
  int rank, G;
  long long timeout_cycles;
};
ExchangeArgs exchange_args(const DeviceState &d);
// K5 + K6 in one launch when the handle is not sharded (identical results)
cudaError_t launch_finalize_merge(const DeviceState &d, cudaStream_t s);

// ordered-uint encoding so that atomicMin(unsigned) orders floats (negative costs included)
MPPI_HD uint32_t float_to_ordered(float f) {
  uint32_t u = float_to_bits(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
MPPI_HD float ordered_to_float(uint32_t u) {
  return bits_to_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// Merge of per-rank partial records, shared by the device merge kernel and mppi_merge_partials (host).
//   rec_g = {m_g, S_g, Q_g, -, N_g[n]} with S_g = sum exp(-(c-m_g)/lambda), Q_g = sum of squares.
//   m = min m_g;  a_g = exp(-(m_g - m)/lambda);  S = sum a_g S_g;  N = sum a_g N_g;  u = N / S.
// Ranks are combined in rank order (deterministic).
MPPI_HD float merge_min(const float *recs, int n_ranks, size_t rank_stride) {
  float m = recs[0];
  for (int g = 1; g < n_ranks; ++g) {
    float mg = recs[(size_t)g * rank_stride];
    m = mg < m ? mg : m;
  }
  return m;
}
MPPI_HD float merge_scale(float m_g, float m, float inv_lambda, int n_ranks) {
  return n_ranks == 1 ? 1.f : expf(-(m_g - m) * inv_lambda);
}
MPPI_HD void merge_sums(const float *recs, int n_ranks, size_t rank_stride, float m, float inv_lambda, float &S,
                        float &Q) {
  S = 0.f;
  Q = 0.f;
  for (int g = 0; g < n_ranks; ++g) {
    const float *r = recs + (size_t)g * rank_stride;
    float a = merge_scale(r[0], m, inv_lambda, n_ranks);
    S = fmaf(a, r[1], S);
    Q = fmaf(a * a, r[2], Q);
  }
}
MPPI_HD float merge_numerator(const float *recs, int n_ranks, size_t rank_stride, float m, float inv_lambda, int p) {
  float N = 0.f;
  for (int g = 0; g < n_ranks; ++g) {
    const float *r = recs + (size_t)g * rank_stride;
    N = fmaf(merge_scale(r[0], m, inv_lambda, n_ranks), r[4 + p], N);
  }
  return N;
}

}  // namespace mppi
