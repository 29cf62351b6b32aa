// mppi_capi.cu -- implementation of the C ABI declared in include/mppi_b200.h.
// Host side of one control cycle: window construction in double (get_CurrentIndex + calc_RefPath, DD:126-181),
// one pinned staging block -> one H2D copy, the kernel sequence K1..K6 (optionally replayed from a CUDA graph),
// one D2H copy.  No CPU fallback: without a usable CUDA device every compute entry point returns MPPI_ERR_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cmath>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mppi_b200.h"
#include "mppi_host.h"
#include "mppi_kernels.h"
#include "philox.h"

using namespace mppi;

// ---- minimal NCCL surface, resolved at run time with dlopen so that single-GPU users need no NCCL ----------
namespace {
struct NcclUniqueId {
  char internal[MPPI_COMM_ID_BYTES];
};
typedef int (*nccl_get_unique_id_fn)(NcclUniqueId *);
typedef int (*nccl_comm_init_rank_fn)(void **, int, NcclUniqueId, int);
typedef int (*nccl_all_gather_fn)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef int (*nccl_comm_destroy_fn)(void *);
typedef const char *(*nccl_get_error_string_fn)(int);
constexpr int kNcclFloat = 7;

struct NcclApi {
  void *lib = nullptr;
  nccl_get_unique_id_fn get_unique_id = nullptr;
  nccl_comm_init_rank_fn comm_init_rank = nullptr;
  nccl_all_gather_fn all_gather = nullptr;
  nccl_comm_destroy_fn comm_destroy = nullptr;
  nccl_get_error_string_fn get_error_string = nullptr;
  std::string err;
  bool load() {
    if (lib) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) {
      err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
      return false;
    }
    get_unique_id = (nccl_get_unique_id_fn)dlsym(lib, "ncclGetUniqueId");
    comm_init_rank = (nccl_comm_init_rank_fn)dlsym(lib, "ncclCommInitRank");
    all_gather = (nccl_all_gather_fn)dlsym(lib, "ncclAllGather");
    comm_destroy = (nccl_comm_destroy_fn)dlsym(lib, "ncclCommDestroy");
    get_error_string = (nccl_get_error_string_fn)dlsym(lib, "ncclGetErrorString");
    if (!get_unique_id || !comm_init_rank || !all_gather || !comm_destroy) {
      err = "libnccl is missing a required symbol";
      lib = nullptr;
      return false;
    }
    return true;
  }
};
NcclApi g_nccl;
thread_local std::string g_create_error;
}  // namespace

struct mppi_handle_s {
  int model = 0, K = 0, T = 0, U = 0, S = 0, R = 0, device = 0;
  mppi_params params;
  DeviceState d;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t staged = nullptr;  // H2D of the staging block has completed
  bool staged_pending = false;
  // one pinned staging block mirroring d_in: header | windows | states | nominals
  char *h_in = nullptr, *d_in = nullptr;
  size_t in_bytes = 0, state_off = 0, nominal_off = 0, window_off = 0;
  size_t last_h2d_bytes = 0;       // what the last upload / solve copied host -> device
  bool host_windows_staged = true; // the staged block carries host-built windows (must be copied)
  int opt_upload_warm_start = 1;   // mppi_solve copies the caller's u_nominal to the device (0: keeps the device's own)
  float *d_window2 = nullptr;  // second slot of the device-built windows (the first one lives in d_in)
  unsigned issue_idx = 0;      // stream-launched solves issued so far: its parity selects the slot
  int last_slot = -1;          // slot of the last stream-launched solve
  float *h_out = nullptr, *d_out = nullptr;
  bool out_mapped = false;  // small results: the tail kernel writes them straight into the pinned host block
  size_t out_bytes = 0;
  // host-side per-robot inputs
  std::vector<std::vector<double>> path;
  std::vector<char> window_fixed;
  std::vector<char> window_yaw_stale;  // the host-built window's yaw_ref column still has to be derived (mppi_get_window)
  std::vector<double> window;  // [R][T][3]
  std::vector<int> cur_index;
  bool have_inputs = false, have_nominal = false;
  // device-side window builder (many-robot handles)
  int window_builder = MPPI_WINDOW_AUTO;
  bool paths_dirty = false;
  double *d_path = nullptr;
  int *d_path_off = nullptr;
  unsigned char *d_win_fixed = nullptr;
  size_t d_path_capacity = 0;
  size_t state64_off = 0;
  std::vector<double> last_state;  // [R][S] pose of the last staged solve (mppi_get_window)
  double last_dt = 0.0;
  bool external_noise = false;
  bool weights_valid = false;  // d.weight holds the last solve's weights (false on the fused-controls path)
  // options (mppi_set_option)
  int opt_fuse_controls = -1, opt_prefetch = -1;
  double opt_timeout_ms = 2000.0;
  // noise prefetch: the current solve's normals are already in their buffer (generated beside the previous solve)
  bool noise_primed = false;
  bool last_fused = false;  // the kernel sequence issued (or captured) last reduces the controls inside K2
  int sm_clock_khz = 1965000;
  int debug_flags = 0, scan_mode = MPPI_SCAN_AUTO;
  uint64_t seed = 0x5EED0000ull;
  int64_t sample_offset = 0, k_global = 0;
  int robot_offset = 0;
  // K0 (candidate grid) runs beside K1 (noise): they are independent
  cudaStream_t side_stream = nullptr;
  // ev_fork / ev_join_cap are only ever recorded inside a stream capture, the others only outside (an event whose
  // last record was captured cannot be waited for by an ordinary stream operation)
  cudaEvent_t ev_fork = nullptr, ev_join_cap = nullptr, ev_grid = nullptr, ev_join = nullptr;
  cudaEvent_t ev_readers[2] = {nullptr, nullptr};  // by slot: the last reader of that slot's window / grid / normals is done
  bool readers_recorded[2] = {false, false};
  bool staged_recorded = false, join_recorded = false;  // the events have been recorded at least once
  // graphs
  bool use_graph = false;
  cudaGraphExec_t exec_kernels = nullptr, exec_solve = nullptr;
  int launch_count = 0;
  // collective
  void *comm = nullptr;
  int rank = 0, n_ranks = 1;
  // NVLink peer exchange (instead of NCCL): own buffer, peers' IPC mappings
  bool p2p = false;
  int xchg_ranks = 0;
  void *xchg_buf = nullptr;
  std::vector<void *> xchg_peer_ptrs;
  void **d_xchg_peers = nullptr;
  std::string err;
};

namespace {

int fail(mppi_handle h, int code, const std::string &msg) {
  if (h) h->err = msg;
  else g_create_error = msg;
  return code;
}
#define CU_TRY(h, expr)                                                                                  \
  do {                                                                                                   \
    cudaError_t e__ = (expr);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return fail(h, MPPI_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));                \
  } while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// u_min <= u_max: the clamp of the FP32 contract is min(max(v, lo), hi) (mppi_math.h), which equals the reference's
// `if (v < lo) v = lo; else if (v > hi) v = hi;` (DD:62-67) only for ordered bounds
bool params_valid(const mppi_params *p) {
  if (!(p && p->lambda > 0.0 && p->resolution > 0.0 && p->control_noise >= 0.0)) return false;
  for (int u = 0; u < kMaxControls; ++u)
    if (!(p->u_min[u] <= p->u_max[u])) return false;
  return true;
}

void invalidate_graphs(mppi_handle h) {
  if (h->exec_kernels) cudaGraphExecDestroy(h->exec_kernels);
  if (h->exec_solve) cudaGraphExecDestroy(h->exec_solve);
  h->exec_kernels = h->exec_solve = nullptr;
}

// K3 folded into K4 (every K4 block recomputes the weights of its samples): pays off while K * planes is small
bool fused_weights(const mppi_handle_s *h) { return h->K <= 32768; }  // re-measured in round 2: see DESIGN.md

bool device_windows(mppi_handle h) {
  return h->window_builder == MPPI_WINDOW_DEVICE || (h->window_builder == MPPI_WINDOW_AUTO && h->R >= 8);
}

// (re)upload all robots' paths, concatenated, when any of them changed
int flush_paths(mppi_handle h) {
  if (!h->paths_dirty) return MPPI_OK;
  std::vector<int> off(h->R + 1, 0);
  for (int r = 0; r < h->R; ++r) off[r + 1] = off[r] + (int)(h->path[r].size() / 2);
  std::vector<double> all((size_t)2 * off[h->R] + 2, 0.0);
  for (int r = 0; r < h->R; ++r)
    if (!h->path[r].empty()) memcpy(all.data() + (size_t)2 * off[r], h->path[r].data(), sizeof(double) * h->path[r].size());
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  if (all.size() > h->d_path_capacity) {
    invalidate_graphs(h);  // the buffer address is a kernel argument of the captured window node
    if (h->d_path) cudaFree(h->d_path);
    h->d_path = nullptr;
    h->d_path_capacity = all.size() + all.size() / 4;
    CU_TRY(h, cudaMalloc((void **)&h->d_path, sizeof(double) * h->d_path_capacity));
    h->d.path_xy = h->d_path;
  }
  CU_TRY(h, cudaMemcpy(h->d_path, all.data(), sizeof(double) * all.size(), cudaMemcpyHostToDevice));
  CU_TRY(h, cudaMemcpy(h->d_path_off, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice));
  CU_TRY(h, cudaMemcpy(h->d_win_fixed, h->window_fixed.data(), (size_t)h->R, cudaMemcpyHostToDevice));
  h->paths_dirty = false;
  return MPPI_OK;
}

void fill_header(mppi_handle h, double dt) {
  SolveHeader *hd = reinterpret_cast<SolveHeader *>(h->h_in);
  hd->v_ref64 = h->params.v_ref;
  hd->dt64 = dt;
  hd->resolution64 = h->params.resolution;
  hd->P = make_solve_params(h->model, h->T, h->params, dt);
  hd->inv_lambda = (float)(1.0 / h->params.lambda);
  hd->key0 = (uint32_t)h->seed;
  hd->key1 = (uint32_t)(h->seed >> 32);
  hd->robot_offset = (uint32_t)h->robot_offset;
  hd->q_offset = (uint32_t)(h->sample_offset / 4);
  // K2 takes the clamp bounds as a kernel parameter: keep them equal to the header's (a captured graph holds a copy)
  ControlBounds b;
  for (int u = 0; u < kMaxControls; ++u) {
    b.lo[u] = hd->P.u_min[u];
    b.hi[u] = hd->P.u_max[u];
  }
  if (memcmp(&b, &h->d.bounds, sizeof b) != 0) {
    h->d.bounds = b;
    invalidate_graphs(h);
  }
}

// which nearest-point scan the next solve runs: the literal kernel is the one that records the predicted states
int effective_scan(mppi_handle h, bool *want_nearest, bool *want_states = nullptr) {
  *want_nearest = (h->debug_flags & MPPI_DEBUG_NEAREST) && h->d.nearest;
  const bool ws = (h->debug_flags & MPPI_DEBUG_STATES) && h->d.states_dbg;
  if (want_states) *want_states = ws;
  int scan = h->scan_mode == MPPI_SCAN_AUTO ? MPPI_SCAN_PRUNED : h->scan_mode;
  if (ws || !pruned_scan_supported(h->d.T, h->d.planes) || !h->d.eps_map_valid) scan = MPPI_SCAN_LITERAL;
  return scan;
}

// K2 also produces the weighted-control records of its CTAs (then K3 + K4 + K5 + K6 become the one-kernel tail).
// AUTO: many-robot handles (K4 is far from the HBM roofline on many small tensors), and single solves whose K2 grid
// covers the machine (>= 256 CTAs, K >= 32768): the CTAs re-read their tile of the normals from L2 right after
// streaming it, in the shadow of the CTAs that are still rolling out and of the next solve's generator, instead of a
// second pass of the whole tensor through HBM plus two more dependent launches.  Measured with the noise prefetch on
// (round 2, back-to-back solves): 0.517 vs 0.562 ms at K = 2^20, 0.096 vs 0.103 ms at 2^17, 0.062 vs 0.070 at 2^16.
// Tiny K keeps K3 + K4: the pass would be serialised inside a few CTAs (K = 4096: 60 vs 51 us per synchronous solve).
bool fused_controls(mppi_handle h, int scan) {
  if (scan != MPPI_SCAN_PRUNED || !h->d.cta_part) return false;
  if (h->opt_fuse_controls >= 0) return h->opt_fuse_controls != 0;
  if (h->R >= 8) return true;
  return (long long)h->R * ((h->K + 127) / 128) >= 256;
}
// noise prefetch: needs the second buffer (allocated at mppi_create when it fits) and the internal generator; the
// pruned scan only (its K0 resets the minimum-cost slot; the literal path is the debug / tiny-horizon path)
bool prefetch_on(mppi_handle h, int scan) {
  return h->d.eps_buffers == 2 && !h->external_noise && scan == MPPI_SCAN_PRUNED && h->opt_prefetch != 0;
}
// K5 + K6 as one launch (one block per robot): unsharded handles whose partial arrays are small (latency path,
// many-robot handles); large-K handles keep the wide finalize + merge pair
bool fused_tail_for(mppi_handle h, const DeviceState &t) {
  return h->n_ranks == 1 && (long long)t.planes * t.nchunk <= 4096;
}

// the kernel sequence of one solve on stream s (also what gets captured into the graph)
//   plain stream launches (back-to-back throughput):
//     main:  . . . . . . . . . . . . . . K2 -> (K3 -> K4 | -) -> tail (finalize / exchange / merge; counter++)
//     side:  [K-1 window] -> K0 candidate grid -> K1 noise of the NEXT solve (prefetch)
//     The side stream starts as soon as ITS inputs are there -- the staged pose / window of this solve, the end of the
//     last reader of this solve's window / grid slot (two solves ago: the slots alternate) and of the noise buffer --
//     so window builder, candidate grid and generator run under the previous solve.  K2 waits for K0.
//     K3, K4 and the tail are programmatic dependents of the kernel in front of them (the one-kernel tail only behind
//     a single-wave K2), and K2 is a programmatic dependent of the PREVIOUS solve's tail: launch latency, header /
//     window staging and barrier set-up overlap the predecessor (pdl_wait() in the kernels).
//   inside a CUDA graph (the synchronous latency path; nothing of another solve to overlap with):
//     main:  [K-1] -> K0 -> K2 -> (K3 -> K4 | -) -> tail          side:  K1 noise of the next solve
//     K2, K3, K4 and the tail are programmatic dependents of the kernel in front of them.
// The generator keeps its own solve index on the device (counters[1]): nothing orders it against the tail, which
// advances the solve counter; the NEXT solve's K2 waits for it (ev_join).
#ifndef MPPI_K2_STREAM_PDL
#define MPPI_K2_STREAM_PDL 1
#endif
int issue_kernels(mppi_handle h, cudaStream_t s, bool capturing) {
  // Stream-launched solves alternate between two slots of window / grid (and, through the solve counter, of the
  // normals), so that the side stream can prepare solve n+1 while K2 of solve n still reads its own slot; a graph
  // always uses slot 0.
  // (windows given by mppi_set_window live in the staging block, i.e. in slot 0 only: such handles stay there)
  const bool alternate = !capturing && !(device_windows(h) && h->host_windows_staged);
  const int slot = alternate ? (int)(h->issue_idx & 1u) : 0;
  const int prev_slot = h->last_slot;  // slot of the previous stream-launched solve, -1: none
  DeviceState d = h->d;
  d.grid_hdr += (size_t)slot * d.R;
  d.grid_cells += (size_t)slot * d.R * d.grid_max_cells;
  if (slot && device_windows(h)) d.window = h->d_window2;
  DeviceState dp = d;  // launch descriptor for the kernels that may start under their predecessor
  dp.pdl = true;
  int n = 0;
  bool want_nearest, want_states;
  const int scan = effective_scan(h, &want_nearest, &want_states);
  const bool prefetch = prefetch_on(h, scan);
  const bool dev_win = device_windows(h);
  const bool pruned = scan == MPPI_SCAN_PRUNED;
  const bool side = pruned && (!capturing || prefetch);  // something runs on the side stream
  cudaStream_t gs = (pruned && !capturing) ? h->side_stream : s;  // where window builder and candidate grid run
  // the normals of THIS solve were generated on the side stream during the previous one: K2 needs them complete
  if (!capturing && h->join_recorded) CU_TRY(h, cudaStreamWaitEvent(s, h->ev_join, 0));
  if (side) {
    if (capturing) {
      CU_TRY(h, cudaEventRecord(h->ev_fork, s));
      CU_TRY(h, cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
    } else {
      if (h->staged_recorded) CU_TRY(h, cudaStreamWaitEvent(h->side_stream, h->staged, 0));
      // this slot's window and grid were last read two solves ago
      if (h->readers_recorded[slot]) CU_TRY(h, cudaStreamWaitEvent(h->side_stream, h->ev_readers[slot], 0));
    }
  }
  if (dev_win) {
    CU_TRY(h, launch_window_builder(d, gs));
    ++n;
  }
  if (pruned) {
    CU_TRY(h, launch_candidate_grid(d, gs));
    ++n;
    if (!capturing) CU_TRY(h, cudaEventRecord(h->ev_grid, h->side_stream));
    if (prefetch) {
      // the buffer of the next solve's normals was last read by the previous solve
      if (!capturing && prev_slot >= 0 && prev_slot != slot)
        CU_TRY(h, cudaStreamWaitEvent(h->side_stream, h->ev_readers[prev_slot], 0));
      CU_TRY(h, launch_noise(d, 1, h->side_stream));
      CU_TRY(h, cudaEventRecord(capturing ? h->ev_join_cap : h->ev_join, h->side_stream));
      if (!capturing) h->join_recorded = true;
      ++n;
    }
  }
  if (!h->external_noise && !prefetch) {
    CU_TRY(h, launch_noise(d, 0, s));
    ++n;
  }
  if (pruned && !capturing) CU_TRY(h, cudaStreamWaitEvent(s, h->ev_grid, 0));
  const bool fused = fused_controls(h, scan);
  // K2 directly behind K0 on the same stream (graph, no generator in between): a programmatic dependent.  With stream
  // launches it follows the previous solve's tail (event waits in between): its CTAs load header and window and set
  // up their barriers while that tail still runs, and read warm start, counter and grid only after it.
  const bool k2_pdl = pruned && (h->external_noise || prefetch) && (capturing || MPPI_K2_STREAM_PDL);
  CU_TRY(h, launch_rollout_cost(k2_pdl ? dp : d, scan, want_nearest, want_states, fused, s));
  ++n;
  h->last_fused = fused;
  h->weights_valid = !fused;  // the per-sample weights are a debug tap on the fused path (mppi_get_weights)
  if (!fused) {
    if (!fused_weights(h)) {
      CU_TRY(h, launch_weights(dp, false, s));
      ++n;
    }
    CU_TRY(h, launch_weighted_controls(dp, fused_weights(h), s));
    ++n;
  }
  if (!capturing) {  // the last reader of this solve's window, grid and normals is queued
    CU_TRY(h, cudaEventRecord(h->ev_readers[slot], s));
    h->readers_recorded[slot] = true;
    h->last_slot = slot;
    ++h->issue_idx;
  }
  // inside a graph the generator has to rejoin the capturing stream (it does not depend on the tail any more: it
  // keeps its own solve index); with stream launches the NEXT solve's K2 waits for it (top of this function)
  auto join = [&]() -> int {
    if (capturing && prefetch) CU_TRY(h, cudaStreamWaitEvent(s, h->ev_join_cap, 0));
    return MPPI_OK;
  };
  const DeviceState &dt = (capturing && !fused) ? dp : d;  // the tail directly behind K4: a programmatic dependent
  h->launch_count = n;
  if (fused) {  // rescale + finalize + merge (+ exchange) in one launch
    const int mode = h->p2p ? 1 : (h->n_ranks > 1 ? 2 : 0);
    // Directly behind K2: a programmatic dependent when K2 is a single wave of CTAs -- the tail's blocks are parked on
    // the SMs, the solve's constants loaded, when K2's last CTA retires (K = 2^17: 91.3 -> 90.5 us per solve, 2^16:
    // 61.7 -> 58.9).  Behind a K2 of several waves the parked blocks only take SM room from the side stream's
    // generator (K = 2^20: 0.515 -> 0.522 ms; 1024 robots: 0.376 -> 0.385 ms): an ordinary launch there.
    CU_TRY(h, launch_rescale_tail(rollout_is_one_wave(d) ? dp : d, mode, s));
    h->launch_count = ++n;
    if (mode != 2) return join();
  } else {
    if (h->p2p) {
      CU_TRY(h, launch_finalize_exchange(dt, s));
      h->launch_count = ++n;
      return join();
    }
    if (fused_tail_for(h, d)) {
      CU_TRY(h, launch_finalize_merge(dt, s));
      h->launch_count = ++n;
      return join();
    }
    CU_TRY(h, launch_finalize(dt, s));
    ++n;
  }
  if (h->n_ranks > 1) {
    int rc = g_nccl.all_gather(d.record, d.gathered, (size_t)d.R * d.rec_stride, kNcclFloat, h->comm, s);
    if (rc != 0)
      return fail(h, MPPI_ERR_NCCL,
                  std::string("ncclAllGather: ") + (g_nccl.get_error_string ? g_nccl.get_error_string(rc) : "?"));
    ++n;
  }
  CU_TRY(h, launch_merge(d, s));
  h->launch_count = ++n;
  return join();
}

// Runs in front of a solve (outside any graph): with the prefetch on, the normals of the solve that is about to
// run must already be in their buffer; they are not after mppi_create, mppi_set_seed, mppi_set_shard or a switch of
// noise source / scan mode -- generate them now, in stream order before the solve.
int prime_noise(mppi_handle h, bool header_staged_only) {
  bool wn;
  const int scan = effective_scan(h, &wn);
  if (!prefetch_on(h, scan)) {
    h->noise_primed = false;  // the next prefetching solve has to start from a fresh buffer
    return MPPI_OK;
  }
  if (h->noise_primed) return MPPI_OK;
  // a generator of an abandoned prefetch may still be running on the side stream (it advances its own solve index)
  if (h->join_recorded) CU_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
  if (header_staged_only)  // the generator reads the shard offsets from the device header: bring it over first
    CU_TRY(h, cudaMemcpyAsync(h->d_in, h->h_in, kHeaderBytes, cudaMemcpyHostToDevice, h->stream));
  CU_TRY(h, launch_noise(h->d, 0, h->stream));
  h->noise_primed = true;
  return MPPI_OK;
}

int wait_staging_free(mppi_handle h) {
  if (h->staged_pending) {
    CU_TRY(h, cudaEventSynchronize(h->staged));
    h->staged_pending = false;
  }
  return MPPI_OK;
}

// host part of the cycle: windows (double), header, robot-centred FP32 inputs into the pinned block
int stage_inputs(mppi_handle h, const double *state, double dt, const double *u_nominal) {
  if (!state || !(dt > 0.0) || !std::isfinite(dt)) return fail(h, MPPI_ERR_INVALID, "state must be non-NULL and dt > 0");
  // The reference lets a NaN pose run through the cycle and publishes NaN commands (DD:117 -> DD:250); here it is an
  // error the caller can see: the kernels' grid lookup and min() do not propagate NaN the way the FP64 loops do.
  for (size_t k = 0; k < (size_t)h->R * h->S; ++k)
    if (!std::isfinite(state[k])) return fail(h, MPPI_ERR_INVALID, "state must be finite");
  if (u_nominal)  // a NaN warm start would live on in the device-resident controls for ever
    for (size_t k = 0; k < (size_t)h->R * h->d.planes; ++k)
      if (!std::isfinite(u_nominal[k])) return fail(h, MPPI_ERR_INVALID, "u_nominal (warm start) must be finite");
  int rc = wait_staging_free(h);
  if (rc) return rc;
  fill_header(h, dt);
  const DeviceState &d = h->d;
  float *win = reinterpret_cast<float *>(h->h_in + h->window_off);
  float *st = reinterpret_cast<float *>(h->h_in + h->state_off);
  float *nom = reinterpret_cast<float *>(h->h_in + h->nominal_off);
  double *st64 = reinterpret_cast<double *>(h->h_in + h->state64_off);
  const bool on_device = device_windows(h);
  if (on_device) {
    rc = flush_paths(h);
    if (rc) return rc;
  }
  memcpy(h->last_state.data(), state, sizeof(double) * (size_t)h->R * h->S);
  h->last_dt = dt;
  h->host_windows_staged = !on_device;
  for (int r = 0; r < h->R; ++r) {
    const double *s = state + (size_t)r * h->S;
    double *w = h->window.data() + (size_t)r * h->T * 3;
    if (!h->window_fixed[r]) {
      if (h->path[r].empty()) return fail(h, MPPI_ERR_STATE, "no path or window set for robot " + std::to_string(r));
      if (!on_device) {  // x, y only: the yaw column is filled in when mppi_get_window asks for it
        h->cur_index[r] = calc_ref_path(h->path[r].data(), (int)(h->path[r].size() / 2), s[0], s[1], h->params.v_ref,
                                        dt, h->params.resolution, h->T, w, false);
        h->window_yaw_stale[r] = 1;
      }
    }
    // windows built on the device (K-1) overwrite this robot's slot after the H2D copy
    if (!on_device || h->window_fixed[r]) {
      window_to_robot_frame(w, h->T, s[0], s[1], win + (size_t)r * d.win_stride);
      h->host_windows_staged = true;
    }
    state_to_robot_frame(h->model, s, st + (size_t)r * kStateStride);
    st64[2 * r] = s[0];
    st64[2 * r + 1] = s[1];
  }
  if (u_nominal) {
    const size_t n = (size_t)h->R * d.planes;
    for (size_t k = 0; k < n; ++k) nom[k] = (float)u_nominal[k];
  }
  return MPPI_OK;
}

// Host -> device copy of the staged block: always header + poses + state records; the warm start only when the host
// supplies it; the windows only when they were built on the host (the device builder writes them in place).
int enqueue_h2d(mppi_handle h, bool with_nominal, cudaStream_t s) {
  const bool win = h->host_windows_staged;
  size_t bytes = 0;
  if (with_nominal) {
    bytes = win ? h->in_bytes : h->window_off;
    CU_TRY(h, cudaMemcpyAsync(h->d_in, h->h_in, bytes, cudaMemcpyHostToDevice, s));
  } else {
    bytes = h->nominal_off;
    CU_TRY(h, cudaMemcpyAsync(h->d_in, h->h_in, bytes, cudaMemcpyHostToDevice, s));
    if (win) {
      CU_TRY(h, cudaMemcpyAsync(h->d_in + h->window_off, h->h_in + h->window_off, h->in_bytes - h->window_off,
                                cudaMemcpyHostToDevice, s));
      bytes += h->in_bytes - h->window_off;
    }
  }
  h->last_h2d_bytes = bytes;
  return MPPI_OK;
}

// graph replay: unsharded handles and the NVLink peer exchange (ordinary kernels); the NCCL transport keeps plain launches
bool graph_capable(mppi_handle h) { return h->use_graph && (h->n_ranks == 1 || h->p2p); }

int capture(mppi_handle h, bool with_copies, cudaGraphExec_t *out) {
  cudaGraph_t g = nullptr;
  CU_TRY(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  int rc = MPPI_OK;
  if (with_copies) rc = enqueue_h2d(h, h->opt_upload_warm_start != 0, h->stream);
  if (rc == MPPI_OK) rc = issue_kernels(h, h->stream, true);
  if (rc == MPPI_OK && with_copies && !h->out_mapped) {
    cudaError_t e = cudaMemcpyAsync(h->h_out, h->d_out, h->out_bytes, cudaMemcpyDeviceToHost, h->stream);
    if (e != cudaSuccess) rc = fail(h, MPPI_ERR_CUDA, cudaGetErrorString(e));
  }
  cudaError_t e = cudaStreamEndCapture(h->stream, &g);
  if (rc != MPPI_OK) {
    if (g) cudaGraphDestroy(g);
    return rc;
  }
  CU_TRY(h, e);
  e = cudaGraphInstantiate(out, g, 0);
  cudaGraphDestroy(g);
  CU_TRY(h, e);
  return MPPI_OK;
}

// the merge of the peer exchange flags a record that did not arrive in stats[3] (h_out holds a fresh copy of d_out)
int exchange_status(mppi_handle h) {
  if (!h->p2p) return MPPI_OK;
  const size_t n = (size_t)h->R * h->d.planes;
  for (int r = 0; r < h->R; ++r)
    if (h->h_out[n + (size_t)r * 4 + 3] != 0.f)
      return fail(h, MPPI_ERR_NCCL,
                  "peer exchange timed out: a rank's record did not arrive (are all ranks solving?); controls and warm "
                  "start keep their previous values");
  return MPPI_OK;
}

int copy_out(mppi_handle h, double *u_nominal) {
  int rc = exchange_status(h);
  if (rc) return rc;
  const size_t n = (size_t)h->R * h->d.planes;
  for (size_t k = 0; k < n; ++k) u_nominal[k] = (double)h->h_out[k];
  return MPPI_OK;
}

}  // namespace

#ifndef MPPI_MAPPED_OUT_BYTES
#define MPPI_MAPPED_OUT_BYTES (1 << 20)
#endif

extern "C" {

int mppi_abi_version(void) { return MPPI_B200_ABI_VERSION; }

const char *mppi_last_error(mppi_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int mppi_create(mppi_handle *out, int model, const mppi_params *params, int num_samples, int horizon, int n_robots,
                int device) {
  if (!out) return fail(nullptr, MPPI_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (model < 0 || model > 2) return fail(nullptr, MPPI_ERR_INVALID, "unknown model");
  if (!params_valid(params)) return fail(nullptr, MPPI_ERR_INVALID, "params: need lambda > 0, resolution > 0, control_noise >= 0, u_min <= u_max");
  if (num_samples < 1 || horizon < 2 || horizon > 4096 || n_robots < 1 || n_robots > 65535)
    return fail(nullptr, MPPI_ERR_INVALID, "need num_samples >= 1, 2 <= horizon <= 4096, 1 <= n_robots <= 65535");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, MPPI_ERR_CUDA, std::string("no CUDA device (this library has no CPU fallback): ") +
                                            (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  if (device < 0 || device >= ndev) return fail(nullptr, MPPI_ERR_INVALID, "device ordinal out of range");
  mppi_handle h = new (std::nothrow) mppi_handle_s();
  if (!h) return fail(nullptr, MPPI_ERR_ALLOC, "out of host memory");
  h->model = model;
  h->K = num_samples;
  h->T = horizon;
  h->U = num_controls(model);
  h->S = num_states(model);
  h->R = n_robots;
  h->device = device;
  h->params = *params;
  h->k_global = num_samples;
  DeviceState &d = h->d;
  d.key0 = (uint32_t)h->seed;
  d.key1 = (uint32_t)(h->seed >> 32);
  d.model = model;
  d.T = horizon;
  d.U = h->U;
  d.K = num_samples;
  d.Kp = (int)align_up((size_t)num_samples, 4);
  d.R = n_robots;
  d.planes = (horizon - 1) * h->U;
  d.win_stride = (int)align_up((size_t)2 * horizon, 4);
  d.rec_stride = (int)align_up((size_t)4 + d.planes, 4);
  d.nb3 = (num_samples + kWeightBlock * 4 - 1) / (kWeightBlock * 4);
  {
    const int chunk = kReduceChunk / reduce_planes_per_block(d.Kp);
    d.nchunk = (d.Kp + chunk - 1) / chunk;
  }
  d.n_ranks = 1;
  // candidate grid of the pruned scan: building it costs ~cells*T distance evaluations per robot and solve, the
  // rollouts K*T*(~100 instr): keep the grid below a few per cent of that
  d.grid_max_cells = num_samples / 2 < 256 ? 256 : (num_samples / 2 > 65536 ? 65536 : num_samples / 2);
  if (d.planes > 65535) {
    delete h;
    return fail(nullptr, MPPI_ERR_INVALID, "(horizon-1)*U exceeds 65535");
  }

  auto bail = [&](int code, const std::string &msg) {
    std::string m = msg;
    mppi_destroy(h);
    return fail(nullptr, code, m);
  };
#define CU_NEW(expr)                                                                   \
  do {                                                                                 \
    cudaError_t e__ = (expr);                                                          \
    if (e__ != cudaSuccess) return bail(MPPI_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)
  CU_NEW(cudaSetDevice(device));
  // the solve's own kernels go first when they compete for SMs with the side stream's candidate grid / noise prefetch
  int prio_least = 0, prio_greatest = 0;
  CU_NEW(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  CU_NEW(cudaStreamCreateWithPriority(&h->own_stream, cudaStreamNonBlocking, prio_greatest));
  h->stream = h->own_stream;
  CU_NEW(cudaEventCreateWithFlags(&h->staged, cudaEventDisableTiming));
  CU_NEW(cudaStreamCreateWithPriority(&h->side_stream, cudaStreamNonBlocking, prio_least));
  CU_NEW(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CU_NEW(cudaEventCreateWithFlags(&h->ev_grid, cudaEventDisableTiming));
  CU_NEW(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  CU_NEW(cudaEventCreateWithFlags(&h->ev_join_cap, cudaEventDisableTiming));
  CU_NEW(cudaEventCreateWithFlags(&h->ev_readers[0], cudaEventDisableTiming));
  CU_NEW(cudaEventCreateWithFlags(&h->ev_readers[1], cudaEventDisableTiming));
  CU_NEW(cudaDeviceGetAttribute(&h->sm_clock_khz, cudaDevAttrClockRate, device));
  CU_NEW(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, device));
  if (h->sm_clock_khz <= 0) h->sm_clock_khz = 1965000;
  d.xchg_timeout_cycles = (long long)(h->opt_timeout_ms * (double)h->sm_clock_khz);

  // staging block: header | pose (FP64 x, y) | state record | warm start | windows -- a solve copies the prefix it needs
  const size_t win_bytes = sizeof(float) * (size_t)d.R * d.win_stride;
  const size_t st_bytes = sizeof(float) * (size_t)d.R * kStateStride;
  const size_t nom_bytes = sizeof(float) * (size_t)d.R * d.planes;
  h->state64_off = kHeaderBytes;
  h->state_off = h->state64_off + sizeof(double) * 2 * (size_t)d.R;
  h->nominal_off = align_up(h->state_off + st_bytes, 16);
  h->window_off = align_up(h->nominal_off + nom_bytes, 16);
  h->in_bytes = align_up(h->window_off + win_bytes, 16);
  h->out_bytes = sizeof(float) * ((size_t)d.R * d.planes + (size_t)d.R * 4);
  CU_NEW(cudaMallocHost((void **)&h->h_in, h->in_bytes));
  CU_NEW(cudaMallocHost((void **)&h->h_out, h->out_bytes));
  memset(h->h_in, 0, h->in_bytes);
  memset(h->h_out, 0, h->out_bytes);
  CU_NEW(cudaMalloc((void **)&h->d_in, h->in_bytes));
  CU_NEW(cudaMemset(h->d_in, 0, h->in_bytes));
  // Results of up to 1 MB ((T-1) x U controls + 4 statistics per robot; 1024 robots at T = 50: 418 KB) need no device
  // copy and no D2H node: the pinned block is addressable from the device (unified addressing), the tail kernel stores
  // into it over PCIe (posted writes, nothing on the device reads them back) and the host reads it after the stream
  // synchronisation it does anyway.  Larger fleets keep the device buffer + one copy.
  h->out_mapped = h->out_bytes <= (size_t)MPPI_MAPPED_OUT_BYTES;
  if (h->out_mapped) {
    h->d_out = h->h_out;
  } else {
    CU_NEW(cudaMalloc((void **)&h->d_out, h->out_bytes));
    CU_NEW(cudaMemset(h->d_out, 0, h->out_bytes));
  }
  d.hdr = reinterpret_cast<SolveHeader *>(h->d_in);
  d.window = reinterpret_cast<float *>(h->d_in + h->window_off);
  d.state = reinterpret_cast<float *>(h->d_in + h->state_off);
  d.nominal = reinterpret_cast<float *>(h->d_in + h->nominal_off);
  d.state64 = reinterpret_cast<double *>(h->d_in + h->state64_off);
  CU_NEW(cudaMalloc((void **)&h->d_path_off, sizeof(int) * ((size_t)d.R + 1)));
  CU_NEW(cudaMemset(h->d_path_off, 0, sizeof(int) * ((size_t)d.R + 1)));
  CU_NEW(cudaMalloc((void **)&h->d_win_fixed, (size_t)d.R));
  CU_NEW(cudaMemset(h->d_win_fixed, 0, (size_t)d.R));
  CU_NEW(cudaMalloc((void **)&d.cur_index, sizeof(int) * (size_t)d.R));
  CU_NEW(cudaMemset(d.cur_index, 0, sizeof(int) * (size_t)d.R));
  d.path_off = h->d_path_off;
  d.win_fixed = h->d_win_fixed;
  d.u_new = h->d_out;
  d.stats = h->d_out + (size_t)d.R * d.planes;
  // noise tensor: two buffers (the next solve's normals are generated beside the current solve) when both fit easily
  d.eps_buf_elems = (size_t)d.R * d.planes * d.Kp;
  {
    size_t free_b = 0, total_b = 0;
    CU_NEW(cudaMemGetInfo(&free_b, &total_b));
    d.eps_buffers = (2 * sizeof(float) * d.eps_buf_elems <= free_b / 2) ? 2 : 1;
  }
  CU_NEW(cudaMalloc((void **)&d.eps, sizeof(float) * d.eps_buffers * d.eps_buf_elems));
  CU_NEW(cudaMalloc((void **)&d.cost, sizeof(float) * (size_t)d.R * d.K));
  CU_NEW(cudaMalloc((void **)&d.weight, sizeof(float) * (size_t)d.R * d.K));
  const int nb3_weights = d.nb3;  // blocks of the stand-alone weight kernel (also the on-demand weights tap)
  if (fused_weights(h)) d.nb3 = d.nchunk;  // the partial (sum w, sum w^2) come from K4's chunks
  CU_NEW(cudaMalloc((void **)&d.wpart, sizeof(float) * (size_t)d.R * (nb3_weights > d.nb3 ? nb3_weights : d.nb3) * 2));
  // K4's per-chunk partial numerators [R][P][nchunk]; the one-kernel tail keeps its per-group partial records
  // [R][groups <= nchunk][rec_stride] here
  CU_NEW(cudaMalloc((void **)&d.npart, sizeof(float) * (size_t)d.R * d.rec_stride * d.nchunk));
  CU_NEW(cudaMalloc((void **)&d.record, sizeof(float) * (size_t)d.R * d.rec_stride));
  // per-CTA records of the fused weighted controls (fused_controls() decides per solve whether K2 writes them)
  CU_NEW(cudaMalloc((void **)&d.cta_part, sizeof(float) * (size_t)d.R * ((d.K + 127) / 128) * d.rec_stride));
  CU_NEW(cudaMalloc((void **)&d.tail_ticket, sizeof(unsigned int) * ((size_t)d.R + 1)));
  CU_NEW(cudaMemset(d.tail_ticket, 0, sizeof(unsigned int) * ((size_t)d.R + 1)));
  CU_NEW(cudaMalloc((void **)&d.cmin, sizeof(unsigned int) * (size_t)d.R));
  CU_NEW(cudaMemset(d.cmin, 0xFF, sizeof(unsigned int) * (size_t)d.R));  // armed; every solve's tail re-arms it
  // {solve counter, next solve the prefetching generator produces, the generator's last-block ticket, -}
  CU_NEW(cudaMalloc((void **)&d.counter, 4 * sizeof(uint32_t)));
  CU_NEW(cudaMemset(d.counter, 0, 4 * sizeof(uint32_t)));
  CU_NEW(cudaMemset(d.eps, 0, sizeof(float) * d.eps_buffers * d.eps_buf_elems));
  if (make_eps_tensor_map(d) != cudaSuccess) d.eps_map_valid = false;  // no TMA descriptor: the literal scan runs
  if (pruned_scan_supported(d.T, d.planes)) d.side_carveout = rollout_carveout_percent(d);
  {
    const SolveParams P0 = make_solve_params(model, horizon, *params, 0.1);
    for (int u = 0; u < kMaxControls; ++u) {
      d.bounds.lo[u] = P0.u_min[u];
      d.bounds.hi[u] = P0.u_max[u];
    }
  }
  // two slots of everything the side stream writes for the NEXT solve while K2 of the running one still reads its own:
  // candidate grid (header + cells) and the device-built windows; back-to-back solves alternate between them
  CU_NEW(cudaMalloc((void **)&d.grid_hdr, sizeof(GridHeader) * 2 * (size_t)d.R));
  CU_NEW(cudaMemset(d.grid_hdr, 0, sizeof(GridHeader) * 2 * (size_t)d.R));
  CU_NEW(cudaMalloc((void **)&d.grid_cells, sizeof(uint32_t) * 2 * (size_t)d.R * d.grid_max_cells));
  CU_NEW(cudaMalloc((void **)&h->d_window2, sizeof(float) * (size_t)d.R * d.win_stride));
  CU_NEW(cudaMemset(h->d_window2, 0, sizeof(float) * (size_t)d.R * d.win_stride));
  d.gathered = d.record;
#undef CU_NEW
  h->path.resize(n_robots);
  h->window_fixed.assign(n_robots, 0);
  h->window_yaw_stale.assign(n_robots, 0);
  h->window.assign((size_t)n_robots * horizon * 3, 0.0);
  h->cur_index.assign(n_robots, 0);
  h->last_state.assign((size_t)n_robots * h->S, 0.0);
  *out = h;
  return MPPI_OK;
}

int mppi_destroy(mppi_handle h) {
  if (!h) return MPPI_OK;
  cudaSetDevice(h->device);
  if (h->own_stream) cudaStreamSynchronize(h->own_stream);
  if (h->side_stream) cudaStreamSynchronize(h->side_stream);
  invalidate_graphs(h);
  if (h->comm && g_nccl.comm_destroy) g_nccl.comm_destroy(h->comm);
  for (size_t g = 0; g < h->xchg_peer_ptrs.size(); ++g)
    if (h->xchg_peer_ptrs[g] && h->xchg_peer_ptrs[g] != h->xchg_buf) cudaIpcCloseMemHandle(h->xchg_peer_ptrs[g]);
  cudaFree(h->xchg_buf); cudaFree(h->d_xchg_peers); cudaFree(h->d.xchg_seq); cudaFree(h->d.xchg_ticket);
  DeviceState &d = h->d;
  if (d.gathered && d.gathered != d.record) cudaFree(d.gathered);
  cudaFree(d.eps); cudaFree(d.cost); cudaFree(d.weight); cudaFree(d.wpart); cudaFree(d.npart);
  cudaFree(d.record); cudaFree(d.cta_part); cudaFree(d.tail_ticket); cudaFree(d.cmin); cudaFree(d.counter);
  cudaFree(d.nearest);
  cudaFree(d.grid_hdr); cudaFree(d.grid_cells); cudaFree(d.states_dbg); cudaFree(h->d_window2);
  cudaFree(h->d_path); cudaFree(h->d_path_off); cudaFree(h->d_win_fixed); cudaFree(d.cur_index);
  cudaFree(h->d_in);
  if (!h->out_mapped) cudaFree(h->d_out);
  if (h->h_in) cudaFreeHost(h->h_in);
  if (h->h_out) cudaFreeHost(h->h_out);
  if (h->staged) cudaEventDestroy(h->staged);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_grid) cudaEventDestroy(h->ev_grid);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->ev_join_cap) cudaEventDestroy(h->ev_join_cap);
  for (auto &e : h->ev_readers)
    if (e) cudaEventDestroy(e);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return MPPI_OK;
}

int mppi_set_params(mppi_handle h, const mppi_params *params) {
  if (!h) return MPPI_ERR_INVALID;
  if (!params_valid(params)) return fail(h, MPPI_ERR_INVALID, "params: need lambda > 0, resolution > 0, control_noise >= 0, u_min <= u_max");
  h->params = *params;
  return MPPI_OK;
}

int mppi_set_debug(mppi_handle h, int debug_flags) {
  if (!h) return MPPI_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  if ((debug_flags & MPPI_DEBUG_NEAREST) && !h->d.nearest) {
    const size_t bytes = sizeof(int) * (size_t)h->R * h->K * h->T;
    CU_TRY(h, cudaMalloc((void **)&h->d.nearest, bytes));
    CU_TRY(h, cudaMemset(h->d.nearest, 0xFF, bytes));
  }
  if ((debug_flags & MPPI_DEBUG_STATES) && !h->d.states_dbg) {
    const size_t bytes = sizeof(float) * (size_t)h->R * h->K * h->T * 5;
    CU_TRY(h, cudaMalloc((void **)&h->d.states_dbg, bytes));
    CU_TRY(h, cudaMemset(h->d.states_dbg, 0, bytes));
  }
  if (debug_flags != h->debug_flags) invalidate_graphs(h);
  h->debug_flags = debug_flags;
  return MPPI_OK;
}

int mppi_set_scan_mode(mppi_handle h, int scan_mode) {
  if (!h) return MPPI_ERR_INVALID;
  if (scan_mode < MPPI_SCAN_AUTO || scan_mode > MPPI_SCAN_PRUNED) return fail(h, MPPI_ERR_INVALID, "unknown scan mode");
  if (scan_mode != h->scan_mode) invalidate_graphs(h);
  h->scan_mode = scan_mode;
  return MPPI_OK;
}

int mppi_set_window_builder(mppi_handle h, int mode) {
  if (!h) return MPPI_ERR_INVALID;
  if (mode < MPPI_WINDOW_AUTO || mode > MPPI_WINDOW_DEVICE) return fail(h, MPPI_ERR_INVALID, "unknown window builder mode");
  if (mode != h->window_builder) invalidate_graphs(h);
  h->window_builder = mode;
  h->paths_dirty = true;
  return MPPI_OK;
}

int mppi_set_option(mppi_handle h, int option, double value) {
  if (!h) return MPPI_ERR_INVALID;
  if (!std::isfinite(value)) return fail(h, MPPI_ERR_INVALID, "option value must be finite");
  const long long iv = (long long)value;
  const bool integral = (double)iv == value;
  DeviceState &d = h->d;
  switch (option) {
    case MPPI_OPT_GRID_MAX_CELLS: {
      if (!integral || iv < 256 || iv > 262144) return fail(h, MPPI_ERR_INVALID, "MPPI_OPT_GRID_MAX_CELLS: 256 .. 262144");
      if ((int)iv == d.grid_max_cells) return MPPI_OK;
      CU_TRY(h, cudaSetDevice(h->device));
      CU_TRY(h, cudaStreamSynchronize(h->stream));
      uint32_t *cells = nullptr;
      if (h->side_stream) CU_TRY(h, cudaStreamSynchronize(h->side_stream));
      CU_TRY(h, cudaMalloc((void **)&cells, sizeof(uint32_t) * 2 * (size_t)d.R * (size_t)iv));
      cudaFree(d.grid_cells);
      d.grid_cells = cells;
      d.grid_max_cells = (int)iv;
      break;
    }
    case MPPI_OPT_GRID_H_MIN:
      if (!(value > 0.0)) return fail(h, MPPI_ERR_INVALID, "MPPI_OPT_GRID_H_MIN: > 0");
      d.grid_h_min = (float)value;
      break;
    case MPPI_OPT_GRID_MARGIN:
      if (value < 0.0) return fail(h, MPPI_ERR_INVALID, "MPPI_OPT_GRID_MARGIN: >= 0 (0 = automatic)");
      d.grid_margin = (float)value;
      break;
    case MPPI_OPT_GRID_LANES:
      if (!integral || !(iv == 0 || iv == 1 || iv == 2 || iv == 4 || iv == 8 || iv == 16 || iv == 32))
        return fail(h, MPPI_ERR_INVALID, "MPPI_OPT_GRID_LANES: 0 (automatic), 1, 2, 4, 8, 16 or 32");
      d.grid_lanes = (int)iv;
      break;
    case MPPI_OPT_REDUCE_GROUPS:
      if (!integral || iv < 0 || iv > 65535) return fail(h, MPPI_ERR_INVALID, "MPPI_OPT_REDUCE_GROUPS: 0 (automatic) .. 65535");
      d.k4_groups = (int)iv;
      break;
    case MPPI_OPT_FUSE_CONTROLS:
      if (!integral || iv < -1 || iv > 1) return fail(h, MPPI_ERR_INVALID, "MPPI_OPT_FUSE_CONTROLS: -1 (automatic), 0 or 1");
      h->opt_fuse_controls = (int)iv;
      break;
    case MPPI_OPT_NOISE_PREFETCH:
      if (!integral || iv < -1 || iv > 1) return fail(h, MPPI_ERR_INVALID, "MPPI_OPT_NOISE_PREFETCH: -1 (automatic), 0 or 1");
      if (iv == 1 && d.eps_buffers < 2)
        return fail(h, MPPI_ERR_INVALID, "MPPI_OPT_NOISE_PREFETCH: the second noise buffer did not fit at mppi_create");
      h->opt_prefetch = (int)iv;
      h->noise_primed = false;
      break;
    case MPPI_OPT_EXCHANGE_TIMEOUT_MS:
      if (!(value >= 1.0 && value <= 600000.0)) return fail(h, MPPI_ERR_INVALID, "MPPI_OPT_EXCHANGE_TIMEOUT_MS: 1 .. 600000");
      h->opt_timeout_ms = value;
      d.xchg_timeout_cycles = (long long)(value * (double)h->sm_clock_khz);
      break;
    case MPPI_OPT_FEEDBACK_WARM_START:
      if (!integral || iv < 0 || iv > 1) return fail(h, MPPI_ERR_INVALID, "MPPI_OPT_FEEDBACK_WARM_START: 0 or 1");
      d.feedback = iv != 0;
      break;
    case MPPI_OPT_UPLOAD_WARM_START:
      if (!integral || iv < 0 || iv > 1) return fail(h, MPPI_ERR_INVALID, "MPPI_OPT_UPLOAD_WARM_START: 0 or 1");
      h->opt_upload_warm_start = (int)iv;
      break;
    default:
      return fail(h, MPPI_ERR_INVALID, "unknown option");
  }
  invalidate_graphs(h);  // every option is a kernel argument or changes the kernel sequence
  return MPPI_OK;
}

int mppi_get_option(mppi_handle h, int option, double *value) {
  if (!h) return MPPI_ERR_INVALID;
  if (!value) return fail(h, MPPI_ERR_INVALID, "value is NULL");
  const DeviceState &d = h->d;
  switch (option) {
    case MPPI_OPT_GRID_MAX_CELLS: *value = d.grid_max_cells; break;
    case MPPI_OPT_GRID_H_MIN: *value = d.grid_h_min; break;
    case MPPI_OPT_GRID_MARGIN: *value = d.grid_margin; break;
    case MPPI_OPT_GRID_LANES: *value = d.grid_lanes; break;
    case MPPI_OPT_REDUCE_GROUPS: *value = d.k4_groups; break;
    case MPPI_OPT_FUSE_CONTROLS: *value = h->opt_fuse_controls; break;
    case MPPI_OPT_NOISE_PREFETCH: *value = h->opt_prefetch; break;
    case MPPI_OPT_EXCHANGE_TIMEOUT_MS: *value = h->opt_timeout_ms; break;
    case MPPI_OPT_FEEDBACK_WARM_START: *value = d.feedback ? 1.0 : 0.0; break;
    case MPPI_OPT_UPLOAD_WARM_START: *value = h->opt_upload_warm_start; break;
    case MPPI_INFO_FUSED_CONTROLS: *value = h->last_fused ? 1.0 : 0.0; break;
    default: return fail(h, MPPI_ERR_INVALID, "unknown option");
  }
  return MPPI_OK;
}

int mppi_set_path(mppi_handle h, int robot, const double *path_xy, int n_points) {
  if (!h) return MPPI_ERR_INVALID;
  if (robot < 0 || robot >= h->R || n_points < 1 || !path_xy) return fail(h, MPPI_ERR_INVALID, "bad robot index or empty path");
  h->path[robot].assign(path_xy, path_xy + (size_t)2 * n_points);
  h->window_fixed[robot] = 0;
  h->paths_dirty = true;
  return MPPI_OK;
}

int mppi_set_window(mppi_handle h, int robot, const double *window_xyyaw) {
  if (!h) return MPPI_ERR_INVALID;
  if (robot < 0 || robot >= h->R || !window_xyyaw) return fail(h, MPPI_ERR_INVALID, "bad robot index or NULL window");
  memcpy(h->window.data() + (size_t)robot * h->T * 3, window_xyyaw, sizeof(double) * 3 * (size_t)h->T);
  h->window_fixed[robot] = 1;
  h->window_yaw_stale[robot] = 0;
  h->cur_index[robot] = 0;
  h->paths_dirty = true;
  return MPPI_OK;
}

int mppi_set_seed(mppi_handle h, uint64_t seed, uint64_t first_solve_counter) {
  if (!h) return MPPI_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  if (seed != h->seed) invalidate_graphs(h);  // the key is a kernel argument of the captured noise node
  h->seed = seed;
  h->d.key0 = (uint32_t)seed;
  h->d.key1 = (uint32_t)(seed >> 32);
  uint32_t c = (uint32_t)first_solve_counter;
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  CU_TRY(h, cudaMemcpy(h->d.counter, &c, sizeof c, cudaMemcpyHostToDevice));
  h->noise_primed = false;
  return MPPI_OK;
}

int mppi_set_shard(mppi_handle h, int64_t sample_offset, int64_t num_samples_global, int robot_offset) {
  if (!h) return MPPI_ERR_INVALID;
  if (sample_offset < 0 || (sample_offset % 4) != 0 || num_samples_global < sample_offset + h->K || robot_offset < 0)
    return fail(h, MPPI_ERR_INVALID, "sample_offset must be a non-negative multiple of 4 inside the global sample range");
  h->sample_offset = sample_offset;
  h->k_global = num_samples_global;
  h->robot_offset = robot_offset;
  h->noise_primed = false;
  return MPPI_OK;
}

int mppi_set_noise(mppi_handle h, const float *eps) {
  if (!h) return MPPI_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  const bool ext = eps != nullptr;
  if (ext != h->external_noise) invalidate_graphs(h);
  h->external_noise = ext;
  h->noise_primed = false;
  if (!ext) return MPPI_OK;
  const DeviceState &d = h->d;
  const int K = h->K, U = h->U, steps = h->T - 1;
  std::vector<float> phys(d.eps_buf_elems, 0.f);
  for (int r = 0; r < d.R; ++r)
    for (int t = 0; t < steps; ++t)
      for (int i = 0; i < K; ++i)
        for (int u = 0; u < U; ++u)
          phys[((size_t)r * d.planes + (size_t)t * U + u) * d.Kp + i] = eps[(((size_t)r * steps + t) * K + i) * U + u];
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  for (int b = 0; b < d.eps_buffers; ++b)  // the same tensor for every following solve, whichever buffer it reads
    CU_TRY(h, cudaMemcpy(d.eps + (size_t)b * d.eps_buf_elems, phys.data(), sizeof(float) * phys.size(),
                         cudaMemcpyHostToDevice));
  return MPPI_OK;
}

int mppi_set_stream(mppi_handle h, void *cuda_stream) {
  if (!h) return MPPI_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  if (h->side_stream) CU_TRY(h, cudaStreamSynchronize(h->side_stream));
  invalidate_graphs(h);
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return MPPI_OK;
}

int mppi_use_graph(mppi_handle h, int enable) {
  if (!h) return MPPI_ERR_INVALID;
  h->use_graph = enable != 0;
  if (!enable) invalidate_graphs(h);
  return MPPI_OK;
}

int mppi_upload(mppi_handle h, const double *state, double dt, const double *u_nominal) {
  if (!h) return MPPI_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  int rc = stage_inputs(h, state, dt, u_nominal);
  if (rc) return rc;
  rc = enqueue_h2d(h, u_nominal != nullptr, h->stream);  // u_nominal == NULL keeps the device-resident warm start
  if (rc) return rc;
  CU_TRY(h, cudaEventRecord(h->staged, h->stream));
  h->staged_pending = true;
  h->staged_recorded = true;
  h->have_inputs = true;
  if (u_nominal) h->have_nominal = true;
  return MPPI_OK;
}

static int issue_kernels_primed(mppi_handle h) {
  int rc = prime_noise(h, false);
  if (rc) return rc;
  return issue_kernels(h, h->stream, false);
}

int mppi_enqueue(mppi_handle h) {
  if (!h) return MPPI_ERR_INVALID;
  if (!h->have_inputs) return fail(h, MPPI_ERR_STATE, "mppi_enqueue before mppi_upload");
  CU_TRY(h, cudaSetDevice(h->device));
  int rc = prime_noise(h, false);
  if (rc) return rc;
  if (graph_capable(h)) {
    if (!h->exec_kernels) {
      rc = capture(h, false, &h->exec_kernels);
      if (rc) return rc;
    }
    if (h->join_recorded) CU_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));  // a generator from a stream-launched solve
    CU_TRY(h, cudaGraphLaunch(h->exec_kernels, h->stream));
    h->weights_valid = !h->last_fused;  // a replay overwrites the costs: weights of an earlier solve are stale
    return MPPI_OK;
  }
  return issue_kernels(h, h->stream, false);
}

int mppi_download(mppi_handle h, double *u_nominal) {
  if (!h) return MPPI_ERR_INVALID;
  if (!u_nominal) return fail(h, MPPI_ERR_INVALID, "u_nominal is NULL");
  CU_TRY(h, cudaSetDevice(h->device));
  if (!h->out_mapped) CU_TRY(h, cudaMemcpyAsync(h->h_out, h->d_out, h->out_bytes, cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  h->staged_pending = false;
  return copy_out(h, u_nominal);
}

int mppi_synchronize(mppi_handle h) {
  if (!h) return MPPI_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  if (h->p2p && !h->out_mapped) {  // bring the statistics over: a peer-exchange time-out of an enqueued solve is reported here too
    const size_t off = sizeof(float) * (size_t)h->R * h->d.planes;
    CU_TRY(h, cudaMemcpyAsync((char *)h->h_out + off, (char *)h->d_out + off, h->out_bytes - off, cudaMemcpyDeviceToHost,
                              h->stream));
  }
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  h->staged_pending = false;
  return exchange_status(h);
}

int mppi_solve(mppi_handle h, const double *state, double dt, double *u_nominal) {
  if (!h) return MPPI_ERR_INVALID;
  if (!u_nominal) return fail(h, MPPI_ERR_INVALID, "u_nominal is NULL");
  CU_TRY(h, cudaSetDevice(h->device));
  // MPPI_OPT_UPLOAD_WARM_START = 0: after the first solve the warm start is the device's own copy of the previous
  // result (what the reference's optimal_solution is when the host does not touch it between cycles)
  const bool with_nominal = h->opt_upload_warm_start != 0 || !h->have_nominal;
  if (graph_capable(h) && with_nominal == (h->opt_upload_warm_start != 0)) {
    int rc = stage_inputs(h, state, dt, with_nominal ? u_nominal : nullptr);
    if (rc) return rc;
    rc = prime_noise(h, true);
    if (rc) return rc;
    if (!h->exec_solve) {
      rc = capture(h, true, &h->exec_solve);
      if (rc) return rc;
    }
    if (h->join_recorded) CU_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    CU_TRY(h, cudaGraphLaunch(h->exec_solve, h->stream));
    h->weights_valid = !h->last_fused;
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    h->have_inputs = h->have_nominal = true;
    return copy_out(h, u_nominal);
  }
  int rc = mppi_upload(h, state, dt, with_nominal ? u_nominal : nullptr);
  if (rc) return rc;
  if (graph_capable(h)) rc = issue_kernels_primed(h);  // a one-off solve in front of the graph: plain launches
  else rc = mppi_enqueue(h);
  if (rc) return rc;
  return mppi_download(h, u_nominal);
}

int mppi_get_costs(mppi_handle h, int robot, float *cost) {
  if (!h) return MPPI_ERR_INVALID;
  if (robot < 0 || robot >= h->R || !cost) return fail(h, MPPI_ERR_INVALID, "bad robot index or NULL buffer");
  CU_TRY(h, cudaSetDevice(h->device));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  CU_TRY(h, cudaMemcpy(cost, h->d.cost + (size_t)robot * h->K, sizeof(float) * (size_t)h->K, cudaMemcpyDeviceToHost));
  return MPPI_OK;
}

int mppi_get_weights(mppi_handle h, int robot, float *weights) {
  if (!h) return MPPI_ERR_INVALID;
  if (robot < 0 || robot >= h->R || !weights) return fail(h, MPPI_ERR_INVALID, "bad robot index or NULL buffer");
  CU_TRY(h, cudaSetDevice(h->device));
  if (!h->weights_valid) {  // fused-controls path: exp(-(c - c_min)/lambda) from the resident costs, on demand
    DeviceState t = h->d;
    t.nb3 = (h->K + kWeightBlock * 4 - 1) / (kWeightBlock * 4);
    CU_TRY(h, launch_weights(t, true, h->stream));
    h->weights_valid = true;
  }
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  CU_TRY(h, cudaMemcpy(weights, h->d.weight + (size_t)robot * h->K, sizeof(float) * (size_t)h->K, cudaMemcpyDeviceToHost));
  return MPPI_OK;
}

int mppi_get_nearest(mppi_handle h, int robot, int32_t *nearest) {
  if (!h) return MPPI_ERR_INVALID;
  if (robot < 0 || robot >= h->R || !nearest) return fail(h, MPPI_ERR_INVALID, "bad robot index or NULL buffer");
  if (!(h->debug_flags & MPPI_DEBUG_NEAREST) || !h->d.nearest)
    return fail(h, MPPI_ERR_STATE, "mppi_get_nearest needs mppi_set_debug(MPPI_DEBUG_NEAREST) before the solve");
  CU_TRY(h, cudaSetDevice(h->device));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  const size_t n = (size_t)h->K * h->T;
  CU_TRY(h, cudaMemcpy(nearest, h->d.nearest + (size_t)robot * n, sizeof(int) * n, cudaMemcpyDeviceToHost));
  return MPPI_OK;
}

int mppi_get_states(mppi_handle h, int robot, double *states) {
  if (!h) return MPPI_ERR_INVALID;
  if (robot < 0 || robot >= h->R || !states) return fail(h, MPPI_ERR_INVALID, "bad robot index or NULL buffer");
  if (!(h->debug_flags & MPPI_DEBUG_STATES) || !h->d.states_dbg)
    return fail(h, MPPI_ERR_STATE, "mppi_get_states needs mppi_set_debug(MPPI_DEBUG_STATES) before the solve");
  CU_TRY(h, cudaSetDevice(h->device));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  const size_t n = (size_t)h->K * h->T;
  std::vector<float> raw(n * 5);
  CU_TRY(h, cudaMemcpy(raw.data(), h->d.states_dbg + (size_t)robot * n * 5, sizeof(float) * raw.size(), cudaMemcpyDeviceToHost));
  // back to the world frame: the kernels work robot-centred
  const double px = h->last_state[(size_t)robot * h->S], py = h->last_state[(size_t)robot * h->S + 1];
  for (size_t k = 0; k < n; ++k) {
    double *o = states + k * h->S;
    o[0] = (double)raw[5 * k] + px;
    o[1] = (double)raw[5 * k + 1] + py;
    o[2] = raw[5 * k + 2];
    if (h->S == 5) {
      o[3] = raw[5 * k + 3];
      o[4] = raw[5 * k + 4];
    }
  }
  return MPPI_OK;
}

int mppi_get_noise(mppi_handle h, int robot, float *eps) {
  if (!h) return MPPI_ERR_INVALID;
  if (robot < 0 || robot >= h->R || !eps) return fail(h, MPPI_ERR_INVALID, "bad robot index or NULL buffer");
  CU_TRY(h, cudaSetDevice(h->device));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  const DeviceState &d = h->d;
  std::vector<float> phys((size_t)d.planes * d.Kp);
  uint32_t counter = 0;  // the merge has advanced it: the last solve used buffer (counter - 1) & (buffers - 1)
  CU_TRY(h, cudaMemcpy(&counter, d.counter, sizeof counter, cudaMemcpyDeviceToHost));
  const size_t buf = (size_t)((counter - 1u) & (uint32_t)(d.eps_buffers - 1));
  CU_TRY(h, cudaMemcpy(phys.data(), d.eps + buf * d.eps_buf_elems + (size_t)robot * d.planes * d.Kp,
                       sizeof(float) * phys.size(), cudaMemcpyDeviceToHost));
  const int K = h->K, U = h->U, steps = h->T - 1;
  for (int t = 0; t < steps; ++t)
    for (int i = 0; i < K; ++i)
      for (int u = 0; u < U; ++u) eps[((size_t)t * K + i) * U + u] = phys[((size_t)t * U + u) * d.Kp + i];
  return MPPI_OK;
}

int mppi_get_window(mppi_handle h, int robot, double *window_xyyaw, int *current_index) {
  if (!h) return MPPI_ERR_INVALID;
  if (robot < 0 || robot >= h->R) return fail(h, MPPI_ERR_INVALID, "bad robot index");
  if (device_windows(h) && !h->window_fixed[robot] && !h->path[robot].empty() && h->last_dt > 0.0) {
    // built on the device: current_index_ comes back from there, the FP64 window is re-derived from it with the
    // host form of calc_RefPath (same expressions)
    CU_TRY(h, cudaSetDevice(h->device));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    int cur = 0;
    CU_TRY(h, cudaMemcpy(&cur, h->d.cur_index + robot, sizeof(int), cudaMemcpyDeviceToHost));
    h->cur_index[robot] = cur;
    window_from_index(h->path[robot].data(), (int)(h->path[robot].size() / 2), cur, h->params.v_ref, h->last_dt,
                      h->params.resolution, h->T, h->window.data() + (size_t)robot * h->T * 3);
  }
  if (h->window_yaw_stale[robot]) {
    window_yaw(h->T, h->window.data() + (size_t)robot * h->T * 3);
    h->window_yaw_stale[robot] = 0;
  }
  if (window_xyyaw) memcpy(window_xyyaw, h->window.data() + (size_t)robot * h->T * 3, sizeof(double) * 3 * (size_t)h->T);
  if (current_index) *current_index = h->cur_index[robot];
  return MPPI_OK;
}

int mppi_get_stats(mppi_handle h, int robot, double *stats) {
  if (!h) return MPPI_ERR_INVALID;
  if (robot < 0 || robot >= h->R || !stats) return fail(h, MPPI_ERR_INVALID, "bad robot index or NULL buffer");
  CU_TRY(h, cudaSetDevice(h->device));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  float s[4];
  CU_TRY(h, cudaMemcpy(s, h->d.stats + (size_t)robot * 4, sizeof s, cudaMemcpyDefault));  // device or mapped host
  stats[0] = s[0];
  stats[1] = s[1];
  stats[2] = s[2];
  return MPPI_OK;
}

int mppi_get_record(mppi_handle h, int robot, float *record) {
  if (!h) return MPPI_ERR_INVALID;
  if (robot < 0 || robot >= h->R || !record) return fail(h, MPPI_ERR_INVALID, "bad robot index or NULL buffer");
  CU_TRY(h, cudaSetDevice(h->device));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  CU_TRY(h, cudaMemcpy(record, h->d.record + (size_t)robot * h->d.rec_stride, sizeof(float) * (4 + (size_t)h->d.planes),
                       cudaMemcpyDeviceToHost));
  return MPPI_OK;
}

int mppi_get_info(mppi_handle h, int *model, int *num_samples, int *horizon, int *num_controls_out, int *n_robots) {
  if (!h) return MPPI_ERR_INVALID;
  if (model) *model = h->model;
  if (num_samples) *num_samples = h->K;
  if (horizon) *horizon = h->T;
  if (num_controls_out) *num_controls_out = h->U;
  if (n_robots) *n_robots = h->R;
  return MPPI_OK;
}

int mppi_get_io_bytes(mppi_handle h, size_t *h2d_bytes, size_t *d2h_bytes) {
  if (!h) return MPPI_ERR_INVALID;
  if (h2d_bytes) *h2d_bytes = h->last_h2d_bytes ? h->last_h2d_bytes : h->in_bytes;
  if (d2h_bytes) *d2h_bytes = h->out_bytes;
  return MPPI_OK;
}

int mppi_time_kernels(mppi_handle h, int n_iters, float *ms) {
  if (!h) return MPPI_ERR_INVALID;
  if (n_iters < 1 || !ms) return fail(h, MPPI_ERR_INVALID, "n_iters >= 1 and ms != NULL required");
  if (!h->have_inputs) return fail(h, MPPI_ERR_STATE, "mppi_time_kernels before mppi_upload");
  CU_TRY(h, cudaSetDevice(h->device));
  cudaEvent_t ev[9] = {};
  struct EventGuard {  // destroyed on every exit path
    cudaEvent_t *e;
    ~EventGuard() {
      for (int k = 0; k < 9; ++k)
        if (e[k]) cudaEventDestroy(e[k]);
    }
  } guard{ev};
  for (auto &e : ev) CU_TRY(h, cudaEventCreate(&e));
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const DeviceState &d = h->d;
  cudaStream_t s = h->stream;
  bool want_nearest;
  const int scan = effective_scan(h, &want_nearest);
  h->noise_primed = false;  // the generator runs in front of its own solve here
  if (h->join_recorded) CU_TRY(h, cudaStreamWaitEvent(s, h->ev_join, 0));  // a prefetching generator still in flight
  for (int it = 0; it <= n_iters; ++it) {
    if (device_windows(h)) CU_TRY(h, launch_window_builder(d, s));
    CU_TRY(h, cudaEventRecord(ev[0], s));
    if (!h->external_noise) CU_TRY(h, launch_noise(d, 0, s));
    CU_TRY(h, cudaEventRecord(ev[1], s));
    if (scan == MPPI_SCAN_PRUNED) CU_TRY(h, launch_candidate_grid(d, s));
    CU_TRY(h, cudaEventRecord(ev[7], s));
    const bool fc = fused_controls(h, scan);
    CU_TRY(h, launch_rollout_cost(d, scan, want_nearest, false, fc, s));
    CU_TRY(h, cudaEventRecord(ev[2], s));
    h->last_fused = fc;
    h->weights_valid = !fc;
    if (!fc && !fused_weights(h)) CU_TRY(h, launch_weights(d, false, s));
    CU_TRY(h, cudaEventRecord(ev[3], s));
    if (!fc) CU_TRY(h, launch_weighted_controls(d, fused_weights(h), s));
    CU_TRY(h, cudaEventRecord(ev[4], s));
    // tail: reported as "finalize" (+ "merge" where that is a launch of its own)
    bool merge_follows = false;
    if (fc) {
      const int mode = h->p2p ? 1 : (h->n_ranks > 1 ? 2 : 0);
      CU_TRY(h, launch_rescale_tail(d, mode, s));
      merge_follows = mode == 2;
    } else if (h->p2p) {
      CU_TRY(h, launch_finalize_exchange(d, s));
    } else if (fused_tail_for(h, d)) {
      CU_TRY(h, launch_finalize_merge(d, s));
    } else {
      CU_TRY(h, launch_finalize(d, s));
      merge_follows = true;
    }
    CU_TRY(h, cudaEventRecord(ev[5], s));
    if (merge_follows) {
      if (h->n_ranks > 1 &&
          g_nccl.all_gather(d.record, d.gathered, (size_t)d.R * d.rec_stride, kNcclFloat, h->comm, s) != 0)
        return fail(h, MPPI_ERR_NCCL, "ncclAllGather failed");
      CU_TRY(h, launch_merge(d, s));
    }
    CU_TRY(h, cudaEventRecord(ev[6], s));
    CU_TRY(h, cudaStreamSynchronize(s));
    if (it == 0) continue;  // warm-up
    for (int k = 0; k < 6; ++k) {
      float t = 0.f;
      CU_TRY(h, cudaEventElapsedTime(&t, k == 1 ? ev[7] : ev[k], ev[k + 1]));
      acc[k] += t;
    }
    float t = 0.f;
    CU_TRY(h, cudaEventElapsedTime(&t, ev[0], ev[6]));
    acc[6] += t;
    CU_TRY(h, cudaEventElapsedTime(&t, ev[1], ev[7]));
    acc[7] += t;
  }
  for (int k = 0; k < 8; ++k) ms[k] = (float)(acc[k] / n_iters);
  return MPPI_OK;
}

int mppi_last_launch_count(mppi_handle h) { return h ? h->launch_count : 0; }

int mppi_comm_get_unique_id(void *id_out) {
  if (!id_out) return MPPI_ERR_INVALID;
  if (!g_nccl.load()) return fail(nullptr, MPPI_ERR_NCCL, g_nccl.err);
  NcclUniqueId id;
  int rc = g_nccl.get_unique_id(&id);
  if (rc != 0) return fail(nullptr, MPPI_ERR_NCCL, "ncclGetUniqueId failed");
  memcpy(id_out, &id, sizeof id);
  return MPPI_OK;
}

int mppi_comm_init(mppi_handle h, const void *id, int rank, int n_ranks) {
  if (!h) return MPPI_ERR_INVALID;
  if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(h, MPPI_ERR_INVALID, "bad rank / n_ranks / id");
  if (h->comm) return fail(h, MPPI_ERR_STATE, "communicator already initialised");
  CU_TRY(h, cudaSetDevice(h->device));
  if (n_ranks == 1) return MPPI_OK;
  if (!g_nccl.load()) return fail(h, MPPI_ERR_NCCL, g_nccl.err);
  NcclUniqueId uid;
  memcpy(&uid, id, sizeof uid);
  int rc = g_nccl.comm_init_rank(&h->comm, n_ranks, uid, rank);
  if (rc != 0)
    return fail(h, MPPI_ERR_NCCL,
                std::string("ncclCommInitRank: ") + (g_nccl.get_error_string ? g_nccl.get_error_string(rc) : "?"));
  CU_TRY(h, cudaMalloc((void **)&h->d.gathered, sizeof(float) * (size_t)n_ranks * h->d.R * h->d.rec_stride));
  h->d.n_ranks = n_ranks;
  h->n_ranks = n_ranks;
  h->rank = rank;
  invalidate_graphs(h);
  return MPPI_OK;
}

int mppi_comm_export(mppi_handle h, int n_ranks, void *handle_out) {
  if (!h) return MPPI_ERR_INVALID;
  if (!handle_out || n_ranks < 2 || n_ranks > 32) return fail(h, MPPI_ERR_INVALID, "need 2 <= n_ranks <= 32 and a handle buffer");
  if (h->comm || h->xchg_buf) return fail(h, MPPI_ERR_STATE, "communicator already initialised");
  CU_TRY(h, cudaSetDevice(h->device));
  const DeviceState &d = h->d;
  // slots[2 parities][n_ranks][R][rec_stride] of 8-byte words {payload, stamp} (mppi_device.cuh)
  const size_t bytes = kExchangeHeaderBytes + sizeof(unsigned long long) * 2 * (size_t)n_ranks * d.R * d.rec_stride;
  CU_TRY(h, cudaMalloc(&h->xchg_buf, bytes));
  CU_TRY(h, cudaMemset(h->xchg_buf, 0, bytes));
  CU_TRY(h, cudaMalloc((void **)&h->d.xchg_seq, sizeof(unsigned int)));
  CU_TRY(h, cudaMemset(h->d.xchg_seq, 0, sizeof(unsigned int)));
  CU_TRY(h, cudaMalloc((void **)&h->d.xchg_ticket, 2 * sizeof(unsigned int)));
  CU_TRY(h, cudaMemset(h->d.xchg_ticket, 0, 2 * sizeof(unsigned int)));
  CU_TRY(h, cudaDeviceSynchronize());
  cudaIpcMemHandle_t ipc;
  CU_TRY(h, cudaIpcGetMemHandle(&ipc, h->xchg_buf));
  static_assert(sizeof(cudaIpcMemHandle_t) == MPPI_IPC_HANDLE_BYTES, "IPC handle size");
  memcpy(handle_out, &ipc, sizeof ipc);
  h->xchg_ranks = n_ranks;
  return MPPI_OK;
}

int mppi_comm_connect(mppi_handle h, const void *handles, int rank, int n_ranks) {
  if (!h) return MPPI_ERR_INVALID;
  if (!handles || !h->xchg_buf || n_ranks != h->xchg_ranks || rank < 0 || rank >= n_ranks)
    return fail(h, MPPI_ERR_INVALID, "mppi_comm_connect needs the handles of all ranks of the preceding mppi_comm_export");
  CU_TRY(h, cudaSetDevice(h->device));
  h->xchg_peer_ptrs.assign(n_ranks, nullptr);
  for (int g = 0; g < n_ranks; ++g) {
    if (g == rank) {
      h->xchg_peer_ptrs[g] = h->xchg_buf;
      continue;
    }
    cudaIpcMemHandle_t ipc;
    memcpy(&ipc, (const char *)handles + (size_t)g * MPPI_IPC_HANDLE_BYTES, sizeof ipc);
    CU_TRY(h, cudaIpcOpenMemHandle(&h->xchg_peer_ptrs[g], ipc, cudaIpcMemLazyEnablePeerAccess));
  }
  CU_TRY(h, cudaMalloc((void **)&h->d_xchg_peers, sizeof(void *) * n_ranks));
  CU_TRY(h, cudaMemcpy(h->d_xchg_peers, h->xchg_peer_ptrs.data(), sizeof(void *) * n_ranks, cudaMemcpyHostToDevice));
  h->d.xchg_buf = h->xchg_buf;
  h->d.xchg_peers = h->d_xchg_peers;
  h->d.xchg_rank = rank;
  h->d.n_ranks = n_ranks;
  h->n_ranks = n_ranks;
  h->rank = rank;
  h->p2p = true;
  invalidate_graphs(h);
  return MPPI_OK;
}

int mppi_merge_partials(const float *partials, int n_ranks, int n, double lambda, float *u_out, double *stats) {
  if (!partials || n_ranks < 1 || n < 0 || !(lambda > 0.0) || (n > 0 && !u_out)) return MPPI_ERR_INVALID;
  // same record layout as the device: {m, S, Q, -, N[n]} with stride 4 + n
  const size_t stride = (size_t)4 + n;
  const float inv_lambda = (float)(1.0 / lambda);
  const float m = merge_min(partials, n_ranks, stride);
  float S, Q;
  merge_sums(partials, n_ranks, stride, m, inv_lambda, S, Q);
  for (int p = 0; p < n; ++p) u_out[p] = merge_numerator(partials, n_ranks, stride, m, inv_lambda, p) / S;
  if (stats) {
    stats[0] = m;
    stats[1] = S;
    stats[2] = (double)S * S / Q;
  }
  return MPPI_OK;
}

int mppi_calc_ref_path(const double *path_xy, int n_points, double px, double py, double v_ref, double dt,
                       double resolution, int horizon, double *window_xyyaw, int *current_index_out) {
  if (!path_xy || n_points < 1 || horizon < 1 || !window_xyyaw || !(resolution > 0.0)) return MPPI_ERR_INVALID;
  int cur = calc_ref_path(path_xy, n_points, px, py, v_ref, dt, resolution, horizon, window_xyyaw);
  if (current_index_out) *current_index_out = cur;
  return MPPI_OK;
}

void mppi_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]) {
  Philox4 r = philox4x32_10(counter[0], counter[1], counter[2], counter[3], key[0], key[1]);
  for (int k = 0; k < 4; ++k) out[k] = r.v[k];
}

}  // extern "C"
