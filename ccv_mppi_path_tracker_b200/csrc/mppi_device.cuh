// mppi_device.cuh -- device-side helpers shared by the kernel translation units (warp/block reductions,
// parameter staging).  Internal; not part of the ABI.
#pragma once
#include "mppi_kernels.h"

namespace mppi {

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialisation attribute may start
// while its predecessor in the stream still runs; pdl_wait() returns once that predecessor has completed and its
// writes are visible (a no-op for an ordinary launch), pdl_trigger() lets the successor of THIS kernel start early.
// Used along the chain K0 -> K2 -> K3 -> K4: the successor's launch latency and input prologue overlap the predecessor.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-level min of the per-thread costs, one atomicMin per block; returns the block's minimum to every thread
// (INFINITY when no thread is valid)
__device__ __forceinline__ float block_min_to_global(float c, bool valid, unsigned int *cmin_slot, float *s_red) {
  float v = valid ? c : INFINITY;
  v = warp_min(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) s_red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float m = lane < nw ? s_red[lane] : INFINITY;
  m = warp_min(m);
  if (wid == 0 && lane == 0 && m < INFINITY) atomicMin(cmin_slot, float_to_ordered(m));
  return m;
}

// The noise tensor of the solve that is running: buffer (solve counter & pmask) of the double-buffered tensor
// (pmask = eps_buffers - 1; 0 = single buffer).  The counter is advanced by the merge at the end of a solve.
__device__ __forceinline__ uint32_t eps_parity(const uint32_t *__restrict__ counter, uint32_t pmask) {
  return pmask ? (*counter & pmask) : 0u;
}

__device__ __forceinline__ void load_params_to_shared(SolveParams *dst, const SolveHeader *hdr) {
  const uint32_t *src = reinterpret_cast<const uint32_t *>(&hdr->P);
  uint32_t *d = reinterpret_cast<uint32_t *>(dst);
  for (int k = threadIdx.x; k < (int)(sizeof(SolveParams) / 4); k += blockDim.x) d[k] = src[k];
}

// True (for every thread of the block) in the block that finishes last among `total` blocks sharing `ticket`; the
// ticket is reset for the next launch.  Everything the other blocks wrote before calling this is visible to the
// last block afterwards.  All threads of the block must call it.
__device__ __forceinline__ bool last_block_of_grid(unsigned int *ticket, unsigned int total) {
  __shared__ bool s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == total - 1u;
    if (s_last) *ticket = 0u;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last;
}

// Exchange buffer of one rank: flags[2][G] then slots[2][G][R][rec]; parity = solve sequence & 1.  A rank can run at
// most one solve ahead of a peer (its next merge needs that peer's next record, which the peer pushes only after its
// own merge of the current solve), so two parities are enough.
__device__ __forceinline__ unsigned int *xchg_flags(void *base, int parity, int G) {
  return reinterpret_cast<unsigned int *>(base) + parity * G;
}
__device__ __forceinline__ float *xchg_slot(void *base, int parity, int g, int G, size_t slot_floats) {
  return reinterpret_cast<float *>(reinterpret_cast<char *>(base) + kExchangeHeaderBytes) +
         ((size_t)parity * G + g) * slot_floats;
}

// The exchange step of a sample-sharded solve, run by ONE block once this rank's records [R][rec_stride] are complete:
// stores them into every rank's exchange buffer (NVLink P2P stores; own buffer included), raises the flags, waits for
// the peers' flags, then merges all ranks' records in rank order (log-sum-exp merge, bit-identical on every rank):
// u_new, warm start, stats, counter++.  A peer that does not arrive within timeout_cycles leaves controls and warm
// start untouched and sets stats[3] = 1 (the host then reports MPPI_ERR_NCCL).
__device__ __forceinline__ void exchange_and_merge(const SolveHeader *__restrict__ hdr, const float *record,
                                                   const ExchangeArgs &x, float *__restrict__ u_new,
                                                   float *__restrict__ nominal, float *__restrict__ stats,
                                                   uint32_t *__restrict__ counter, int planes, int rec_stride, int R,
                                                   bool own_in_regs = false, float own0 = 0.f, float own1 = 0.f) {
  __shared__ int s_timeout;
  const int G = x.G;
  const unsigned int n = *x.seq;
  const int parity = (int)(n & 1u);
  const size_t slot_floats = (size_t)R * rec_stride;
  if (threadIdx.x == 0) s_timeout = 0;
  // every element is read once (or is still in the caller's registers: own_in_regs, one robot, <= 2 elements per
  // thread) and stored to all G buffers: the stores are posted, nothing below depends on them before the fence
  int slot = 0;
  for (size_t k = threadIdx.x; k < slot_floats; k += blockDim.x, ++slot) {
    const float v = own_in_regs ? (slot == 0 ? own0 : own1) : __ldcg(record + k);
    for (int g = 0; g < G; ++g) xchg_slot(x.peers[g], parity, x.rank, G, slot_floats)[k] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < G) {
    volatile unsigned int *f = xchg_flags(x.peers[threadIdx.x], parity, G) + x.rank;
    *f = n + 1u;
    // flags of one parity only grow (n+1, n+3, ...): a later value can never be mistaken for this solve's
    volatile unsigned int *mine = xchg_flags(x.xbuf, parity, G) + threadIdx.x;
    const long long t0 = clock64();
    while (*mine != n + 1u) {
      if (clock64() - t0 > x.timeout_cycles) {
        s_timeout = 1;
        break;
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  const bool timed_out = s_timeout != 0;
  const float inv_lambda = hdr->inv_lambda;
  for (int robot = 0; robot < R; ++robot) {
    // the records were written by peers into this GPU's memory: read them through L2 (__ldcg), never L1
    const float *recs = xchg_slot(x.xbuf, parity, 0, G, slot_floats) + (size_t)robot * rec_stride;
    float m = __ldcg(recs);
    for (int g = 1; g < G; ++g) {
      const float mg = __ldcg(recs + (size_t)g * slot_floats);
      m = mg < m ? mg : m;
    }
    float S = 0.f, Q = 0.f;
    for (int g = 0; g < G; ++g) {
      const float *r = recs + (size_t)g * slot_floats;
      const float a = merge_scale(__ldcg(r), m, inv_lambda, G);
      S = fmaf(a, __ldcg(r + 1), S);
      Q = fmaf(a * a, __ldcg(r + 2), Q);
    }
    if (!timed_out) {
      for (int p = threadIdx.x; p < planes; p += blockDim.x) {
        float N = 0.f;
        for (int g = 0; g < G; ++g) {
          const float *r = recs + (size_t)g * slot_floats;
          N = fmaf(merge_scale(__ldcg(r), m, inv_lambda, G), __ldcg(r + 4 + p), N);
        }
        const float u = N / S;
        u_new[(size_t)robot * planes + p] = u;
        if (nominal) nominal[(size_t)robot * planes + p] = u;
      }
    }
    if (threadIdx.x == 0) {
      stats[robot * 4 + 0] = m;
      stats[robot * 4 + 1] = S;
      stats[robot * 4 + 2] = S * S / Q;
      stats[robot * 4 + 3] = timed_out ? 1.f : 0.f;  // != 0: a peer's record did not arrive in time
    }
  }
  if (threadIdx.x == 0) {
    *counter = *counter + 1u;
    *x.seq = n + 1u;
  }
}

}  // namespace mppi
