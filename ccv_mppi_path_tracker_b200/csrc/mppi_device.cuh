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

// block-level min of the per-thread costs, one atomicMin per block (cmin_slot may be null: none); returns the block's
// minimum to every thread (INFINITY when no thread is valid)
__device__ __forceinline__ float block_min_to_global(float c, bool valid, unsigned int *cmin_slot, float *s_red) {
  float v = valid ? c : INFINITY;
  v = warp_min(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) s_red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float m = lane < nw ? s_red[lane] : INFINITY;
  m = warp_min(m);
  if (cmin_slot && wid == 0 && lane == 0 && m < INFINITY) atomicMin(cmin_slot, float_to_ordered(m));
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

// Exchange buffer of one rank: slots[2][G][R][rec] of 8-byte words {payload (float bits), stamp}; parity = solve
// sequence & 1, stamp = sequence + 1.  Data and "it has arrived" travel in ONE aligned 8-byte store (the protocol NCCL
// calls LL): no fence, no separate flag, no second trip -- the receiver polls the words it needs and has the payload
// in its registers the moment the stamp matches.  A rank can run at most one solve ahead of a peer (its next merge
// needs that peer's next record, which the peer pushes only after its own merge of the current solve), so two
// parities are enough and a stamp can never be mistaken for another solve's.
__device__ __forceinline__ unsigned long long *xchg_slot(void *base, int parity, int g, int G, size_t slot_elems) {
  return reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(base) + kExchangeHeaderBytes) +
         ((size_t)parity * G + g) * slot_elems;
}
__device__ __forceinline__ unsigned long long xchg_pack(float v, unsigned int stamp) {
  return ((unsigned long long)stamp << 32) | (unsigned long long)__float_as_uint(v);
}
constexpr int kExchangeRegRanks = 8;  // ranks whose records the merging block keeps in registers (one robot)

// The exchange step of a sample-sharded solve, run by ONE block once this rank's records [R][rec_stride] are complete
// (in `record`, or -- one robot, at most two elements per thread -- still in the caller's registers): stores them into
// every rank's exchange buffer (NVLink P2P stores; own buffer included), polls the peers' words, then merges all ranks'
// records in rank order (log-sum-exp merge, bit-identical on every rank and to merge_kernel): u_new, warm start, stats,
// counter++.  A peer that does not arrive within timeout_cycles leaves controls and warm start untouched and sets
// stats[3] = 1 (the host then reports MPPI_ERR_NCCL).
__device__ __forceinline__ void exchange_and_merge(const SolveHeader *__restrict__ hdr, const float *record,
                                                   const ExchangeArgs &x, float *__restrict__ u_new,
                                                   float *__restrict__ nominal, float *__restrict__ stats,
                                                   uint32_t *__restrict__ counter, int planes, int rec_stride, int R,
                                                   bool own_in_regs = false, float own0 = 0.f, float own1 = 0.f) {
  __shared__ int s_timeout;
  __shared__ float s_head[3][kExchangeRegRanks];  // {m_g, S_g, Q_g} of every rank (register path)
  const int G = x.G;
  const unsigned int n = *x.seq;
  const unsigned int stamp = n + 1u;
  const int parity = (int)(n & 1u);
  const size_t slot_elems = (size_t)R * rec_stride;
  const float inv_lambda = hdr->inv_lambda;
  const long long t0 = clock64();
  if (threadIdx.x == 0) s_timeout = 0;
  __syncthreads();
  const bool reg_path = R == 1 && G <= kExchangeRegRanks && slot_elems <= 2 * (size_t)blockDim.x;
  if (reg_path) {
    // ---- one robot: every thread owns (up to) two record elements of every rank, start to finish ----------------
    const size_t k0 = threadIdx.x, k1 = threadIdx.x + blockDim.x;
    const bool has0 = k0 < slot_elems, has1 = k1 < slot_elems;
    const float v0 = own_in_regs ? own0 : (has0 ? __ldcg(record + k0) : 0.f);
    const float v1 = own_in_regs ? own1 : (has1 ? __ldcg(record + k1) : 0.f);
    for (int g = 0; g < G; ++g) {  // posted stores, all peers
      unsigned long long *dst = xchg_slot(x.peers[g], parity, x.rank, G, slot_elems);
      if (has0) *reinterpret_cast<volatile unsigned long long *>(dst + k0) = xchg_pack(v0, stamp);
      if (has1) *reinterpret_cast<volatile unsigned long long *>(dst + k1) = xchg_pack(v1, stamp);
    }
    // Poll all ranks' words of this thread TOGETHER: one batch of (up to) 2 G independent loads per trip, repeated
    // until every stamp matches.  (Polling word after word -- each a loop of its own -- cost one L2 round trip per
    // word even when everything had arrived: 2 G dependent trips, ~10 us at G = 8.)
    float a0[kExchangeRegRanks], a1[kExchangeRegRanks];
    bool ok = true;
    {
      const volatile unsigned long long *src0 = xchg_slot(x.xbuf, parity, 0, G, slot_elems);
      for (;;) {
        unsigned long long w0[kExchangeRegRanks], w1[kExchangeRegRanks];
#pragma unroll
        for (int g = 0; g < kExchangeRegRanks; ++g) {
          w0[g] = w1[g] = (unsigned long long)stamp << 32;
          if (g < G) {
            if (has0) w0[g] = src0[(size_t)g * slot_elems + k0];
            if (has1) w1[g] = src0[(size_t)g * slot_elems + k1];
          }
        }
        bool all = true;
#pragma unroll
        for (int g = 0; g < kExchangeRegRanks; ++g) {
          all = all && (unsigned int)(w0[g] >> 32) == stamp && (unsigned int)(w1[g] >> 32) == stamp;
          a0[g] = __uint_as_float((unsigned int)w0[g]);
          a1[g] = __uint_as_float((unsigned int)w1[g]);
        }
        if (all) break;
        if (clock64() - t0 > x.timeout_cycles) {
          ok = false;
          break;
        }
      }
    }
    if (!ok) s_timeout = 1;
    if (threadIdx.x < 3) {
#pragma unroll
      for (int g = 0; g < kExchangeRegRanks; ++g) s_head[threadIdx.x][g] = a0[g];
    }
    __syncthreads();
    const bool timed_out = s_timeout != 0;
    float m = s_head[0][0];
    for (int g = 1; g < G; ++g) {
      const float mg = s_head[0][g];
      m = mg < m ? mg : m;
    }
    float S = 0.f, Q = 0.f, scale[kExchangeRegRanks];
#pragma unroll
    for (int g = 0; g < kExchangeRegRanks; ++g) {
      scale[g] = 0.f;
      if (g < G) {
        scale[g] = merge_scale(s_head[0][g], m, inv_lambda, G);
        S = fmaf(scale[g], s_head[1][g], S);
        Q = fmaf(scale[g] * scale[g], s_head[2][g], Q);
      }
    }
    if (!timed_out) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int col = (int)(half ? k1 : k0);
        if (col >= 4 && col < 4 + planes) {
          float N = 0.f;
#pragma unroll
          for (int g = 0; g < kExchangeRegRanks; ++g)
            if (g < G) N = fmaf(scale[g], half ? a1[g] : a0[g], N);
          const float u = N / S;
          u_new[col - 4] = u;
          if (nominal) nominal[col - 4] = u;
        }
      }
    }
    if (threadIdx.x == 0) {
      stats[0] = m;
      stats[1] = S;
      stats[2] = S * S / Q;
      stats[3] = timed_out ? 1.f : 0.f;  // != 0: a peer's record did not arrive in time
      *counter = *counter + 1u;
      *x.seq = n + 1u;
    }
    return;
  }
  // ---- general case (several robots, many ranks): push, poll every word, merge from the buffer ---------------------
  {
    int slot = 0;
    for (size_t k = threadIdx.x; k < slot_elems; k += blockDim.x, ++slot) {
      const float v = own_in_regs ? (slot == 0 ? own0 : own1) : __ldcg(record + k);
      for (int g = 0; g < G; ++g)
        *reinterpret_cast<volatile unsigned long long *>(xchg_slot(x.peers[g], parity, x.rank, G, slot_elems) + k) =
            xchg_pack(v, stamp);
    }
    // poll the words of eight ranks at a time (independent loads: one L2 round trip per batch once they have arrived)
    bool ok = true;
    for (size_t k = threadIdx.x; k < slot_elems && ok; k += blockDim.x)
      for (int g0 = 0; g0 < G && ok; g0 += 8) {
        const volatile unsigned long long *src = xchg_slot(x.xbuf, parity, g0, G, slot_elems) + k;
        for (;;) {
          bool all = true;
          unsigned long long w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w[j] = g0 + j < G ? src[(size_t)j * slot_elems] : (unsigned long long)stamp << 32;
#pragma unroll
          for (int j = 0; j < 8; ++j) all = all && (unsigned int)(w[j] >> 32) == stamp;
          if (all) break;
          if (clock64() - t0 > x.timeout_cycles) {
            ok = false;
            break;
          }
        }
      }
    if (!ok) s_timeout = 1;
  }
  __syncthreads();
  const bool timed_out = s_timeout != 0;
  auto payload = [&](int g, size_t k) {  // arrived (polled above): a plain L2 read of the word's low half
    return __uint_as_float((unsigned int)__ldcg(xchg_slot(x.xbuf, parity, g, G, slot_elems) + k));
  };
  for (int robot = 0; robot < R; ++robot) {
    const size_t base = (size_t)robot * rec_stride;
    float m = payload(0, base);
    for (int g = 1; g < G; ++g) {
      const float mg = payload(g, base);
      m = mg < m ? mg : m;
    }
    float S = 0.f, Q = 0.f;
    for (int g = 0; g < G; ++g) {
      const float a = merge_scale(payload(g, base), m, inv_lambda, G);
      S = fmaf(a, payload(g, base + 1), S);
      Q = fmaf(a * a, payload(g, base + 2), Q);
    }
    if (!timed_out) {
      for (int p = threadIdx.x; p < planes; p += blockDim.x) {
        float N = 0.f;
        for (int g = 0; g < G; ++g)
          N = fmaf(merge_scale(payload(g, base), m, inv_lambda, G), payload(g, base + 4 + p), N);
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
