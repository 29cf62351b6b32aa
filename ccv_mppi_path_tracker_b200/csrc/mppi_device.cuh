// mppi_device.cuh -- device-side helpers shared by the kernel translation units (warp/block reductions,
// parameter staging).  Internal; not part of the ABI.
#pragma once
#include "mppi_kernels.h"

namespace mppi {

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

__device__ __forceinline__ void load_params_to_shared(SolveParams *dst, const SolveHeader *hdr) {
  const uint32_t *src = reinterpret_cast<const uint32_t *>(&hdr->P);
  uint32_t *d = reinterpret_cast<uint32_t *>(dst);
  for (int k = threadIdx.x; k < (int)(sizeof(SolveParams) / 4); k += blockDim.x) d[k] = src[k];
}

}  // namespace mppi
