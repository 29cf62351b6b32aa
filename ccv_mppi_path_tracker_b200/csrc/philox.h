// philox.h -- Philox4x32-10 counter-based generator (Salmon et al., SC'11), host + device.
// Replaces the reference's per-call std::mt19937 + std::normal_distribution (src/diff_drive_mppi.cpp:83-97):
// a counter-based stream lets every (robot, t, u, sample) element be generated independently, in any order,
// on any GPU of a sharded solve.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PHILOX_HD __host__ __device__ __forceinline__
#else
#define PHILOX_HD inline
#endif

namespace mppi {

struct Philox4 {
  uint32_t v[4];
};

PHILOX_HD void philox_mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
#if defined(__CUDA_ARCH__)
  hi = __umulhi(a, b);
  lo = a * b;
#else
  uint64_t p = (uint64_t)a * (uint64_t)b;
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
#endif
}

PHILOX_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u;
  const uint32_t kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    philox_mulhilo(kM0, c0, hi0, lo0);
    philox_mulhilo(kM1, c2, hi1, lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += kW0;
    k1 += kW1;
  }
  Philox4 out;
  out.v[0] = c0;
  out.v[1] = c1;
  out.v[2] = c2;
  out.v[3] = c3;
  return out;
}

// Counter layout of the noise tensor: one Philox block yields the 4 normals of samples 4q..4q+3 of one
// (robot, t, u) plane.  c0 = q (global sample index / 4), c1 = plane index t*U+u, c2 = global robot index,
// c3 = solve counter; key = 64-bit seed.
struct NoiseCounter {
  uint32_t c0, c1, c2, c3;
};

}  // namespace mppi
