// Micro-benchmark: FP32 issue rates on sm_100a -- scalar FFMA vs packed fma.rn.f32x2, and the scan's op mix.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp32 ubench_fp32.cu
#include <cuda_runtime.h>
#include <stdio.h>

#define ITERS 4096

__global__ void k_ffma(float *out, float a, float b) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float *out, float a, float b) {
  unsigned long long x[8];
  unsigned long long aa, bb;
  asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float f = threadIdx.x * 0.001f + i;
    asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(f));
  }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the literal scan's mix per point: 2 sub, 1 mul, 1 fma, 1 min (scalar)
__global__ void k_scan_scalar(float *out, const float2 *win, int T, float px, float py) {
  extern __shared__ float2 sw[];
  for (int j = threadIdx.x; j < T; j += blockDim.x) sw[j] = win[j];
  __syncthreads();
  float x = px + threadIdx.x * 1e-3f, y = py;
  float acc = 0;
  for (int it = 0; it < 256; ++it) {
    float best = 1e4f;
    for (int j = 0; j < T; ++j) {
      float dx = x - sw[j].x, dy = y - sw[j].y;
      best = fminf(best, fmaf(dy, dy, dx * dx));
    }
    acc += best;
    x += 1e-3f;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// packed: two points per f32x2 op; window stored as {x0,x1},{y0,y1} pairs
__global__ void k_scan_packed(float *out, const float4 *win2, int T2, float px, float py) {
  extern __shared__ float4 sw4[];
  for (int j = threadIdx.x; j < T2; j += blockDim.x) sw4[j] = win2[j];
  __syncthreads();
  float x = px + threadIdx.x * 1e-3f, y = py;
  float acc = 0;
  for (int it = 0; it < 256; ++it) {
    float best = 1e4f;
    unsigned long long xx, yy;
    asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
    asm("mov.b64 %0, {%1, %1};" : "=l"(yy) : "f"(y));
    for (int j = 0; j < T2; ++j) {
      float4 w = sw4[j];  // {x0, x1, y0, y1}
      unsigned long long wx, wy, dx, dy, m, d;
      asm("mov.b64 %0, {%1, %2};" : "=l"(wx) : "f"(w.x), "f"(w.y));
      asm("mov.b64 %0, {%1, %2};" : "=l"(wy) : "f"(w.z), "f"(w.w));
      asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(xx), "l"(wx));
      asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(yy), "l"(wy));
      asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(m) : "l"(dx));
      asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(d) : "l"(dy), "l"(m));
      float d0, d1;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
      best = fminf(best, fminf(d0, d1));
    }
    acc += best;
    x += 1e-3f;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  float *out;
  const int blocks = 148 * 8, threads = 256;
  cudaMalloc(&out, blocks * threads * sizeof(float));
  const int T = 104;
  float2 hw[T];
  float4 hw4[T / 2];
  for (int j = 0; j < T; ++j) hw[j] = make_float2(0.12f * j, 0.3f * sinf(0.2f * j));
  for (int j = 0; j < T / 2; ++j) hw4[j] = make_float4(hw[2 * j].x, hw[2 * j + 1].x, hw[2 * j].y, hw[2 * j + 1].y);
  float2 *dw; float4 *dw4;
  cudaMalloc(&dw, sizeof hw); cudaMalloc(&dw4, sizeof hw4);
  cudaMemcpy(dw, hw, sizeof hw, cudaMemcpyHostToDevice);
  cudaMemcpy(dw4, hw4, sizeof hw4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * blocks * threads * 8.0 * ITERS;
    printf("FFMA   : %.3f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
    cudaEventRecord(e0); k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("FFMA2  : %.3f ms  %.1f TFLOP/s\n", ms, 2 * fl / ms / 1e9);
    cudaEventRecord(e0); k_scan_scalar<<<blocks, threads, T * sizeof(float2)>>>(out, dw, T, 1.0f, 0.2f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double pairs = (double)blocks * threads * 256.0 * T;
    printf("scan scalar: %.3f ms  %.2f Gpair/s  (%.2f pair/clk/SM @1.965GHz)\n", ms, pairs / ms / 1e6, pairs / (ms * 1e-3) / 148 / 1.965e9);
    cudaEventRecord(e0); k_scan_packed<<<blocks, threads, T / 2 * sizeof(float4)>>>(out, dw4, T / 2, 1.0f, 0.2f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("scan packed: %.3f ms  %.2f Gpair/s  (%.2f pair/clk/SM @1.965GHz)\n", ms, pairs / ms / 1e6, pairs / (ms * 1e-3) / 148 / 1.965e9);
  }
  printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
