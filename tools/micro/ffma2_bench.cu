// Micro-benchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100+) throughput per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ unsigned long long pack(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

template <int kMode>  // 0: 16 scalar FFMA chains, 1: 8 FFMA2 chains (= 16 FMAs), 2: 8 FFMA2 + 4 MUFU per round, 3: 16 FFMA + 4 MUFU
__global__ void __launch_bounds__(256) k(float* out, int iters, float x, float y) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  unsigned long long p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = pack(acc[2 * i], acc[2 * i + 1]);
  const unsigned long long xx = pack(x, x), yy = pack(y, y);
  float m[4] = {1.1f, 1.2f, 1.3f, 1.4f};
  for (int it = 0; it < iters; ++it) {
    if (kMode == 0 || kMode == 3) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(x), "f"(y));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = ffma2(xx, p[i], yy);
    }
    if (kMode >= 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
    }
  }
  float s = m[0] + m[1] + m[2] + m[3];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int kMode>
void run(const char* name, float* d) {
  const int iters = 20000, grid = 148 * 4, block = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<kMode><<<grid, block>>>(d, 100, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  k<kMode><<<grid, block>>>(d, iters, 1.0001f, 0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double)grid * block * iters * 16;
  printf("%-28s %8.3f ms  %7.2f T FMA/s  (%.1f TFLOP/s)\n", name, ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
}

int main() {
  float* d;
  cudaMalloc(&d, 148 * 4 * 256 * sizeof(float));
  run<0>("16 x FFMA", d);
  run<1>("8 x FFMA2", d);
  run<3>("16 x FFMA + 4 MUFU", d);
  run<2>("8 x FFMA2 + 4 MUFU", d);
  return 0;
}
