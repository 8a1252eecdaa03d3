// Microbenchmark: per-SM throughput of ex2 variants (lane-results per clock per SM) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu && ./mufu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8];
  uint32_t p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i + 1); p[i] = 0xBC00BC00u + threadIdx.x * 65537u * (i + 1); }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(p[i]));
      if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(p[i]));
      if (MODE == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (MODE == 4) asm volatile("{.reg .b32 t; cvt.rn.bf16x2.f32 t, %0, %0; mov.b32 %0, t;}" : "+f"(a[i]));
    }
  }
  long long t1 = clock64();
  __syncthreads();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(p[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_lane) {
  float* out; long long* cyc;
  const int blocks = 148, threads = 512, iters = 4096;
  cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
  k<MODE><<<blocks, threads>>>(out, cyc, 16);
  k<MODE><<<blocks, threads>>>(out, cyc, iters);
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < blocks; ++i) c += h[i]; c /= blocks;
  double ops = (double)threads * iters * 8 * per_lane;
  printf("%-28s %8.2f results/clk/SM  (%.0f cycles)\n", name, ops / c, c);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.ftz.bf16x2", 2);
  run<2>("ex2.approx.f16x2", 2);
  run<3>("fma.rn.f32 (3-reg)", 1);
  run<4>("cvt.rn.bf16x2.f32", 2);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
