// FMA-pipe microbenchmark behind profiles/r01_lstm_microbench.txt: FFMA2 (fma.rn.f32x2) vs scalar FFMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench tools/microbench/ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float2 acc[8], w[8], h[8];
  for (int i = 0; i < 8; ++i) { acc[i] = make_float2(threadIdx.x, i); w[i] = make_float2(1.0001f + i, 0.999f); h[i] = make_float2(0.5f + threadIdx.x * 1e-3f, 0.25f + i); }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) acc[i] = ffma2(w[(i + r) & 7], h[i], acc[i]);
        else { acc[i].x = fmaf(w[(i + r) & 7].x, h[i].x, acc[i].x); acc[i].y = fmaf(w[(i + r) & 7].y, h[i].y, acc[i].y); }
      }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&cyc, 8);
  for (int warps : {4, 8, 16}) for (int mode = 0; mode < 2; ++mode) {
    int iters = 2000;
    if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, cyc); else k<1><<<148, warps * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    double per_smsp = (double)iters * 64 * (warps / 4.0) * (mode == 0 ? 1 : 2);
    printf("warps/SM %d mode %s: %lld clk, %.2f clk per warp-instr per SMSP, FMA/clk/SM %.1f\n", warps, mode == 0 ? "FFMA2" : "FFMA",
           *cyc, *cyc / per_smsp, (double)iters * 64 * 2 * warps * 32 / *cyc);
  }
  return 0;
}
