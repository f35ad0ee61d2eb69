// Microbenchmark: sustained fp32 FMA throughput of one B200 — scalar FFMA vs packed FFMA2 — to know
// what "100 % of the CUDA-core roofline" means for gemm_ffma.cuh.   nvcc -arch=sm_100a -O3 ffma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a0, float b0) {
  float2 acc[16];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  float a = a0 + threadIdx.x * 1e-6f;
  float2 b = make_float2(b0, b0 * 0.5f);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 16; i++) {
        if (MODE == 2) acc[i] = __ffma2_rn(make_float2(a, a), b, acc[i]);
        else { acc[i].x = fmaf(a, b.x, acc[i].x); acc[i].y = fmaf(a, b.y, acc[i].y); }
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 20000;
  for (int mode = 1; mode <= 2; mode++)
    for (int bps = 1; bps <= 4; bps *= 2) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (mode == 1) k<1><<<148 * bps, 256>>>(out, iters, 1.0001f, 0.9999f); else k<2><<<148 * bps, 256>>>(out, iters, 1.0001f, 0.9999f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double flops = 2.0 * 148 * bps * 256 * (double)iters * 4 * 16 * 2;
      printf("mode %s  CTAs/SM %d (warps/SM %d): %.2f TFLOP/s\n", mode == 1 ? "FFMA " : "FFMA2", bps, bps * 8, flops / ms / 1e9);
    }
  return 0;
}
