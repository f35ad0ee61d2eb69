// Microbenchmark 2: the GEMM inner product shape without memory — 8 scalar a's x 4 (or 8) b pairs,
// 32 (64) FFMA2 into distinct accumulator pairs per step, operands rotated every step so nothing is
// loop-invariant.  Tells whether register-file operand bandwidth caps the 8x8 / 8x16 FFMA2 tiles.
#include <cstdio>
#include <cuda_runtime.h>
template <int NB>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
  float2 acc[8][NB];
  float a[8]; float2 b[NB];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = seed + i * 1e-3f + threadIdx.x * 1e-6f;
#pragma unroll
    for (int j = 0; j < NB; j++) acc[i][j] = make_float2(i, j); }
#pragma unroll
  for (int j = 0; j < NB; j++) b[j] = make_float2(seed * (j + 1), seed * 0.5f * (j + 1));
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) acc[i][j] = __ffma2_rn(make_float2(a[i], a[i]), b[j], acc[i][j]);
      // rotate operands (cheap, keeps them live and changing)
      float t = a[0];
#pragma unroll
      for (int i = 0; i < 7; i++) a[i] = a[i + 1];
      a[7] = t;
      float2 tb = b[0];
#pragma unroll
      for (int j = 0; j < NB - 1; j++) b[j] = b[j + 1];
      b[NB - 1] = tb;
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) s += acc[i][j].x + acc[i][j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 4 * 256 * 4);
  const int iters = 5000;
  for (int nb = 4; nb <= 8; nb += 4)
    for (int bps = 1; bps <= 2; bps++) {
      if (nb == 8 && bps == 2) continue;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (nb == 4) k<4><<<148 * bps, 256>>>(out, iters, 1.0001f); else k<8><<<148 * bps, 256>>>(out, iters, 1.0001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double flops = 2.0 * 148 * bps * 256 * (double)iters * 4 * 8 * nb * 2;
      printf("8 x %d-pair outer product, CTAs/SM %d: %.2f TFLOP/s\n", nb, bps, flops / ms / 1e9);
    }
  return 0;
}
