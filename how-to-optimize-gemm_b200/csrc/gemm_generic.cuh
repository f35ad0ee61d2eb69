// gemm_generic.cuh — shape/stride-agnostic CUDA-core kernels (still GPU; there is no CPU path).
//
// They serve what TMA cannot describe: leading dimensions or base pointers that are not 16-byte
// aligned (TMA needs ld*sizeof % 16 == 0), and degenerate sizes.  The reference's hand kernels
// simply assume m,n % 128 == 0 and ignore lda/ldb/ldc (cuda/MMult_cuda_12.cu:231-234); chgemm's
// headline feature is that it does not (aarch64-int8/int8kernel_m4.S:62-93).  Sequential-k
// accumulation, so the fp32 instance keeps the strict contract of gemm_ffma.cuh.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>
#include <type_traits>

#include "ptx.cuh"

namespace b200 {

template <typename T> struct LoadAs;
template <> struct LoadAs<float>   { using Acc = float;   __device__ static float   ld(const float* p)   { return *p; } };
template <> struct LoadAs<int8_t>  { using Acc = int32_t; __device__ static int32_t ld(const int8_t* p)  { return (int32_t)*p; } };
template <> struct LoadAs<uint16_t>{ using Acc = float;   __device__ static float   ld(const uint16_t* p){ return __uint_as_float((uint32_t)*p << 16); } };

template <typename Acc, typename OutT> __device__ __forceinline__ void store_out(OutT* p, Acc v);
template <> __device__ __forceinline__ void store_out<float, float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void store_out<int32_t, int32_t>(int32_t* p, int32_t v) { *p = v; }
template <> __device__ __forceinline__ void store_out<float, uint16_t>(uint16_t* p, float v) {
  *p = __bfloat16_as_ushort(__float2bfloat16_rn(v));
}

// 64x64 tile, 256 threads, 4x4 per thread, BK = 16.
template <typename InT, typename OutT>
__global__ void __launch_bounds__(256)
gemm_generic_kernel(int M, int N, int K, const InT* __restrict__ A, long long lda,
                    const InT* __restrict__ B, long long ldb, OutT* __restrict__ C, long long ldc,
                    int accumulate, const float* __restrict__ rq_scale = nullptr,
                    const float* __restrict__ rq_bias = nullptr) {
  using Acc = typename LoadAs<InT>::Acc;
  __shared__ Acc As[16][64 + 4];
  __shared__ Acc Bs[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  Acc acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0;
  if constexpr (std::is_same<Acc, OutT>::value) {
    if (accumulate) {          // chain starts from C(i,j): the CPU harness contract C += A*B
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int gm = m0 + ty + 16 * i, gn = n0 + tx + 16 * j;
          if (gm < M && gn < N) acc[i][j] = C[(long long)gm * ldc + gn];
        }
    }
  }

  for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int idx = threadIdx.x + r * 256;          // 0..1023
      const int am = idx >> 4, ak = idx & 15;          // A tile 64 x 16, k fastest
      const int gm = m0 + am, gk = k0 + ak;
      As[ak][am] = (gm < M && gk < K) ? LoadAs<InT>::ld(A + (long long)gm * lda + gk) : (Acc)0;
      const int bk = idx >> 6, bn = idx & 63;          // B tile 16 x 64, n fastest
      const int gk2 = k0 + bk, gn = n0 + bn;
      Bs[bk][bn] = (gk2 < K && gn < N) ? LoadAs<InT>::ld(B + (long long)gk2 * ldb + gn) : (Acc)0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; kk++) {
      Acc a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if constexpr (sizeof(Acc) == 4 && !std::is_same<Acc, int32_t>::value)
            acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
          else
            acc[i][j] += a[i] * b[j];
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int gm = m0 + ty + 16 * i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int gn = n0 + tx + 16 * j;
      if (gn < N) {
        if constexpr (std::is_same<OutT, int8_t>::value && std::is_same<Acc, int32_t>::value) {
          // requantising store (int8 C): per-row scale and optional per-row bias
          C[(long long)gm * ldc + gn] = (int8_t)requant_s8(acc[i][j], rq_scale[gm], rq_bias ? rq_bias[gm] : 0.0f,
                                                            rq_bias != nullptr);
        } else {
          store_out<Acc, OutT>(C + (long long)gm * ldc + gn, acc[i][j]);
        }
      }
    }
  }
}

// ---- small element-wise helpers ------------------------------------------------------------
__global__ void convert_f32_to_bf16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst,
                                           size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = __bfloat16_as_ushort(__float2bfloat16_rn(src[i]));
}

// C(i,j) += T(i,j): the CPU harness contract C += A*B (aarch64/MMult0.cpp:16) on top of C = A*B.
// C *= s over an m x n window (the beta / alpha passes of the general epilogue on the CUDA-core paths)
__global__ void scale_inplace_kernel(int M, int N, float* __restrict__ C, long long ldc, float s) {
  for (int r = blockIdx.y; r < M; r += gridDim.y)
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x) C[(long long)r * ldc + c] *= s;
}

template <typename T>
__global__ void fill_zero_kernel(int M, int N, T* __restrict__ C, long long ldc) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  for (int i = blockIdx.y; i < M; i += gridDim.y) C[(long long)i * ldc + j] = (T)0;
}

}  // namespace b200
