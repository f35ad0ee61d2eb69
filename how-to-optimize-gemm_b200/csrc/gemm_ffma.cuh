// gemm_ffma.cuh — the strict-fp32 path: CUDA-core FFMA SGEMM for sm_100a, fed by TMA.
//
// Arithmetic contract (what makes this the drop-in for cuda/MMult_cuda_12.cu:200-206 and bit-exact
// against the reference's naive oracle as its own makefile builds it, aarch64/REF_MMult.cpp:24 with
// GCC's fused multiply-add): every C(i,j) is ONE accumulator chain
//     c = 0;  for p = 0..k-1 (ascending):  c = fma(A(i,p), B(p,j), c)
// no split-K, no reassociation.
//
// Structure: 128x128 CTA tile, 256 threads (8 warps laid out 4(ty) x 8(tx) lanes), 8x8 outputs per
// thread as rows ty+16*i, column groups tx*4+64*j.  The reference's gmem->reg->smem double buffer
// (cuda/MMult_cuda_12.cu:113-198) becomes a 3-deep TMA ring guarded by mbarriers: no LDG/STS issue
// slots are spent on staging.  A lands K-contiguous with SWIZZLE_128B so four consecutive rows read
// by a warp hit distinct banks; B lands N-contiguous (512-byte rows).
#pragma once
#include "ptx.cuh"

namespace b200 {

struct FfmaParams {
  float* C;
  long long ldc;
  int M, N, K;
  int vec_ok;
  int accumulate;   // 1: accumulator chains start from C(i,j) (C += A*B), 0: from zero (C = A*B)
};

struct FfmaCfg {
  static constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3;
  static constexpr int A_STAGE = BM * BK * 4;     // 16 KB, 128 rows x 128 B (swizzled)
  static constexpr int B_STAGE = BK * BN * 4;     // 16 KB, 32 rows x 512 B
  static constexpr int STAGE_BYTES = A_STAGE + B_STAGE;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 2 * STAGES * 8;
  static constexpr int THREADS = 256;
};

__global__ void __launch_bounds__(256, 2)
gemm_ffma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const FfmaParams p) {
  using Cfg = FfmaCfg;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + Cfg::STAGES * Cfg::A_STAGE;
  const uint32_t bar_full = sB + Cfg::STAGES * Cfg::B_STAGE;
  const uint32_t bar_empty = bar_full + 8 * Cfg::STAGES;
  const uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ty = (warp >> 1) * 4 + (lane >> 3);   // 0..15
  const int tx = (warp & 1) * 8 + (lane & 7);     // 0..15
  const int m0 = blockIdx.y * Cfg::BM, n0 = blockIdx.x * Cfg::BN;
  const int num_kb = (p.K + Cfg::BK - 1) / Cfg::BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < Cfg::STAGES; i++) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 8);             // one arrive per consumer warp
    }
    fence_barrier_init();
  }
  __syncthreads();

  auto issue = [&](int kb) {                        // thread 0 only
    const int s = kb % Cfg::STAGES;
    const uint32_t full = bar_full + 8 * s;
    mbar_arrive_expect_tx(full, Cfg::STAGE_BYTES);
    tma_load_2d(sA + s * Cfg::A_STAGE, &tmA, full, kb * Cfg::BK, m0);
    tma_load_2d(sB + s * Cfg::B_STAGE, &tmB, full, n0, kb * Cfg::BK);
  };
  if (threadIdx.x == 0) {
    for (int kb = 0; kb < Cfg::STAGES - 1 && kb < num_kb; kb++) issue(kb);
  }

  // Accumulators as 64-bit pairs: Blackwell's packed FFMA2 (fma.rn.f32x2) performs two fused
  // multiply-adds per lane per instruction; ptxas folds the scalar A operand into the instruction's
  // broadcast form (FFMA2 Rd, Ra.F32, Rb.F32x2, Rc.F32x2), so one k-step is 32 FFMA2 instead of 64
  // FFMA — half the issue slots and register-port reads, same rounding (each lane is an IEEE fma).
  float2 acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = make_float2(0.0f, 0.0f);
  if (p.accumulate) {
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int gm = m0 + ty + 16 * i, gn = n0 + tx * 4 + 64 * (j >> 2) + (j & 3);
        if (gm < p.M && gn < p.N) {
          const float v = p.C[(long long)gm * p.ldc + gn];
          if (j & 1) acc[i][j >> 1].y = v; else acc[i][j >> 1].x = v;
        }
      }
  }

  // per-thread smem offsets: A rows ty+16*i -> (row & 7) == (ty & 7) for every i
  const uint32_t a_row_off = ty * 128;
  const uint32_t a_swz = (ty & 7);
  const uint32_t b_col_off = tx * 16;

  for (int kb = 0; kb < num_kb; kb++) {
    const int s = kb % Cfg::STAGES;
    const uint32_t use = kb / Cfg::STAGES;
    if (threadIdx.x == 0) {
      const int nk = kb + Cfg::STAGES - 1;           // refill the slot consumed in iteration kb-1
      if (nk < num_kb) {
        if (kb >= 1) mbar_wait(bar_empty + 8 * (nk % Cfg::STAGES), ((nk / Cfg::STAGES) - 1) & 1);
        issue(nk);
      }
    }
    mbar_wait(bar_full + 8 * s, use & 1);
    const uint8_t* As = smem_gen + (sA - smem_base) + s * Cfg::A_STAGE;
    const uint8_t* Bs = smem_gen + (sB - smem_base) + s * Cfg::B_STAGE;
#pragma unroll 1
    for (int kc = 0; kc < Cfg::BK / 4; kc++) {
      float4 a4[8];
#pragma unroll
      for (int i = 0; i < 8; i++)
        a4[i] = *reinterpret_cast<const float4*>(As + a_row_off + i * (16 * 128) +
                                                 ((kc ^ a_swz) << 4));
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        const uint8_t* brow = Bs + (kc * 4 + kk) * 512 + b_col_off;
        const float4 b0 = *reinterpret_cast<const float4*>(brow);
        const float4 b1 = *reinterpret_cast<const float4*>(brow + 256);
        const float2 bv[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w),
                              make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const float av = kk == 0 ? a4[i].x : kk == 1 ? a4[i].y : kk == 2 ? a4[i].z : a4[i].w;
          const float2 aa = make_float2(av, av);
#pragma unroll
          for (int j = 0; j < 4; j++) acc[i][j] = __ffma2_rn(aa, bv[j], acc[i][j]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * s);
  }

  // epilogue: 16-byte stores; for fixed (i, j) a warp writes 4 rows x 128 contiguous bytes
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int gm = m0 + ty + 16 * i;
    if (gm >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int gn = n0 + tx * 4 + 64 * j;
      float* dst = p.C + (long long)gm * p.ldc + gn;
      if (p.vec_ok && gn + 4 <= p.N) {
        *reinterpret_cast<float4*>(dst) =
            make_float4(acc[i][2 * j].x, acc[i][2 * j].y, acc[i][2 * j + 1].x, acc[i][2 * j + 1].y);
      } else {
        const float ev[4] = {acc[i][2 * j].x, acc[i][2 * j].y, acc[i][2 * j + 1].x, acc[i][2 * j + 1].y};
#pragma unroll
        for (int e = 0; e < 4; e++)
          if (gn + e < p.N) dst[e] = ev[e];
      }
    }
  }
}

}  // namespace b200
