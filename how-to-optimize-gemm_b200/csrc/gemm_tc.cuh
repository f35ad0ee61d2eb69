// gemm_tc.cuh — the tensor-core path: persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[m x n] = A[m x k] * B[k x n],  all row-major.
//
// Mapping of the reference's roles onto Blackwell:
//   packA / packB  (aarch64/MMult_4x4_21.cpp:459-572; the gmem->smem staging of
//                   cuda/MMult_cuda_12.cu:113-198)      -> TMA bulk-tensor copies with hardware swizzle
//                                                          into a STAGES-deep shared-memory ring
//   4x4 / 8x12 register micro-kernel (kernel_8x12,
//                   cuda/MMult_cuda_12.cu:200-206)      -> one thread issuing tcgen05.mma 128 x BN x K
//                                                          into a TMEM accumulator (fp32 / int32)
//   stg128 epilogue (cuda/MMult_cuda_12.cu:210-222)     -> tcgen05.ld -> swizzled smem transpose ->
//                                                          coalesced 16-byte st.global
//
// Row-major B is the MMA's "MN-major" operand: no transpose pass (the job of reorder_b / trans_w in
// aarch64-int8/MMult_4x8_21.c:45-71) exists here; TMA drops [BK x 128B] column blocks of B straight
// into the canonical MN-major swizzled layout and the descriptor walks them (LBO = block stride).
//
// Split-precision fp32 (B200_F32_BF16X3 / _BF16X2): A and B arrive as NPA / NPB stacked bf16 "planes"
// (a = a1 + a2 + a3, produced by split_planes_kernel); each k-block stage holds every plane once and
// the MMA warp issues the listed plane products into the SAME fp32 accumulator, smallest terms first.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner),
// warps 2..5 = epilogue (TMEM lane quadrant = warp_idx % 4).  Three pipelines: smem full/empty,
// TMEM full/empty (two accumulator stages, so the epilogue of tile i overlaps the mainloop of i+1),
// and the static persistent tile schedule.
#pragma once
#include <cuda_fp16.h>
#include <string.h>
#include <type_traits>

#include "ptx.cuh"

namespace b200 {

template <int KIND> struct KindTraits;
// B_LAYOUT / B_SBO: MN-major B uses SWIZZLE_128B (8 k-rows of 128 B per atom, SBO 1024) except for
// 32-bit elements, where the only MN-major layout the tensor core accepts is SWIZZLE_128B_BASE32B
// (32-byte swizzle granules, 4 k-rows per atom, SBO 512; TMA side CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).
// Measured on B200: the 16-byte-atom layout with kind::tf32 returns all-zero accumulators.
template <> struct KindTraits<KIND_F16>  { static constexpr int ELEM = 2, UMMA_K = 16, AB_FMT = 1, C_FMT = 1, B_LAYOUT = 2, B_SBO = 1024; };
template <> struct KindTraits<KIND_TF32> { static constexpr int ELEM = 4, UMMA_K = 8,  AB_FMT = 2, C_FMT = 1, B_LAYOUT = 1, B_SBO = 512; };
template <> struct KindTraits<KIND_FP16> { static constexpr int ELEM = 2, UMMA_K = 16, AB_FMT = 0, C_FMT = 1, B_LAYOUT = 2, B_SBO = 1024; };
template <> struct KindTraits<KIND_I8>   { static constexpr int ELEM = 1, UMMA_K = 32, AB_FMT = 1, C_FMT = 2, B_LAYOUT = 2, B_SBO = 1024; };

// Plane products issued per k-step.  Single: plain GEMM.  X3: a=a1+a2+a3, b likewise, all terms down
// to 2^-16 relative (a1b3, a3b1, a2b2, a1b2, a2b1, a1b1) — dropped terms are <= 2^-24.  X2: two planes,
// three terms, dropped a2b2 ~ 2^-16.
struct ProdSingle {
  static constexpr int N = 1, NPA = 1, NPB = 1;
  __host__ __device__ static constexpr int ia(int) { return 0; }
  __host__ __device__ static constexpr int ib(int) { return 0; }
};
struct ProdX3 {   // (ia,ib): (0,2) (2,0) (1,1) (0,1) (1,0) (0,0)
  static constexpr int N = 6, NPA = 3, NPB = 3;
  __host__ __device__ static constexpr int ia(int i) { return i == 1 ? 2 : (i == 2 || i == 4) ? 1 : 0; }
  __host__ __device__ static constexpr int ib(int i) { return i == 0 ? 2 : (i == 2 || i == 3) ? 1 : 0; }
};
struct ProdX2 {   // (0,1) (1,0) (0,0)
  static constexpr int N = 3, NPA = 2, NPB = 2;
  __host__ __device__ static constexpr int ia(int i) { return i == 1 ? 1 : 0; }
  __host__ __device__ static constexpr int ib(int i) { return i == 0 ? 1 : 0; }
};

struct TcParams {
  void* C;
  long long ldc;           // elements
  int M, N, K;
  int tiles_m, tiles_n;
  int group_m;             // rasterisation: tiles are walked m-fastest inside groups of group_m rows
  int vec_ok;              // C base and ldc allow 16-byte stores
  int a_plane_rows;        // row offset between stacked A planes (split modes), else 0
  int b_plane_rows;        // row offset between stacked B planes (split modes), else 0
  int chunk_kb;            // k-blocks per accumulation chunk (two-level accumulation), >= num_kb: off
  // Wave-quantisation tail: work items [0, full_tiles) are whole tiles; every later tile is cut
  // into `split` K-ranges processed by different CTAs/pairs in the last round and folded into C in
  // order (part p waits for flag == p on its 32-row strip, adds, then publishes p+1).
  int full_tiles, split;
  int halfn;               // 1: tail tiles are issued as two half-width (BN/2) tiles instead of K parts
  int* flags;              // [tail tile][cta rank][epilogue warp], zero between launches
  // Dynamic tile scheduler: work items are handed out in order by an atomic counter (zero between launches; the
  // unit that draws the last sentinel resets it), so a CTA that starts late — SMs held by a co-running kernel such
  // as NCCL's copy kernels — simply draws fewer tiles instead of delaying the whole grid.  Null = static round robin.
  int* sched_counter;
  // Scaled split mode (B200_F32_F16X2): operands were multiplied by 2^-e(row) / 2^-e(col) before the
  // fp16 split; the epilogue multiplies back by 2^e(row) * 2^e(col), exact.  Null = no scaling.
  const float* row_max;    // [M] max |A(i,:)|
  const float* col_max;    // [N] max |B(:,j)|
  // General epilogue C = alpha * (A*B) + beta * C (cuBLAS semantics, cuda/MMult_cuBLAS_1.cpp:11-19); fp32 output
  // only.  axpby == 0: alpha = 1 and beta = accumulate (the two contracts the reference's harnesses use).
  float alpha, beta;
  int axpby;
  int accumulate;
  int stream_c;       // 1: the split modes' single pass over C uses streaming (evict-first) stores, so C does not push the operand planes out of L2
  int epi_direct;     // 1: non-folding passes store straight from registers (tuning hook)          // 1: C += A*B (every partial, including the first, is folded into C); fp32/int32 only
  int dbg_b_lbo, dbg_b_sbo;  // 0 = defaults (probe hook, see b200_gemm_debug_set_b_desc)
};

// REGACC (split-precision fp32 modes): the running fp32 sum of a tile lives in the REGISTERS of eight
// epilogue warps (lane = row, 128 columns per warp) and every K-chunk's TMEM accumulator is added to it
// with a rounded fp32 add; C is touched once per tile.  Otherwise four epilogue warps drain one
// accumulator per tile.
template <int KIND, int BN, int STAGES, class Prod, int A_ROW_BYTES, int CG = 1, int EPIW = 4>
struct TcConfig {
  using T = KindTraits<KIND>;
  static constexpr bool REGACC = Prod::N > 1;
  static constexpr int EPI_WARPS = REGACC ? 8 : EPIW;          // EPIW = 8: two warps per TMEM lane quadrant, half the column passes each
  static constexpr int EPI_WARP0 = REGACC ? 4 : 2;             // first epilogue warp (warpgroup-aligned for setmaxnreg)
  static constexpr int BM = 128;
  static constexpr int BK = A_ROW_BYTES / T::ELEM;          // one swizzled row of K per stage
  static constexpr int A_PLANE = BM * A_ROW_BYTES;          // 16 KB (SW128) or 8 KB (SW64)
  static constexpr int A_LAYOUT = A_ROW_BYTES == 128 ? 2 : 4;   // UMMA layout type: SWIZZLE_128B / SWIZZLE_64B
  static constexpr int A_SBO = 8 * A_ROW_BYTES;             // 8-row core-matrix group stride
  static constexpr int B_BOX_COLS = 128 / T::ELEM;          // elements per 128B-wide column block
  static constexpr int BN_CTA = BN / CG;                    // B columns held by each CTA of the pair
  static constexpr int B_BOXES = BN_CTA / B_BOX_COLS;
  static constexpr int B_BOX_BYTES = BK * 128;
  static constexpr int B_PLANE = B_BOXES * B_BOX_BYTES;
  static constexpr int A_STAGE = Prod::NPA * A_PLANE;
  static constexpr int B_STAGE = Prod::NPB * B_PLANE;
  static constexpr int STAGE_BYTES = A_STAGE + B_STAGE;     // per CTA
  static constexpr int TX_BYTES = CG * STAGE_BYTES;         // bytes landing on the (leader's) full barrier
  static constexpr int MMAS_PER_STAGE = BK / T::UMMA_K;
  static constexpr int A_KADV = T::UMMA_K * T::ELEM;        // 32 B inside the swizzled row
  static constexpr int B_KADV = T::UMMA_K * 128;            // UMMA_K k-rows of 128 B
  static constexpr int EPI_STAGING = EPI_WARPS * 32 * 128;  // per epilogue warp: 32 rows x 128 B
  static constexpr int ACC_STRIDE = BN <= 128 ? 128 : 256;  // TMEM columns between the two accumulators
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int SCHED_SLOTS = 8;                     // ring of published work items (never the limiter: draws are gated by the producer, see run_tile_scheduler)
  static constexpr int SCHED_WARP = REGACC ? 2 : EPI_WARP0 + EPI_WARPS;   // REGACC: an otherwise idle warp of the data-movement warpgroup
  static constexpr int NUM_BARS = 2 * STAGES + 4 + 2 * SCHED_SLOTS + 1;
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + EPI_STAGING +
                                    NUM_BARS * 8 + 16 + 4 * SCHED_SLOTS;
  static constexpr int THREADS = 32 * (EPI_WARP0 + EPI_WARPS + (REGACC ? 0 : 1));   // + the scheduler warp
  static_assert(!REGACC || BN % 64 == 0, "register accumulation splits the tile columns over two warp sets of 32-column groups");
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory of sm_100");
  static_assert(BN_CTA % B_BOX_COLS == 0 && BN % 16 == 0 && BN <= 256, "invalid BN");
  static constexpr int TILE_M = 128 * CG;                   // rows of C per work unit (CTA or CTA pair)
};

__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int group_m, int& mb,
                                            int& nb) {
  const int per_group = group_m * tiles_n;
  const int g = t / per_group;
  const int first_m = g * group_m;
  const int rows = min(group_m, tiles_m - first_m);
  const int r = t - g * per_group;
  mb = first_m + r % rows;
  nb = r / rows;
}

// Scaled split mode (B200_F32_F16X2): exponent e with maxv * 2^-e in [0.5, 1); 0 for zero, denormal,
// inf or NaN maxima.  Range [-125, 128].
__host__ __device__ __forceinline__ int pow2_exp(float maxv) {
  uint32_t bits;
#ifdef __CUDA_ARCH__
  bits = __float_as_uint(maxv);
#else
  memcpy(&bits, &maxv, 4);
#endif
  const int ef = (int)((bits >> 23) & 0xFF);
  return (ef == 0 || ef == 255) ? 0 : ef - 126;
}
// 2^e as a float, e in [-126, 127]
__device__ __forceinline__ float exp2i(int e) { return __uint_as_float((uint32_t)(127 + e) << 23); }
// x * 2^e for any e in [-256, 256] through two normal power-of-two factors (exact unless the result
// itself leaves the fp32 range)
__device__ __forceinline__ float mul_pow2(float x, int e) {
  const int h = e >> 1;
  return x * exp2i(h) * exp2i(min(e - h, 127));
}

struct WorkItem { int tile, part, kb0, kb1, nsub, bn; };
template <int BN>
__device__ __forceinline__ WorkItem work_item(int w, const TcParams& p, int num_kb) {
  WorkItem it;
  it.nsub = 0; it.bn = BN;
  if (w < p.full_tiles) { it.tile = w; it.part = 0; it.kb0 = 0; it.kb1 = num_kb; return it; }
  const int r = w - p.full_tiles;
  if (p.halfn) {        // tail round as half-width tiles: same K chain, no cross-CTA reduction
    it.tile = p.full_tiles + (r >> 1); it.nsub = r & 1; it.bn = BN / 2;
    it.part = 0; it.kb0 = 0; it.kb1 = num_kb;
    return it;
  }
  it.tile = p.full_tiles + r / p.split;
  it.part = r - (r / p.split) * p.split;
  it.kb0 = (int)((long long)num_kb * it.part / p.split);
  it.kb1 = (int)((long long)num_kb * (it.part + 1) / p.split);
  return it;
}

// Per-warp view of the tile schedule.  Static: item = unit, unit + num_units, ...  Dynamic: items arrive through a
// ring of SCHED_SLOTS shared-memory slots written (in both CTAs of a pair) by the scheduler warp of the leader CTA;
// every consumer warp reads each slot once and releases it on the LEADER's empty barrier.  -1 ends the loop.
struct TileSched {
  int slot;
  uint32_t phase;
  int w;                           // static mode: the next item of this unit
  // bar_sfull: own CTA's full[] (shared::cta); sempty_leader: the leader's empty[] (shared::cluster when pair)
  // gate != 0 (the leader's producer warp only): barrier to arrive on after taking an item — the scheduler draws the
  // next item only then, so a unit never claims work more than one item ahead of its producer
  __device__ __forceinline__ int next(int lane, bool dynamic, bool pair, uint32_t bar_sfull, uint32_t sempty_leader,
                                      const volatile int* items, int step, int num_items, uint32_t gate = 0) {
    if (!dynamic) {
      const int r = w;
      w += step;
      return r < num_items ? r : -1;
    }
    mbar_wait_cluster(bar_sfull + 8 * slot, phase);
    const int r = items[slot];
    __syncwarp();
    if (lane == 0) {
      if (pair) mbar_arrive_cluster(sempty_leader + 8 * slot);
      else mbar_arrive(sempty_leader + 8 * slot);
      if (gate) mbar_arrive(gate);
    }
    if (++slot == 8) { slot = 0; phase ^= 1; }
    return r;
  }
};

template <typename OutT> struct OutPack;
template <> struct OutPack<float> {
  static constexpr int COLS = 32;   // accumulator columns per 128-byte staging row
  __device__ static void pack(const uint32_t (&a)[32], const uint32_t (&)[32], uint32_t (&w)[32]) {
#pragma unroll
    for (int i = 0; i < 32; i++) w[i] = a[i];
  }
};
template <> struct OutPack<int32_t> {
  static constexpr int COLS = 32;
  __device__ static void pack(const uint32_t (&a)[32], const uint32_t (&)[32], uint32_t (&w)[32]) {
#pragma unroll
    for (int i = 0; i < 32; i++) w[i] = a[i];
  }
};
struct bf16_out {};   // tag: C stored as bf16 (RNE from the fp32 accumulator)
template <> struct OutPack<bf16_out> {
  static constexpr int COLS = 64;
  __device__ static uint32_t cvt2(uint32_t lo, uint32_t hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
    return r;
  }
  __device__ static void pack(const uint32_t (&a)[32], const uint32_t (&b)[32], uint32_t (&w)[32]) {
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = cvt2(a[2 * i], a[2 * i + 1]);
#pragma unroll
    for (int i = 0; i < 16; i++) w[16 + i] = cvt2(b[2 * i], b[2 * i + 1]);
  }
};
struct s8_out {};     // tag: C stored as int8 through requant_s8 (TcParams::row_max = per-row scale,
                      // TcParams::col_max = per-row bias or null); int8 kernels only
template <> struct OutPack<s8_out> {
  static constexpr int COLS = 32;
  __device__ static void pack(const uint32_t (&)[32], const uint32_t (&)[32], uint32_t (&)[32]) {}
};
template <typename OutT> struct OutBytes { static constexpr int V = 4; };
template <> struct OutBytes<bf16_out> { static constexpr int V = 2; };
template <> struct OutBytes<s8_out> { static constexpr int V = 1; };

// The scheduler warp of the leader CTA: draws work items from the global counter and publishes each one in a slot of
// BOTH CTAs of the pair (remote shared-memory store + remote barrier arrive); -1 after the last item.
template <int CG, int SLOTS>
__device__ __forceinline__ void run_tile_scheduler(const TcParams& p, int lane, int num_items, int num_units,
                                                   uint32_t bar_sfull, uint32_t bar_sempty, uint32_t s_items, uint32_t bar_gate) {
  int slot = 0;
  uint32_t ph = 0, gph = 0;
  const int total = num_items + num_units;              // every unit draws exactly one sentinel
  for (int i = 0;; i++) {
    if (i > 0) {                                        // draw item i only once the producer has taken item i - 1:
      mbar_wait(bar_gate, gph);                         // work is claimed late, so late or slow units claim less
      gph ^= 1;
    }
    mbar_wait_cluster(bar_sempty + 8 * slot, ph ^ 1);
    int w = 0;
    if (lane == 0) {
      w = atomicAdd(p.sched_counter, 1);
      if (w == total - 1) atomicExch(p.sched_counter, 0);              // last draw of the launch: re-arm for a later one
      const uint32_t item = (uint32_t)(w < num_items ? w : -1);
#pragma unroll
      for (int r = 0; r < CG; r++) {
        st_shared_cluster_u32(mapa(s_items + 4 * slot, r), item);
        mbar_arrive_cluster(mapa(bar_sfull + 8 * slot, r));            // release.cluster: orders the store above
      }
    }
    w = __shfl_sync(0xffffffffu, w, 0);
    if (w >= num_items) break;
    if (++slot == SLOTS) { slot = 0; ph ^= 1; }
  }
}

template <int KIND, int BN, int STAGES, typename OutT, class Prod, int A_ROW_BYTES, int CG, int EPIW>
__global__ void __launch_bounds__((TcConfig<KIND, BN, STAGES, Prod, A_ROW_BYTES, CG, EPIW>::THREADS), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const TcParams p) {
  using Cfg = TcConfig<KIND, BN, STAGES, Prod, A_ROW_BYTES, CG, EPIW>;
  using T = KindTraits<KIND>;
  constexpr int OB = OutBytes<OutT>::V;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzle atoms need 1 KB
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + STAGES * Cfg::A_STAGE;
  const uint32_t sEpi = sB + STAGES * Cfg::B_STAGE;
  const uint32_t sBar = sEpi + Cfg::EPI_STAGING;
  const uint32_t bar_full = sBar;
  const uint32_t bar_empty = sBar + 8 * STAGES;
  const uint32_t bar_tfull = sBar + 16 * STAGES;
  const uint32_t bar_tempty = bar_tfull + 16;
  const uint32_t s_tmem_ptr = bar_tempty + 16;
  const uint32_t bar_sfull = s_tmem_ptr + 16;
  const uint32_t bar_sempty = bar_sfull + 8 * Cfg::SCHED_SLOTS;
  const uint32_t bar_gate = bar_sempty + 8 * Cfg::SCHED_SLOTS;
  const uint32_t s_items = bar_gate + 8;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CTA pair (CG == 2): rank 0 is the leader — it owns the full[] and tmem_empty[] barriers the whole
  // pair signals, and its MMA thread issues tcgen05.mma.cta_group::2 for both SMs.
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int unit = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int num_units = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; i++) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, Cfg::EPI_WARPS * CG);     // one arrive per epilogue warp of every CTA in the pair
    }
    for (int i = 0; i < Cfg::SCHED_SLOTS; i++) {
      mbar_init(bar_sfull + 8 * i, 1);                                  // the scheduler warp's arrive
      mbar_init(bar_sempty + 8 * i, CG * (1 + Cfg::EPI_WARPS) + 1);     // producer + epilogue warps of every CTA, + the MMA warp
    }
    mbar_init(bar_gate, 1);                                             // the leader's producer warp: "took an item"
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) tmem_alloc_cg2<Cfg::TMEM_COLS>(s_tmem_ptr);
    else tmem_alloc<Cfg::TMEM_COLS>(s_tmem_ptr);
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync(); else __syncthreads();   // peers touch each other's barriers
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (s_tmem_ptr - smem_base));
  // PDL: everything above (barriers, TMEM, descriptor prefetch) may overlap the tail of the previous kernel in
  // the stream; nothing below (operand loads, C stores) may.
  griddep_launch();
  griddep_wait();

  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_items = p.full_tiles + (num_tiles - p.full_tiles) * (p.halfn ? 2 : p.split);
  const int num_kb = (p.K + Cfg::BK - 1) / Cfg::BK;
  TileSched sched;
  sched.slot = 0; sched.phase = 0; sched.w = unit;
  const bool dyn = p.sched_counter != nullptr;
#define NEXT_ITEM_G(GATE) sched.next(lane, dyn, CG == 2, bar_sfull, CG == 2 ? mapa(bar_sempty, 0) : bar_sempty, \
                                     reinterpret_cast<const volatile int*>(smem_gen + (s_items - smem_base)), num_units, num_items, GATE)
#define NEXT_ITEM() NEXT_ITEM_G(0u)

  if (warp < Cfg::EPI_WARP0) {
  // register pool of the CTA = 384 threads x 168 (launch bound); afterwards 128 x 88 + 256 x 208 = the same 64512
  if constexpr (Cfg::REGACC) setmaxnreg_dec<88>();   // data-movement warpgroup (incl. its two idle warps) gives registers back
  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp walks the loop (warp-uniform control flow keeps stage/phase/coordinates in uniform
    // registers, so the UTMALDG operands need no per-lane R2UR waterfall); one elected lane issues.
    {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t gate = leader ? bar_gate : 0u;
      for (int w = NEXT_ITEM_G(gate); w >= 0; w = NEXT_ITEM_G(gate)) {
        const WorkItem it = work_item<BN>(w, p, num_kb);
        int mb, nb;
        tile_coords(it.tile, p.tiles_m, p.tiles_n, p.group_m, mb, nb);
        // this CTA's slice of the unit: its 128 rows of A, its BN/CG columns of B
        const int bn_cta = it.bn / CG;
        const int boxes = bn_cta / Cfg::B_BOX_COLS;       // B column blocks this CTA stages per plane
        const uint32_t tx_bytes = CG * (Cfg::A_STAGE + Prod::NPB * boxes * Cfg::B_BOX_BYTES);
        const int m0 = mb * Cfg::TILE_M + (int)cta_rank * Cfg::BM, n0 = nb * BN + it.nsub * it.bn + (int)cta_rank * bn_cta;
        for (int kb = it.kb0; kb < it.kb1; kb++) {
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          // bytes from both CTAs complete on the LEADER's full barrier; only the leader arms it
          const uint32_t full = CG == 2 ? mapa(bar_full + 8 * s, 0) : bar_full + 8 * s;
          if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(bar_full + 8 * s, tx_bytes);
#pragma unroll
          for (int pa = 0; pa < Prod::NPA; pa++) {
            const uint32_t dst = sA + s * Cfg::A_STAGE + pa * Cfg::A_PLANE;
            if constexpr (CG == 2) tma_load_2d_cg2(dst, &tmA, full, kb * Cfg::BK, pa * p.a_plane_rows + m0);
            else tma_load_2d(dst, &tmA, full, kb * Cfg::BK, pa * p.a_plane_rows + m0);
          }
#pragma unroll
          for (int pb = 0; pb < Prod::NPB; pb++)
#pragma unroll
            for (int j = 0; j < Cfg::B_BOXES; j++) {
              if (j >= boxes) break;
              const uint32_t dst = sB + s * Cfg::B_STAGE + pb * Cfg::B_PLANE + j * Cfg::B_BOX_BYTES;
              if constexpr (CG == 2)
                tma_load_2d_cg2(dst, &tmB, full, n0 + j * Cfg::B_BOX_COLS, pb * p.b_plane_rows + kb * Cfg::BK);
              else
                tma_load_2d(dst, &tmB, full, n0 + j * Cfg::B_BOX_COLS, pb * p.b_plane_rows + kb * Cfg::BK);
            }
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Same shape as the producer: warp-uniform loop, descriptors live in uniform registers, one elected
    // lane issues the MMAs and the commits of a k-block.
    if (leader) {
      const uint32_t b_lbo = p.dbg_b_lbo ? (uint32_t)p.dbg_b_lbo : (uint32_t)Cfg::B_BOX_BYTES;
      const uint32_t b_sbo = p.dbg_b_sbo ? (uint32_t)p.dbg_b_sbo : (uint32_t)T::B_SBO;
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int w = NEXT_ITEM(); w >= 0; w = NEXT_ITEM()) {
       const WorkItem it = work_item<BN>(w, p, num_kb);
       const uint32_t idesc = make_idesc(T::C_FMT, T::AB_FMT, /*a_mn=*/0, /*b_mn=*/1, Cfg::TILE_M, it.bn);
       for (int c0 = it.kb0; c0 < it.kb1; c0 += p.chunk_kb) {     // one TMEM accumulator per K-chunk
        mbar_wait(bar_tempty + 8 * as, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * Cfg::ACC_STRIDE;
        const int c1 = min(c0 + p.chunk_kb, it.kb1);
        for (int kb = c0; kb < c1; kb++) {
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          const uint32_t a0 = sA + s * Cfg::A_STAGE;
          const uint32_t b0 = sB + s * Cfg::B_STAGE;
          if (elect_one()) {
#pragma unroll
          for (int pr = 0; pr < Prod::N; pr++) {
#pragma unroll
            for (int k = 0; k < Cfg::MMAS_PER_STAGE; k++) {
              const uint64_t ad = make_sdesc(a0 + Prod::ia(pr) * Cfg::A_PLANE + k * Cfg::A_KADV, 16,
                                             Cfg::A_SBO, Cfg::A_LAYOUT);
              const uint64_t bd = make_sdesc(b0 + Prod::ib(pr) * Cfg::B_PLANE + k * Cfg::B_KADV, b_lbo,
                                             b_sbo, T::B_LAYOUT);
              if constexpr (CG == 2) tc_mma_cg2<KIND>(d_tmem, ad, bd, idesc, ((kb - c0) | k | pr) != 0 ? 1u : 0u);
              else tc_mma<KIND>(d_tmem, ad, bd, idesc, ((kb - c0) | k | pr) != 0 ? 1u : 0u);
            }
          }
          // frees the smem slot (in both CTAs of a pair) when these MMAs retire
          if constexpr (CG == 2) tc_commit_cg2(bar_empty + 8 * s, 3); else tc_commit(bar_empty + 8 * s);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if (elect_one()) {
          if constexpr (CG == 2) tc_commit_cg2(bar_tfull + 8 * as, 3); else tc_commit(bar_tfull + 8 * as);
        }
        __syncwarp();
        if (++as == 2) { as = 0; aph ^= 1; }
       }
      }
    }
  } else if (Cfg::REGACC && warp == Cfg::SCHED_WARP) {
    if (dyn && leader)
      run_tile_scheduler<CG, Cfg::SCHED_SLOTS>(p, lane, num_items, num_units, bar_sfull, bar_sempty, s_items, bar_gate);
  }
  } else if (!Cfg::REGACC && warp == Cfg::SCHED_WARP) {
    if (dyn && leader)
      run_tile_scheduler<CG, Cfg::SCHED_SLOTS>(p, lane, num_items, num_units, bar_sfull, bar_sempty, s_items, bar_gate);
  } else {
   if constexpr (Cfg::REGACC) {
    setmaxnreg_inc<208>();
    // ===================== epilogue, register-resident two-level accumulation (8 warps) =====================
    // The tensor core adds into its fp32 accumulator with truncation, so a long K chain drifts (measured:
    // error grows ~K).  K is cut into chunks; each chunk gets a fresh TMEM accumulator and is added HERE,
    // with a rounded fp32 add, to the tile's running sum, which lives in registers: lane = row (TMEM lane
    // quadrant warp % 4), warps 4-7 own tile columns [0, bn/2), warps 8-11 own [bn/2, bn).  C is written
    // once per tile (read once more only for C += A*B and for the K-split tail parts).
    const int ew = warp - Cfg::EPI_WARP0;
    const int q = ew & 3;
    const int half = ew >> 2;
    uint8_t* stg = smem_gen + (sEpi - smem_base) + ew * 4096;   // 32 rows x 128 B, chunk-swizzled
    constexpr int HC = BN / 2;                       // columns per warp in a full-width tile
    constexpr int NJ = HC / 32;                      // 32-column register groups
    int as = 0;
    uint32_t aph = 0;
    const uint32_t tempty_base = CG == 2 ? mapa(bar_tempty, 0) : bar_tempty;   // leader's barrier
    for (int w = NEXT_ITEM(); w >= 0; w = NEXT_ITEM()) {
      const WorkItem it = work_item<BN>(w, p, num_kb);
      int mb, nb;
      tile_coords(it.tile, p.tiles_m, p.tiles_n, p.group_m, mb, nb);
      const int hc = it.bn / 2;                      // columns this warp owns (HC, or HC/2 in a half-width tile)
      const int nj = hc / 32;
      const int m0 = mb * Cfg::TILE_M + (int)cta_rank * Cfg::BM + q * 32;
      const int n0 = nb * BN + it.nsub * it.bn + half * hc;
      const uint32_t t_col = (uint32_t)(half * hc);
      float acc[NJ][32];
      const int my_re = (p.row_max != nullptr && m0 + lane < p.M) ? pow2_exp(__ldg(p.row_max + m0 + lane)) : 0;
      bool first = true;
      for (int c0 = it.kb0; c0 < it.kb1; c0 += p.chunk_kb) {
        mbar_wait(bar_tfull + 8 * as, aph);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + as * Cfg::ACC_STRIDE + t_col;
#pragma unroll
        for (int j = 0; j < NJ; j++) {
          if (j < nj) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_addr + j * 32, r);
            tmem_ld_wait();
            if (first) {
#pragma unroll
              for (int i = 0; i < 32; i++) acc[j][i] = __uint_as_float(r[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; i++) acc[j][i] = __fadd_rn(acc[j][i], __uint_as_float(r[i]));
            }
          }
        }
        first = false;
        tc_fence_before();                            // accumulator drained: hand the TMEM stage back
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_cluster(tempty_base + 8 * as);
          else mbar_arrive(bar_tempty + 8 * as);
        }
        if (++as == 2) { as = 0; aph ^= 1; }
      }
      // ---- the tile's single pass over C ----
      int* flag = p.flags + ((it.tile - p.full_tiles) * CG + (int)cta_rank) * Cfg::EPI_WARPS + ew;
      if (it.part > 0) {                               // K-split tail: wait until parts < it.part are in C
        if (lane == 0) {
          const long long t0 = clock64();
          while (true) {
            int v;
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if (v == it.part) break;
            if (clock64() - t0 > 4000000000LL) { asm volatile("trap;"); }
          }
        }
        __syncwarp();
      }
      // beta * C is read by part 0 only; later K parts add alpha * partial to what is already there
      const bool fold = p.accumulate != 0 || it.part > 0 || (p.axpby && p.beta != 0.f);
      const float al = p.axpby ? p.alpha : 1.f;
      const float be = (p.axpby && it.part == 0) ? p.beta : 1.f;
      const int chunk = lane & 7;
      // (ncu, round 2: the first version fetched the row maximum inside the store loop — 32 dependent L2 round trips per
      //  tile — and unrolled that loop 32 times into 150 KB of code: stall_long_sb + stall_no_inst, 13-19 us per tile.)
      // my_re: scale exponent of the row this lane OWNS (row = lane), one coalesced load per tile issued before the
      // accumulation loop's results are needed; the store loop gets row exponents by shuffle.
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        if (j < nj) {
          const int col = n0 + j * 32 + chunk * 4;
          const bool vec = p.vec_ok && col + 4 <= p.N;
          int ce[4] = {0, 0, 0, 0};
          if (p.col_max != nullptr) {                 // issued before the staging round trip: latency overlaps it
#pragma unroll
            for (int e = 0; e < 4; e++)
              if (col + e < p.N) ce[e] = pow2_exp(__ldg(p.col_max + col + e));
          }
          if (fold && vec) {                          // C tile lines of this group towards L1 while the transpose runs
#pragma unroll
            for (int i = 0; i < 8; i++) {
              const int gm = m0 + i * 4 + (lane >> 3);
              if (gm < p.M) asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const float*>(p.C) + (long long)gm * p.ldc + col));
            }
          }
          // registers (row = lane) -> staging, 16-byte chunk index XOR (row & 7): conflict-free both ways
#pragma unroll
          for (int g = 0; g < 8; g++) {
            uint4 v = make_uint4(__float_as_uint(acc[j][4 * g]), __float_as_uint(acc[j][4 * g + 1]),
                                 __float_as_uint(acc[j][4 * g + 2]), __float_as_uint(acc[j][4 * g + 3]));
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((g ^ (lane & 7)) << 4)) = v;
          }
          __syncwarp();
#pragma unroll 1
          for (int i = 0; i < 8; i++) {
            const int row = i * 4 + (lane >> 3);
            const int gm = m0 + row;
            const int re = __shfl_sync(0xffffffffu, my_re, row);
            float4 v = *reinterpret_cast<const float4*>(stg + row * 128 + ((chunk ^ (row & 7)) << 4));
            if (gm < p.M) {
              if (p.row_max != nullptr) {                // undo the operand scaling: exact powers of two
                v.x = mul_pow2(v.x, re + ce[0]); v.y = mul_pow2(v.y, re + ce[1]);
                v.z = mul_pow2(v.z, re + ce[2]); v.w = mul_pow2(v.w, re + ce[3]);
              }
              float* dst = reinterpret_cast<float*>(p.C) + (long long)gm * p.ldc + col;
              if (p.axpby) { v.x *= al; v.y *= al; v.z *= al; v.w *= al; }
              if (vec) {
                if (fold) {                              // once per tile: C += A*B, beta * C, or an earlier K part
                  // parts of a K-split tile were written by another SM in this launch: read them at L2, not through L1
                  const float4 o = it.part > 0 ? __ldcg(reinterpret_cast<const float4*>(dst)) : *reinterpret_cast<const float4*>(dst);
                  if (p.axpby) { v.x = fmaf(be, o.x, v.x); v.y = fmaf(be, o.y, v.y); v.z = fmaf(be, o.z, v.z); v.w = fmaf(be, o.w, v.w); }
                  else { v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                }
                if (p.stream_c) __stcs(reinterpret_cast<float4*>(dst), v);
                else *reinterpret_cast<float4*>(dst) = v;
              } else {
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; e++)
                  if (col + e < p.N) dst[e] = !fold ? vv[e] : p.axpby ? fmaf(be, __ldcg(dst + e), vv[e]) : vv[e] + __ldcg(dst + e);
              }
            }
          }
          __syncwarp();
        }
      }
      if (w >= p.full_tiles && p.split > 1) {          // publish this part (the last one re-arms the flag)
        __threadfence();
        __syncwarp();
        if (lane == 0) {
          const int nv = it.part + 1 == p.split ? 0 : it.part + 1;
          asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(nv) : "memory");
        }
      }
    }
   } else {
    // ===================== epilogue (4 or 8 warps) =====================
    const int q = warp & 3;                          // TMEM lane quadrant this warp may read
    const int ew = warp - Cfg::EPI_WARP0;
    const int ehalf = ew >> 2;                       // 8 warps: which half of the column passes this warp drains
    uint8_t* stg = smem_gen + (sEpi - smem_base) + ew * 4096;   // 32 rows x 128 B, chunk-swizzled
    constexpr int COLS = OutPack<OutT>::COLS;
    constexpr int VEC_ELEMS = 16 / OB;
    int as = 0;
    uint32_t aph = 0;
    const uint32_t tempty_base = CG == 2 ? mapa(bar_tempty, 0) : bar_tempty;   // leader's barrier
    for (int w = NEXT_ITEM(); w >= 0; w = NEXT_ITEM()) {
      const WorkItem it = work_item<BN>(w, p, num_kb);
      int mb, nb;
      tile_coords(it.tile, p.tiles_m, p.tiles_n, p.group_m, mb, nb);
      const int m0 = mb * Cfg::TILE_M + (int)cta_rank * Cfg::BM + q * 32, n0 = nb * BN + it.nsub * it.bn;
      const int passes = it.bn / COLS;
      // requant: this lane's row keeps one scale / bias for the whole tile (fetched before the wait on the accumulator)
      float sc = 0.0f, bi = 0.0f;
      bool hb = false;
      if constexpr (std::is_same<OutT, s8_out>::value) {
        hb = p.col_max != nullptr;
        if (m0 + lane < p.M) {
          sc = __ldg(p.row_max + m0 + lane);
          if (hb) bi = __ldg(p.col_max + m0 + lane);
        }
      }
      int* flag = p.flags + ((it.tile - p.full_tiles) * CG + (int)cta_rank) * Cfg::EPI_WARPS + ew;
      if (it.part > 0) {                               // wait until parts < it.part are in C
        if (lane == 0) {
          const long long t0 = clock64();
          while (true) {
            int v;
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if (v == it.part) break;
            if (clock64() - t0 > 4000000000LL) { asm volatile("trap;"); }
          }
        }
        __syncwarp();
      }
     // Two-level accumulation (split-precision modes): the tensor core adds into its fp32
     // accumulator with truncation, so a long K chain drifts (measured: error grows ~K).  Each
     // K-chunk gets a fresh TMEM accumulator and is folded into C here with a rounded fp32 add.
     for (int c0 = it.kb0; c0 < it.kb1; c0 += p.chunk_kb) {
      const bool fold = p.accumulate != 0 || c0 != it.kb0 || it.part > 0 || (p.axpby && p.beta != 0.f);
      const float al = p.axpby ? p.alpha : 1.f;
      const float be = (p.axpby && it.part == 0 && c0 == it.kb0) ? p.beta : 1.f;
      mbar_wait(bar_tfull + 8 * as, aph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + as * Cfg::ACC_STRIDE;
#pragma unroll 1
      const int ps0 = Cfg::EPI_WARPS == 8 ? ehalf * (passes >> 1) : 0;
      const int ps1 = Cfg::EPI_WARPS == 8 ? (ehalf == 0 ? (passes >> 1) : passes) : passes;
      if (ps0 == ps1) {                              // (8 warps, a single-pass tile) nothing to drain here: still release the stage
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_cluster(tempty_base + 8 * as);
          else mbar_arrive(bar_tempty + 8 * as);
        }
      }
      for (int ps = ps0; ps < ps1; ps++) {
       if constexpr (std::is_same<OutT, s8_out>::value) {
        // Requantising epilogue: lane = row, 32 accumulator columns -> 32 bytes, stored straight from
        // registers as two 16-byte vectors (whole 32-byte sectors; no staging transpose needed).
        uint32_t ra[32];
        tmem_ld_32x32b_x32(t_addr + ps * 32, ra);
        tmem_ld_wait();
        if (ps == ps1 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2) mbar_arrive_cluster(tempty_base + 8 * as);
            else mbar_arrive(bar_tempty + 8 * as);
          }
        }
        const int gm = m0 + lane, col0 = n0 + ps * 32;
        if (gm < p.M && col0 < p.N) {
          uint32_t w8[8];
#pragma unroll
          for (int j = 0; j < 8; j++) {
            uint32_t v = 0;
#pragma unroll
            for (int e = 0; e < 4; e++)
              v |= ((uint32_t)requant_s8((int32_t)ra[4 * j + e], sc, bi, hb) & 0xFFu) << (8 * e);
            w8[j] = v;
          }
          uint8_t* dst = reinterpret_cast<uint8_t*>(p.C) + (long long)gm * p.ldc + col0;
          if (p.vec_ok && col0 + 32 <= p.N) {
            reinterpret_cast<uint4*>(dst)[0] = make_uint4(w8[0], w8[1], w8[2], w8[3]);
            reinterpret_cast<uint4*>(dst)[1] = make_uint4(w8[4], w8[5], w8[6], w8[7]);
          } else {
#pragma unroll
            for (int e = 0; e < 32; e++)
              if (col0 + e < p.N) dst[e] = (uint8_t)(w8[e >> 2] >> (8 * (e & 3)));
          }
        }
       } else {
        const int chunk = lane & 7;
        const int col = n0 + ps * COLS + chunk * VEC_ELEMS;
        const bool vec = p.vec_ok && col + VEC_ELEMS <= p.N;
        // folding pass: fetch all eight partial-C vectors first, so their L2 latency overlaps the
        // TMEM load and the staging transpose below
        float4 old[8];
        if constexpr (OB == 4) {
          if (fold && vec) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
              const int gm = m0 + i * 4 + (lane >> 3);
              old[i] = gm < p.M ? __ldcg(reinterpret_cast<const float4*>(
                                      reinterpret_cast<const float*>(p.C) + (long long)gm * p.ldc + col))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }
        uint32_t ra[32], rb[32], w[32];
        tmem_ld_32x32b_x32(t_addr + ps * (COLS == 64 ? 64 : 32), ra);
        if constexpr (COLS == 64) tmem_ld_32x32b_x32(t_addr + ps * 64 + 32, rb);
        tmem_ld_wait();
        if (ps == ps1 - 1) {                         // this warp's share of the TMEM stage is drained: hand it back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2) mbar_arrive_cluster(tempty_base + 8 * as);
            else mbar_arrive(bar_tempty + 8 * as);
          }
        }
        OutPack<OutT>::pack(ra, rb, w);
        if (p.epi_direct && !fold && !p.axpby && p.row_max == nullptr && p.vec_ok && n0 + (ps + 1) * COLS <= p.N) {
          // Direct epilogue: lane = row holds 128 contiguous bytes of C for this pass; eight 16-byte stores
          // straight from registers (32 rows per instruction, whole lines after the eighth) — no staging
          // round trip, no warp syncs.
          const int gm = m0 + lane;
          if (gm < p.M) {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.C) +
                                                  ((long long)gm * p.ldc + n0 + ps * COLS) * OB);
#pragma unroll
            for (int j = 0; j < 8; j++) dst[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
          }
          continue;
        }
        // registers (row = lane) -> staging, 16-byte chunk index XOR (row & 7): conflict-free both ways
#pragma unroll
        for (int j = 0; j < 8; j++) {
          uint4 v = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = v;
        }
        __syncwarp();
        // staging -> global: 8 lanes cover one 128-byte row segment, 4 rows per instruction
        // (operand scalings exist only in the split modes, whose tiles take the register-accumulation epilogue above)
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int row = i * 4 + (lane >> 3);
          const int gm = m0 + row;
          uint4 v = *reinterpret_cast<const uint4*>(stg + row * 128 + ((chunk ^ (row & 7)) << 4));
          if (gm < p.M) {
            uint8_t* dst = reinterpret_cast<uint8_t*>(p.C) + ((long long)gm * p.ldc + col) * OB;
            if (vec) {
              if constexpr (std::is_same<OutT, float>::value) {
                if (p.axpby) {
                  const float4 o = fold ? old[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                  v.x = __float_as_uint(fmaf(be, o.x, al * __uint_as_float(v.x)));
                  v.y = __float_as_uint(fmaf(be, o.y, al * __uint_as_float(v.y)));
                  v.z = __float_as_uint(fmaf(be, o.z, al * __uint_as_float(v.z)));
                  v.w = __float_as_uint(fmaf(be, o.w, al * __uint_as_float(v.w)));
                } else if (fold) {
                  v.x = __float_as_uint(__uint_as_float(v.x) + old[i].x);
                  v.y = __float_as_uint(__uint_as_float(v.y) + old[i].y);
                  v.z = __float_as_uint(__uint_as_float(v.z) + old[i].z);
                  v.w = __float_as_uint(__uint_as_float(v.w) + old[i].w);
                }
              } else if constexpr (std::is_same<OutT, int32_t>::value) {
                if (fold) {                           // exact: integer partial sums
                  v.x += __float_as_uint(old[i].x); v.y += __float_as_uint(old[i].y);
                  v.z += __float_as_uint(old[i].z); v.w += __float_as_uint(old[i].w);
                }
              }
              *reinterpret_cast<uint4*>(dst) = v;
            } else {
              uint32_t vv[4] = {v.x, v.y, v.z, v.w};
              if constexpr (OB == 4) {
#pragma unroll
                for (int e = 0; e < 4; e++)
                  if (col + e < p.N) {
                    if constexpr (std::is_same<OutT, float>::value) {
                      if (p.axpby) vv[e] = __float_as_uint(fmaf(be, fold ? __ldcg(reinterpret_cast<const float*>(dst) + e) : 0.f, al * __uint_as_float(vv[e])));
                      else if (fold) vv[e] = __float_as_uint(__uint_as_float(vv[e]) + __ldcg(reinterpret_cast<const float*>(dst) + e));
                    } else if constexpr (std::is_same<OutT, int32_t>::value) {
                      if (fold) vv[e] += __ldcg(reinterpret_cast<const uint32_t*>(dst) + e);
                    }
                    reinterpret_cast<uint32_t*>(dst)[e] = vv[e];
                  }
              } else {
#pragma unroll
                for (int e = 0; e < 8; e++)
                  if (col + e < p.N)
                    reinterpret_cast<uint16_t*>(dst)[e] = (uint16_t)(vv[e >> 1] >> ((e & 1) * 16));
              }
            }
          }
        }
        __syncwarp();
       }
      }
      if (++as == 2) { as = 0; aph ^= 1; }
     }
      if (w >= p.full_tiles && p.split > 1) {          // publish this part (the last one re-arms the flag)
        __threadfence();
        __syncwarp();
        if (lane == 0) {
          const int nv = it.part + 1 == p.split ? 0 : it.part + 1;
          asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(nv) : "memory");
        }
      }
    }
  }

  }

#undef NEXT_ITEM
#undef NEXT_ITEM_G
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync(); else __syncthreads();   // no CTA may exit while its peer still signals it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_cg2<Cfg::TMEM_COLS>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---- fp32 -> stacked bf16 planes (the split-precision pre-pass) -----------------------------------
// src: rows x cols fp32, pitch ld.  dst: NP planes stacked along rows, plane p at row p*plane_rows,
// pitch dld (elements, multiple of 8).  x = p1 + p2 + p3 with p1 = bf16(x), p2 = bf16(x - p1),
// p3 = bf16(x - p1 - p2); the subtractions are exact in fp32.  Rows [rows, plane_rows) and columns
// [cols, dld) of every plane are written as zero (K padding of B must contribute nothing).
struct SplitJob {
  const float* src; long long ld; int rows, cols;
  uint16_t* dst; long long dld; int plane_rows;
};

// One launch splits both operands (blockIdx.z picks A or B).  A thread owns 8 consecutive columns
// (two 16-byte loads, one 16-byte store per plane) and walks rows blockIdx.y, +gridDim.y, ... two at
// a time so that 64 B of loads are in flight per thread; HBM-bound: 4 + 2*NP bytes per element.
template <int NP>
__global__ void __launch_bounds__(256) split_planes_kernel(const SplitJob ja, const SplitJob jb) {
  griddep_launch();
  griddep_wait();
  const SplitJob& jo = blockIdx.z == 0 ? ja : jb;
  const float* __restrict__ src = jo.src;
  uint16_t* __restrict__ dst = jo.dst;
  const long long ld = jo.ld, dld = jo.dld;
  const int rows = jo.rows, cols = jo.cols, plane_rows = jo.plane_rows;
  const int c = (int)(blockIdx.x * 256 + threadIdx.x) * 8;
  if (c >= dld) return;
  const bool vec = c + 8 <= cols && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  for (int r0 = blockIdx.y * 2; r0 < plane_rows; r0 += gridDim.y * 2) {
    float x[2][8];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int r = r0 + u;
#pragma unroll
      for (int e = 0; e < 8; e++) x[u][e] = 0.f;
      if (r < rows) {
        const float* s = src + (long long)r * ld + c;
        if (vec) {
          const float4 v0 = __ldcs(reinterpret_cast<const float4*>(s));
          const float4 v1 = __ldcs(reinterpret_cast<const float4*>(s) + 1);
          x[u][0] = v0.x; x[u][1] = v0.y; x[u][2] = v0.z; x[u][3] = v0.w;
          x[u][4] = v1.x; x[u][5] = v1.y; x[u][6] = v1.z; x[u][7] = v1.w;
        } else {
#pragma unroll
          for (int e = 0; e < 8; e++)
            if (c + e < cols) x[u][e] = s[e];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int r = r0 + u;
      if (r >= plane_rows) break;
#pragma unroll
      for (int pl = 0; pl < NP; pl++) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          // packs {hi half <- x[2e+1], lo half <- x[2e]}, round-to-nearest-even
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[e]) : "f"(x[u][2 * e + 1]), "f"(x[u][2 * e]));
          x[u][2 * e] -= __uint_as_float(w[e] << 16);
          x[u][2 * e + 1] -= __uint_as_float(w[e] & 0xFFFF0000u);
        }
        *reinterpret_cast<uint4*>(dst + ((long long)pl * plane_rows + r) * dld + c) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
}


// ---- scaled fp16 split (B200_F32_F16X2) ------------------------------------------------------------
// x' = x * 2^-e (e from the row maximum of A / the column maximum of B, so |x'| < 1), then
// x' = h1 + h2 with h1 = fp16(x'), h2 = fp16(x' - h1): 22 significant bits; h2 may be an fp16
// subnormal (absolute quantum 2^-24 relative to the row/column scale).  Three launches per GEMM:
// rows of A (maximum, scaling and split fused: each row is read from HBM once), column maxima of B,
// columns of B.  All HBM-bound: 4 bytes read + 4 bytes written per element (+ 4 read for B's maxima,
// mostly L2 hits on the second touch).
__device__ __forceinline__ void split_f16x8(const float (&x)[8], uint4& h1, uint4& h2) {
  uint32_t a[4], b[4];
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const __half2 p = __floats2half2_rn(x[2 * e], x[2 * e + 1]);
    const float2 pf = __half22float2(p);
    const __half2 q = __floats2half2_rn(x[2 * e] - pf.x, x[2 * e + 1] - pf.y);
    a[e] = *reinterpret_cast<const uint32_t*>(&p);
    b[e] = *reinterpret_cast<const uint32_t*>(&q);
  }
  h1 = make_uint4(a[0], a[1], a[2], a[3]);
  h2 = make_uint4(b[0], b[1], b[2], b[3]);
}
__device__ __forceinline__ void load8(const float* __restrict__ s, int c, int cols, bool vec, float (&x)[8]) {
  if (vec && c + 8 <= cols) {
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(s + c));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(s + c) + 1);
    x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w;
    x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; e++) x[e] = c + e < cols ? __ldg(s + c + e) : 0.f;
  }
}

// A side: one warp per row.  VPL > 0: the row (cols <= 256 * VPL) stays in registers between the
// maximum and the split; VPL == 0: any length, second pass re-reads the row (L1 / L2 hits).
// dst: 2 planes stacked along rows (plane p at row p * plane_rows), pitch dld (multiple of 8); columns
// [cols, dld) and rows [rows, plane_rows) are written as zero.
template <int VPL>
__global__ void __launch_bounds__(256) split_f16_rows_kernel(const float* __restrict__ src, long long ld, int rows,
                                                             int cols, float* __restrict__ rmax,
                                                             uint16_t* __restrict__ dst, long long dld, int plane_rows) {
  griddep_launch();
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const bool vec = (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < plane_rows; r += nwarps) {
    uint16_t* d1 = dst + (long long)r * dld;
    uint16_t* d2 = dst + ((long long)plane_rows + r) * dld;
    if (r >= rows) {
      for (int c = lane * 8; c < dld; c += 256) {
        *reinterpret_cast<uint4*>(d1 + c) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(d2 + c) = make_uint4(0, 0, 0, 0);
      }
      continue;
    }
    const float* s = src + (long long)r * ld;
    float mx = 0.f;
    if constexpr (VPL > 0) {
      float x[VPL][8];
#pragma unroll
      for (int j = 0; j < VPL; j++) {
        const int c = (j * 32 + lane) * 8;
#pragma unroll
        for (int e = 0; e < 8; e++) x[j][e] = 0.f;
        if (c < cols) load8(s, c, cols, vec, x[j]);
#pragma unroll
        for (int e = 0; e < 8; e++) mx = fmaxf(mx, fabsf(x[j][e]));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (lane == 0) rmax[r] = mx;
      const int ex = -pow2_exp(mx);
#pragma unroll
      for (int j = 0; j < VPL; j++) {
        const int c = (j * 32 + lane) * 8;
        if (c < dld) {
#pragma unroll
          for (int e = 0; e < 8; e++) x[j][e] = mul_pow2(x[j][e], ex);
          uint4 h1, h2;
          split_f16x8(x[j], h1, h2);
          *reinterpret_cast<uint4*>(d1 + c) = h1;
          *reinterpret_cast<uint4*>(d2 + c) = h2;
        }
      }
    } else {
      for (int c0 = lane * 8; c0 < cols; c0 += 1024) {       // 4 vectors in flight per lane
        float x[4][8];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int c = c0 + u * 256;
#pragma unroll
          for (int e = 0; e < 8; e++) x[u][e] = 0.f;
          if (c < cols) load8(s, c, cols, vec, x[u]);
#pragma unroll
          for (int e = 0; e < 8; e++) mx = fmaxf(mx, fabsf(x[u][e]));
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (lane == 0) rmax[r] = mx;
      const int ex = -pow2_exp(mx);
      for (int c0 = lane * 8; c0 < dld; c0 += 512) {
        float x[2][8];
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const int c = c0 + u * 256;
#pragma unroll
          for (int e = 0; e < 8; e++) x[u][e] = 0.f;
          if (c < cols) load8(s, c, cols, vec, x[u]);
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const int c = c0 + u * 256;
          if (c < dld) {
#pragma unroll
            for (int e = 0; e < 8; e++) x[u][e] = mul_pow2(x[u][e], ex);
            uint4 h1, h2;
            split_f16x8(x[u], h1, h2);
            *reinterpret_cast<uint4*>(d1 + c) = h1;
            *reinterpret_cast<uint4*>(d2 + c) = h2;
          }
        }
      }
    }
  }
}

// Column maxima of B into out[cols] (zero on entry; non-negative floats order like their bit patterns,
// so atomicMax on the uint view works).  A block walks 16 rows of a 1024-column strip: every row read is one
// contiguous 4 KB segment (DRAM page locality; the first version read 512-byte pieces of eight rows at a time
// and reached 3 TB/s), a thread keeps its 4 columns' maxima in registers, no cross-thread reduction.
__global__ void __launch_bounds__(256) col_absmax_kernel(const float* __restrict__ src, long long ld, int rows,
                                                         int cols, unsigned int* __restrict__ out) {
  constexpr int ROWS = 16;      // all 16 row loads of a thread in flight at once (ncu: 32 rows in 4 batches of 8 reached 3 TB/s)
  griddep_launch();
  griddep_wait();
  const int c = blockIdx.x * 1024 + threadIdx.x * 4;
  if (c >= cols) return;
  const int r0 = blockIdx.y * ROWS, r1 = min(r0 + ROWS, rows);
  const bool vec = (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && c + 4 <= cols;
  float4 mx = make_float4(0.f, 0.f, 0.f, 0.f);
  if (vec) {
#pragma unroll 16
    for (int r = r0; r < r1; r++) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + (long long)r * ld + c));
      mx.x = fmaxf(mx.x, fabsf(v.x)); mx.y = fmaxf(mx.y, fabsf(v.y));
      mx.z = fmaxf(mx.z, fabsf(v.z)); mx.w = fmaxf(mx.w, fabsf(v.w));
    }
  } else {
    for (int r = r0; r < r1; r++) {
      const float* s = src + (long long)r * ld + c;
      mx.x = fmaxf(mx.x, fabsf(s[0]));
      if (c + 1 < cols) mx.y = fmaxf(mx.y, fabsf(s[1]));
      if (c + 2 < cols) mx.z = fmaxf(mx.z, fabsf(s[2]));
      if (c + 3 < cols) mx.w = fmaxf(mx.w, fabsf(s[3]));
    }
  }
  // Only a block that can raise the running maximum touches it: hundreds of row blocks hit the same 4 addresses and
  // same-address atomics serialise in L2 (ncu: this kernel sat at 37 % of peak DRAM throughput); after the first few
  // blocks almost every candidate is below the current value.  A stale (smaller) value read here only costs an atomic.
  const float m4[4] = {mx.x, mx.y, mx.z, mx.w};
#pragma unroll
  for (int e = 0; e < 4; e++)
    if (c + e < cols && __float_as_uint(m4[e]) > __ldcg(out + c + e)) atomicMax(out + c + e, __float_as_uint(m4[e]));
}

// B side: a thread owns 8 consecutive columns (their scale exponents live in registers) and walks rows
// blockIdx.y, +gridDim.y, ... two at a time.  Also re-zeroes `zero_buf` (the idle half of the double-
// buffered column maxima) for the next call.
__global__ void __launch_bounds__(256) split_f16_cols_kernel(const float* __restrict__ src, long long ld, int rows,
                                                             int cols, const float* __restrict__ cmax,
                                                             uint16_t* __restrict__ dst, long long dld, int plane_rows,
                                                             float* __restrict__ zero_buf, int zero_n) {
  griddep_launch();
  griddep_wait();
  const int c = (int)(blockIdx.x * 256 + threadIdx.x) * 8;
  if (blockIdx.y == 0 && zero_buf != nullptr) {
#pragma unroll
    for (int e = 0; e < 8; e++)
      if (c + e < zero_n) zero_buf[c + e] = 0.f;
  }
  if (c >= dld) return;
  const bool vec = (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  int ex[8];
#pragma unroll
  for (int e = 0; e < 8; e++) ex[e] = c + e < cols ? -pow2_exp(__ldg(cmax + c + e)) : 0;
  for (int r0 = blockIdx.y * 2; r0 < plane_rows; r0 += gridDim.y * 2) {
    float x[2][8];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int r = r0 + u;
#pragma unroll
      for (int e = 0; e < 8; e++) x[u][e] = 0.f;
      if (r < rows) load8(src + (long long)r * ld, c, cols, vec, x[u]);
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int r = r0 + u;
      if (r >= plane_rows) break;
#pragma unroll
      for (int e = 0; e < 8; e++) x[u][e] = mul_pow2(x[u][e], ex[e]);
      uint4 h1, h2;
      split_f16x8(x[u], h1, h2);
      *reinterpret_cast<uint4*>(dst + (long long)r * dld + c) = h1;
      *reinterpret_cast<uint4*>(dst + ((long long)plane_rows + r) * dld + c) = h2;
    }
  }
}

}  // namespace b200
