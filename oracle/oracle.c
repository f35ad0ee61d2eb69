/*
 * oracle.c — CPU restatement of the reference's GEMM hot path.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Parity status: PINNED against
 * oracle/_ref/libref.so (the reference's own sources) by
 * tests/test_oracle_vs_ref.py and the vectors under tests/golden/.
 *
 * Build: gcc -O3 -ffp-contract=off -mavx2 -mfma -pthread -fPIC -shared
 * (oracle/Makefile); explicit fmaf() calls become vfmadd, nothing else fuses.
 * -ffp-contract=off keeps multiply and add separately rounded, which is what
 * the reference's `g++ -O2` build of aarch64/REF_MMult.cpp produces on x86-64.
 */
#define _XOPEN_SOURCE 600
#include "oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ---- tiny row-parallel helper (this image's gcc has no libgomp) ------------- */
static int g_threads = 0;
void oracle_set_threads(int n) { g_threads = n; }
int oracle_get_threads(void) {
  if (g_threads > 0) return g_threads;
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}
typedef void (*row_fn)(int i0, int i1, void* ctx);
typedef struct { row_fn fn; void* ctx; int i0, i1; } row_job;
static void* row_tramp(void* p) { row_job* j = (row_job*)p; j->fn(j->i0, j->i1, j->ctx); return NULL; }
static void par_rows(int m, row_fn fn, void* ctx) {
  int nt = oracle_get_threads();
  if (nt > m) nt = m;
  if (nt <= 1) { fn(0, m, ctx); return; }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nt);
  row_job* jobs = (row_job*)malloc(sizeof(row_job) * nt);
  for (int t = 0; t < nt; t++) {
    jobs[t].fn = fn; jobs[t].ctx = ctx;
    jobs[t].i0 = (int)((long long)m * t / nt);
    jobs[t].i1 = (int)((long long)m * (t + 1) / nt);
    pthread_create(&th[t], NULL, row_tramp, &jobs[t]);
  }
  for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
}
typedef struct { int n, k; const void* a; int lda; const void* b; int ldb; void* c; int ldc; } mm_ctx;

void oracle_seed(long seed) { srand48(seed); }

/* cuda/random_matrix.cpp:3,6-16 — note the column-index-outer macro
 * A(i,j) = a[j*lda+i] and the double-precision expression around a float
 * cast of drand48(). */
void oracle_random_matrix_cuda(int m, int n, float* a, int lda) {
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++)
      a[(size_t)j * lda + i] = (float)(2.0 * (double)(float)drand48() - 1.0);
}

/* aarch64/random_matrix.cpp:16 */
void oracle_random_matrix_ones(int m, int n, float* a) {
  for (size_t i = 0; i < (size_t)m * n; i++) a[i] = 1.0f;
}

/* aarch64-int8/random_matrix.c:13-21 */
void oracle_random_int8_ramp(int m, int n, int8_t* a, int lda) {
  int val = 0;
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      a[(size_t)i * lda + j] = (int8_t)(val % 3);
      val++;
    }
}

void oracle_random_int8_uniform(int m, int n, int8_t* a, int lda, uint64_t seed) {
  uint64_t s = seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      s = s * 6364136223846793005ull + 1442695040888963407ull;
      uint32_t r = (uint32_t)(s >> 33);
      a[(size_t)i * lda + j] = (int8_t)((int)(r % 255u) - 127);
    }
}

/* aarch64/REF_MMult.cpp:18-28 */
void oracle_ref_mmult_f32(int m, int n, int k, const float* a, int lda,
                          const float* b, int ldb, float* c, int ldc) {
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++)
      for (int p = 0; p < k; p++)
        c[(size_t)i * ldc + j] =
            c[(size_t)i * ldc + j] + a[(size_t)i * lda + p] * b[(size_t)p * ldb + j];
}

/* aarch64/REF_MMult.cpp:24 as compiled by aarch64/makefile:14 (fused). */
void oracle_ref_mmult_f32_fma(int m, int n, int k, const float* a, int lda,
                              const float* b, int ldb, float* c, int ldc) {
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++)
      for (int p = 0; p < k; p++)
        c[(size_t)i * ldc + j] =
            fmaf(a[(size_t)i * lda + p], b[(size_t)p * ldb + j], c[(size_t)i * ldc + j]);
}

static void rows_f32_fma(int i0, int i1, void* vc) {
  mm_ctx* x = (mm_ctx*)vc;
  const float* a = (const float*)x->a; const float* b = (const float*)x->b; float* c = (float*)x->c;
  for (int i = i0; i < i1; i++) {
    float* __restrict ci = c + (size_t)i * x->ldc;
    for (int p = 0; p < x->k; p++) {
      const float aip = a[(size_t)i * x->lda + p];
      const float* __restrict bp = b + (size_t)p * x->ldb;
      for (int j = 0; j < x->n; j++) ci[j] = fmaf(aip, bp[j], ci[j]);
    }
  }
}
void oracle_ref_mmult_f32_fma_fast(int m, int n, int k, const float* a, int lda,
                                   const float* b, int ldb, float* c, int ldc) {
  mm_ctx x = {n, k, a, lda, b, ldb, c, ldc};
  par_rows(m, rows_f32_fma, &x);
}

/* Same per-element operation sequence (c_ij += a_ip*b_pj for p = 0..k-1, each
 * step rounded to fp32), reordered i,p,j so the inner loop is unit-stride. */
static void rows_f32(int i0, int i1, void* vc) {
  mm_ctx* x = (mm_ctx*)vc;
  const float* a = (const float*)x->a; const float* b = (const float*)x->b; float* c = (float*)x->c;
  for (int i = i0; i < i1; i++) {
    float* __restrict ci = c + (size_t)i * x->ldc;
    for (int p = 0; p < x->k; p++) {
      const float aip = a[(size_t)i * x->lda + p];
      const float* __restrict bp = b + (size_t)p * x->ldb;
      for (int j = 0; j < x->n; j++) ci[j] = ci[j] + aip * bp[j];
    }
  }
}
void oracle_ref_mmult_f32_fast(int m, int n, int k, const float* a, int lda,
                               const float* b, int ldb, float* c, int ldc) {
  mm_ctx x = {n, k, a, lda, b, ldb, c, ldc};
  par_rows(m, rows_f32, &x);
}

static void rows_f64acc(int i0, int i1, void* vc) {
  mm_ctx* x = (mm_ctx*)vc;
  const float* a = (const float*)x->a; const float* b = (const float*)x->b; double* c = (double*)x->c;
  for (int i = i0; i < i1; i++) {
    double* __restrict ci = c + (size_t)i * x->ldc;
    for (int j = 0; j < x->n; j++) ci[j] = 0.0;
    for (int p = 0; p < x->k; p++) {
      const double aip = (double)a[(size_t)i * x->lda + p];
      const float* __restrict bp = b + (size_t)p * x->ldb;
      for (int j = 0; j < x->n; j++) ci[j] += aip * (double)bp[j];
    }
  }
}
void oracle_ref_mmult_f64acc(int m, int n, int k, const float* a, int lda,
                             const float* b, int ldb, double* c, int ldc) {
  mm_ctx x = {n, k, a, lda, b, ldb, c, ldc};
  par_rows(m, rows_f64acc, &x);
}

/* aarch64-int8/REF_MMult.c:10-23 */
void oracle_ref_mmult_s8s32(int m, int n, int k, const int8_t* a, int lda,
                            const int8_t* b, int ldb, int32_t* c, int ldc) {
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++)
      for (int p = 0; p < k; p++)
        c[(size_t)i * ldc + j] =
            c[(size_t)i * ldc + j] + a[(size_t)i * lda + p] * b[(size_t)p * ldb + j];
}

/* Integer addition is associative, so any order is bit-identical. */
static void rows_s8(int i0, int i1, void* vc) {
  mm_ctx* x = (mm_ctx*)vc;
  const int8_t* a = (const int8_t*)x->a; const int8_t* b = (const int8_t*)x->b; int32_t* c = (int32_t*)x->c;
  for (int i = i0; i < i1; i++) {
    int32_t* __restrict ci = c + (size_t)i * x->ldc;
    for (int p = 0; p < x->k; p++) {
      const int32_t aip = a[(size_t)i * x->lda + p];
      const int8_t* __restrict bp = b + (size_t)p * x->ldb;
      for (int j = 0; j < x->n; j++) ci[j] += aip * (int32_t)bp[j];
    }
  }
}
void oracle_ref_mmult_s8s32_fast(int m, int n, int k, const int8_t* a, int lda,
                                 const int8_t* b, int ldb, int32_t* c, int ldc) {
  mm_ctx x = {n, k, a, lda, b, ldb, c, ldc};
  par_rows(m, rows_s8, &x);
}

/* aarch64-int8/int8kernel_m4.S:386-426.  One C statement per instruction; built with -ffp-contract=off so
 * the fmul / fadd pair stays two roundings as in the assembly. */
void oracle_requant_s32_to_s8(int m, int n, const int32_t* c, int ldc, const float* scales,
                              const float* bias, int8_t* out, int ldo) {
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      float f = (float)c[(size_t)i * ldc + j];          /* scvtf  :389 (round to nearest even) */
      f = f * scales[i];                                /* fmul   :394, scale of the ROW (v12.s[i]) */
      if (bias) f = f + bias[i];                        /* fadd   :405, skipped when bias == NULL :399-400 */
      int32_t r;                                        /* fcvtas :415 nearest, ties away, saturating, NaN -> 0 */
      if (f != f) r = 0;
      else {
        float t = roundf(f);
        r = t >= 2147483648.0f ? INT32_MAX : t <= -2147483648.0f ? INT32_MIN : (int32_t)t;
      }
      int16_t h = r > 32767 ? 32767 : r < -32768 ? -32768 : (int16_t)r;   /* sqxtn .4h :420 */
      out[(size_t)i * ldo + j] = h > 127 ? 127 : h < -128 ? -128 : (int8_t)h;   /* sqxtn .8b :425 */
    }
}

/* cuda/compare_matrices.cpp:17-29, NaN-aware. */
float oracle_compare_matrices_f32(int m, int n, const float* a, int lda,
                                  const float* b, int ldb) {
  float max_diff = 0.0f;
  int saw_nan = 0;
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      float d = a[(size_t)i * lda + j] - b[(size_t)i * ldb + j];
      if (d != d) saw_nan = 1;
      d = d < 0.0f ? -d : d;
      if (d > max_diff) max_diff = d;
    }
  return saw_nan ? NAN : max_diff;
}

/* aarch64-int8/compare_matrices.c:19-31 */
int32_t oracle_compare_matrices_s32(int m, int n, const int32_t* a, int lda,
                                    const int32_t* b, int ldb) {
  int64_t max_diff = 0;
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      int64_t d = (int64_t)a[(size_t)i * lda + j] - (int64_t)b[(size_t)i * ldb + j];
      if (d < 0) d = -d;
      if (d > max_diff) max_diff = d;
    }
  return max_diff > INT32_MAX ? INT32_MAX : (int32_t)max_diff;
}

float oracle_max_abs_f32(int m, int n, const float* a, int lda) {
  float mx = 0.0f;
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      float v = fabsf(a[(size_t)i * lda + j]);
      if (v != v) return NAN;
      if (v > mx) mx = v;
    }
  return mx;
}

double oracle_max_err_vs_f64(int m, int n, const float* c, int ldc,
                             const double* t, int ldt) {
  double mx = 0.0;
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      double d = fabs((double)c[(size_t)i * ldc + j] - t[(size_t)i * ldt + j]);
      if (d != d) return NAN;
      if (d > mx) mx = d;
    }
  return mx;
}

uint16_t oracle_f32_to_bf16(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x007FFFFFu))
    return (uint16_t)((u >> 16) | 0x0040u);             /* quiet NaN */
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7FFFu + lsb;                                    /* round to nearest even */
  return (uint16_t)(u >> 16);
}

float oracle_bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float x;
  memcpy(&x, &u, 4);
  return x;
}

void oracle_round_to_bf16_inplace(size_t count, float* a) {
  for (size_t i = 0; i < count; i++) a[i] = oracle_bf16_to_f32(oracle_f32_to_bf16(a[i]));
}

/* ---- MXFP4 (the 4-bit path; parity UNPINNED: the reference ships only "WIP", cuda-int4/README.md:1) --------
 * Restates the published OCP Microscaling Formats (MX) v1.0 specification: E2M1 elements
 * {0, .5, 1, 1.5, 2, 3, 4, 6} with sign, one UE8M0 scale 2^(e-127) per 32 consecutive elements,
 * shared exponent = floor(log2(max|x|)) - emax_elem (emax_elem = 2), elements rounded to nearest even and
 * saturated.  Nothing in /root/reference can pin this; the GPU path is checked against THIS restatement. */
static const float k_e2m1[8] = {0.f, .5f, 1.f, 1.5f, 2.f, 3.f, 4.f, 6.f};
uint8_t oracle_e2m1_encode(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  if (v != v) return 0;
  float a = fabsf(v);
  uint8_t code = a <= 0.25f ? 0 : a < 0.75f ? 1 : a <= 1.25f ? 2 : a < 1.75f ? 3 : a <= 2.5f ? 4 : a < 3.5f ? 5 : a <= 5.0f ? 6 : 7;
  return (uint8_t)(((u >> 31) << 3) | code);
}
float oracle_e2m1_decode(uint8_t c) { return (c & 8) ? -k_e2m1[c & 7] : k_e2m1[c & 7]; }
uint8_t oracle_ue8m0_from_max(float mx) {
  uint32_t u;
  memcpy(&u, &mx, 4);
  int ef = (int)((u >> 23) & 0xFF);
  if (ef == 255) return 255;
  int e = ef - 2;
  return (uint8_t)(e < 0 ? 0 : e);
}
double oracle_ue8m0_value(uint8_t e) { return e == 255 ? NAN : ldexp(1.0, (int)e - 127); }
/* rows x cols fp32 -> q (rows x kpad/2 bytes, element 2i in the low nibble) + sf (rows x kpad/32, plain row-major);
 * kpad = cols rounded up to 128, padding elements are zero. */
void oracle_mxf4_quantize(int rows, int cols, const float* src, int ld, uint8_t* q, uint8_t* sf) {
  int kpad = (cols + 127) & ~127, kblocks = kpad / 32;
  for (int r = 0; r < rows; r++)
    for (int kb = 0; kb < kblocks; kb++) {
      float x[32], mx = 0.f;
      for (int e = 0; e < 32; e++) {
        int c = kb * 32 + e;
        x[e] = c < cols ? src[(size_t)r * ld + c] : 0.f;
        if (fabsf(x[e]) > mx) mx = fabsf(x[e]);
      }
      uint8_t se = oracle_ue8m0_from_max(mx);
      int xe = 254 - (int)se;                               /* 2^-(se-127), clamped to normal floats like the device code */
      xe = xe < 1 ? 1 : (xe > 254 ? 254 : xe);
      uint32_t ib = (uint32_t)xe << 23;
      float inv;
      memcpy(&inv, &ib, 4);
      sf[(size_t)r * kblocks + kb] = se;
      for (int e = 0; e < 32; e += 2)
        q[(size_t)r * (kpad / 2) + kb * 16 + e / 2] =
            (uint8_t)(oracle_e2m1_encode(x[e] * inv) | (oracle_e2m1_encode(x[e + 1] * inv) << 4));
    }
}
/* plain scales (rows x kblocks) -> the 512-byte atom layout the tensor core's scale copy consumes
 * ([rows_pad/128][kpad/128][512]; byte (r%32)*16 + ((r/32)%4)*4 + kb%4); padding rows get scale 0. */
void oracle_mxf4_sf_to_atoms(int rows, int kpad, const uint8_t* sf, uint8_t* atoms) {
  int kblocks = kpad / 32, katoms = kpad / 128, rows_pad = (rows + 127) & ~127;
  memset(atoms, 0, (size_t)(rows_pad / 128) * katoms * 512);
  for (int r = 0; r < rows; r++)
    for (int kb = 0; kb < kblocks; kb++)
      atoms[((size_t)(r >> 7) * katoms + (kb >> 2)) * 512 + (size_t)(r & 31) * 16 + ((r >> 5) & 3) * 4 + (kb & 3)] =
          sf[(size_t)r * kblocks + kb];
}
/* C[m x n] = sum_k dq(A)[m,k] * dq(B^T)[n,k] in double: qa (m x kpad/2), qb (n x kpad/2), plain scales. */
typedef struct { int n, kpad; const uint8_t *qa, *sa, *qb, *sb; double* c; } mx_ctx;
static void rows_mxf4(int i0, int i1, void* vc) {
  mx_ctx* x = (mx_ctx*)vc;
  int kblocks = x->kpad / 32;
  for (int i = i0; i < i1; i++)
    for (int j = 0; j < x->n; j++) {
      double acc = 0.0;
      for (int kb = 0; kb < kblocks; kb++) {
        double s = oracle_ue8m0_value(x->sa[(size_t)i * kblocks + kb]) * oracle_ue8m0_value(x->sb[(size_t)j * kblocks + kb]);
        double part = 0.0;
        for (int e = 0; e < 16; e++) {
          uint8_t ba = x->qa[(size_t)i * (x->kpad / 2) + kb * 16 + e], bb = x->qb[(size_t)j * (x->kpad / 2) + kb * 16 + e];
          part += (double)oracle_e2m1_decode(ba & 15) * oracle_e2m1_decode(bb & 15) +
                  (double)oracle_e2m1_decode(ba >> 4) * oracle_e2m1_decode(bb >> 4);
        }
        acc += part * s;
      }
      x->c[(size_t)i * x->n + j] = acc;
    }
}
void oracle_mxf4_gemm(int m, int n, int kpad, const uint8_t* qa, const uint8_t* sa, const uint8_t* qb, const uint8_t* sb,
                      double* c) {
  mx_ctx x = {n, kpad, qa, sa, qb, sb, c};
  par_rows(m, rows_mxf4, &x);
}

