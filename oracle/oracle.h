/*
 * oracle.h — CPU restatement of the reference's GEMM hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (libb200gemm.so, the
 * shims, the harness) may include, link or call this.  Allowed users: tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * Parity status: PINNED.  oracle/Makefile builds oracle/_ref/libref.so from the
 * reference's own sources where they lie under /root/reference (its REF_MMult,
 * random_matrix, compare_matrices for the cuda/, aarch64/ and aarch64-int8/
 * harnesses, plus the vendored OpenBLAS-0.2.20 that cuda/REF_MMult.cpp:11
 * calls); tests/test_oracle_vs_ref.py checks every function below against it
 * bit-for-bit (integer, generators, naive fp32) or to 1 ulp-scale tolerance
 * (OpenBLAS, different summation order), and tests/golden/ holds vectors
 * generated from libref.so by tests/golden/make_golden.py.
 */
#ifndef ORACLE_H_
#define ORACLE_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Threads used by the *_fast variants (0 = all online cores). */
void oracle_set_threads(int n);
int  oracle_get_threads(void);

/* ---- input generators ------------------------------------------------------ */
/* srand48 wrapper: the reference never seeds; tests do. */
void oracle_seed(long seed);
/* cuda/random_matrix.cpp:6-16: A(i,j)=a[j*lda+i] = 2.0*(float)drand48()-1.0,
 * i outer, j inner.  Called by the harness as (m,k,a,m) / (k,n,b,k)
 * (cuda/test_MMult.cpp:77-78). */
void oracle_random_matrix_cuda(int m, int n, float* a, int lda);
/* aarch64/random_matrix.cpp:3-20: every element 1.0f (row-major m x n). */
void oracle_random_matrix_ones(int m, int n, float* a);
/* aarch64-int8/random_matrix.c:9-25: a[i*lda+j] = (int8_t)(val++ % 3). */
void oracle_random_int8_ramp(int m, int n, int8_t* a, int lda);
/* NOT in the reference: seeded uniform integers in [-127,127] (never -128, the
 * chgemm input contract, /root/reference/README.md:82), 64-bit LCG. */
void oracle_random_int8_uniform(int m, int n, int8_t* a, int lda, uint64_t seed);

/* ---- the contraction -------------------------------------------------------- */
/* aarch64/REF_MMult.cpp:18-28 (and armv7/REF_MMult.c:9-22): naive i,j,p,
 * C(i,j) += A(i,p)*B(p,j) in fp32, p ascending, separate multiply and add
 * (no FMA contraction — this file is built with -ffp-contract=off). */
void oracle_ref_mmult_f32(int m, int n, int k, const float* a, int lda,
                          const float* b, int ldb, float* c, int ldc);
/* The same loop nest with the multiply-add FUSED (fmaf): this is what the
 * reference's own build flags produce — aarch64/makefile:14 compiles with
 * `-O2 -march=native`, GCC's default -ffp-contract=fast turns the statement at
 * aarch64/REF_MMult.cpp:24 into one fmadd/vfmadd.  A sequential-k FFMA GPU
 * kernel is bit-identical to THIS variant. */
void oracle_ref_mmult_f32_fma(int m, int n, int k, const float* a, int lda,
                              const float* b, int ldb, float* c, int ldc);
/* Same arithmetic per element (identical bit results), loop order i,p,j and
 * pthreads over rows: the form fast enough for N = 1024..4096 fixtures. */
void oracle_ref_mmult_f32_fast(int m, int n, int k, const float* a, int lda,
                               const float* b, int ldb, float* c, int ldc);
void oracle_ref_mmult_f32_fma_fast(int m, int n, int k, const float* a, int lda,
                                   const float* b, int ldb, float* c, int ldc);
/* fp64-accumulated truth for error analysis (not a reference function). */
void oracle_ref_mmult_f64acc(int m, int n, int k, const float* a, int lda,
                             const float* b, int ldb, double* c, int ldc);
/* aarch64-int8/REF_MMult.c:10-23: C(i,j) += A(i,p)*B(p,j), int32 accumulate. */
void oracle_ref_mmult_s8s32(int m, int n, int k, const int8_t* a, int lda,
                            const int8_t* b, int ldb, int32_t* c, int ldc);
void oracle_ref_mmult_s8s32_fast(int m, int n, int k, const int8_t* a, int lda,
                                 const int8_t* b, int ldb, int32_t* c, int ldc);

/* Requant tail of chgemm's kernels, aarch64-int8/int8kernel_m4.S:386-426 (same in _m2 :647-684, _m1 :851-):
 * out(i,j) = sqxtn(sqxtn(fcvtas(fadd(fmul(scvtf(c(i,j)), scales[i]), bias[i])))) ; bias may be NULL (:399-400).
 * PARITY UNPINNED for this function: the reference ships no C caller, test or vector for the requant
 * kernels and the .S files are AArch64-only, so this restates the instruction sequence and nothing here
 * can be checked against the reference running. */
void oracle_requant_s32_to_s8(int m, int n, const int32_t* c, int ldc, const float* scales,
                              const float* bias, int8_t* out, int ldo);

/* ---- checkers --------------------------------------------------------------- */
/* cuda/compare_matrices.cpp:7-30: max_ij |A(i,j)-B(i,j)| (row-major).  Unlike
 * the reference's macro abs (Appendix B-7) a NaN anywhere returns NaN. */
float oracle_compare_matrices_f32(int m, int n, const float* a, int lda,
                                  const float* b, int ldb);
/* aarch64-int8/compare_matrices.c:8-33: integer max |a-b|. */
int32_t oracle_compare_matrices_s32(int m, int n, const int32_t* a, int lda,
                                    const int32_t* b, int ldb);
/* max_ij |c(i,j)| — denominator of north_star's "max relative error". */
float oracle_max_abs_f32(int m, int n, const float* a, int lda);
/* max |float(c) - truth| against the fp64 truth. */
double oracle_max_err_vs_f64(int m, int n, const float* c, int ldc,
                             const double* t, int ldt);

/* ---- dtype helpers ---------------------------------------------------------- */
/* Round-to-nearest-even fp32 -> bf16 bit pattern (config 3 input rounding,
 * SURVEY §8d); NaN stays NaN (quiet). */
uint16_t oracle_f32_to_bf16(float x);
float    oracle_bf16_to_f32(uint16_t h);
void     oracle_round_to_bf16_inplace(size_t count, float* a);

/* ---- MXFP4 (4-bit path): OCP MX v1.0 restated; parity UNPINNED (reference: cuda-int4/README.md:1 "WIP") ---- */
uint8_t oracle_e2m1_encode(float v);
float   oracle_e2m1_decode(uint8_t code);
uint8_t oracle_ue8m0_from_max(float max_abs);
double  oracle_ue8m0_value(uint8_t e);
void    oracle_mxf4_quantize(int rows, int cols, const float* src, int ld, uint8_t* q, uint8_t* sf);
void    oracle_mxf4_sf_to_atoms(int rows, int kpad, const uint8_t* sf, uint8_t* atoms);
void    oracle_mxf4_gemm(int m, int n, int kpad, const uint8_t* qa, const uint8_t* sa, const uint8_t* qb,
                         const uint8_t* sb, double* c);

#ifdef __cplusplus
}
#endif
#endif
