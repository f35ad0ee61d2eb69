// MY_MMult_b200.cpp — the link-time plug-in the reference's harnesses expect.
//
// The reference selects an implementation by linking exactly one object that defines MY_MMult
// (cuda/makefile:25, aarch64/makefile:24).  This translation unit is that object for libb200gemm:
//
//   void MY_MMult(cublasHandle_t, int m, int n, int k, float* dA, int lda, float* dB, int ldb,
//                 float* dC, int ldc)              cuda/test_MMult.cpp:13-14   device ptrs, C  = A*B
//   void MY_MMult(int m, int n, int k, float* a, int lda, float* b, int ldb, float* c, int ldc)
//                                                  aarch64/MMult0.cpp:3-4      host ptrs,   C += A*B
//
// Both have C++ linkage in the reference (the files are .cpp with no extern "C"), so the mangled
// names produced here are the ones test_MMult.o references.  Error policy follows
// cuda/helper.h:7-17: print and exit(EXIT_FAILURE).
#include <cstdio>
#include <cstdlib>

#include <cuda_runtime.h>

#include "../../include/b200gemm.h"

struct cublasContext;                       // same opaque type cublas_v2.h declares
typedef struct cublasContext* cublasHandle_t;

static void die_on(int rc, const char* what) {
  if (rc != 0) {
    std::fprintf(stderr, "b200gemm error at %s: code=%d \"%s\"\n", what, rc, b200_gemm_strerror(rc));
    std::exit(EXIT_FAILURE);
  }
}

// cuda/ harness: 20 back-to-back asynchronous launches on the default stream
// (cuda/test_MMult.cpp:98-110).  The handle is ignored, as in cuda/MMult_cuda_12.cu:228.
void MY_MMult(cublasHandle_t, int m, int n, int k, float* d_A, int lda, float* d_B, int ldb,
              float* d_C, int ldc) {
  die_on(b200_gemm_f32(m, n, k, d_A, lda, d_B, ldb, d_C, ldc, B200_F32_AUTO, nullptr), "MY_MMult(cuda)");
}

// aarch64/ harness: host pointers, C pre-zeroed by the caller, C += A*B
// (aarch64/test_MMult.cpp:107-110).  Device pointers are accepted too, with the SAME contract: the 9-argument
// form always means C += A*B (aarch64/MMult0.cpp:16), wherever the operands live.
void MY_MMult(int m, int n, int k, float* a, int lda, float* b, int ldb, float* c, int ldc) {
  cudaPointerAttributes at;
  const bool on_device =
      cudaPointerGetAttributes(&at, c) == cudaSuccess && at.type == cudaMemoryTypeDevice;
  cudaGetLastError();
  if (on_device)
    die_on(b200_gemm_f32_acc(m, n, k, a, lda, b, ldb, c, ldc, B200_F32_AUTO, nullptr), "MY_MMult(device)");
  else
    die_on(b200_gemm_f32_host(m, n, k, a, lda, b, ldb, c, ldc, B200_F32_AUTO), "MY_MMult(host)");
}
