// Where does a kernel's static shared memory start in the shared window on sm_100?
// Root-cause probe for the reference's MMult_cuda_11/12 failing their own harness on B200: those kernels
// declare `__shared__ __align__(16 * 1024) char smem[24 * 1024]` and ping-pong their buffers with
// `addr ^= 0x2000` / `addr ^= 0x1000` on the 32-bit shared-window address (cuda/MMult_cuda_12.cu:91-96,
// 138-139), which only works when the window address of smem[0] is a multiple of 16 KB.  sm_100 reserves the
// first 1 KB of the window for the system, so the array starts at 0x400 (also visible statically:
// `MOV R24, 0x400` / `LDS.128 [R3+0x400]` in the sm_100a SASS of the reference kernel, against base 0 for
// sm_86).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o smem_base_probe.x smem_base_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void probe(unsigned* out) {
  __shared__ __align__(16 * 1024) char smem[24 * 1024];
  smem[threadIdx.x] = (char)threadIdx.x;
  __syncthreads();
  if (threadIdx.x == 0) {
    out[0] = (unsigned)__cvta_generic_to_shared(smem);
    out[1] = (unsigned)smem[1];
  }
}

int main() {
  unsigned *d, h[2] = {0, 0};
  cudaMalloc(&d, 8);
  probe<<<1, 32>>>(d);
  cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s sm_%d%d: shared-window address of a 16KB-aligned static array = 0x%x (%s a multiple of 16 KB); "
         "0x%x ^ 0x2000 = 0x%x, (0x%x + 0x1c00) ^ 0x2000 = 0x%x\n",
         p.name, p.major, p.minor, h[0], (h[0] & 0x3fff) ? "NOT" : "is", h[0], h[0] ^ 0x2000, h[0], (h[0] + 0x1c00) ^ 0x2000);
  return 0;
}
