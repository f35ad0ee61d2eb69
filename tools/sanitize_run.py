"""Small-shape pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import _libs
g = _libs.load_pkg()
dev = "cuda"
for (m, n, k) in [(256, 256, 128), (300, 392, 96), (128, 256, 64)]:
    A = torch.rand(m, k, device=dev) - 0.5
    B = torch.rand(k, n, device=dev) - 0.5
    for mode in (0, 1, 2, 3, 5):
        g.gemm_f32(A, B, mode=mode)
        print(m, n, k, g.last_kernel())
    g.gemm_bf16(A.bfloat16(), B.bfloat16())
    g.gemm_bf16(A.bfloat16(), B.bfloat16(), out_dtype=torch.bfloat16)
    kk = (k + 15) // 16 * 16; nn = (n + 15) // 16 * 16
    g.gemm_s8s32(torch.randint(-127, 128, (m, kk), device=dev, dtype=torch.int8), torch.randint(-127, 128, (kk, nn), device=dev, dtype=torch.int8))
    print(m, n, k, g.last_kernel())
g.gemm_f32(torch.rand(77, 77, device=dev), torch.rand(77, 77, device=dev), mode=0)
torch.cuda.synchronize()
print("sanitize_run ok")
