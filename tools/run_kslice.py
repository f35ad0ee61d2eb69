import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import _libs
g = _libs.load_pkg()
N, K = 4096, int(sys.argv[1])
A = torch.rand(N, N, device="cuda") - 0.5; B = torch.rand(N, N, device="cuda") - 0.5; C = torch.empty(N, N, device="cuda")
pa = g.PackedA(A, mode=5); pb = g.PackedB(B[:K], mode=5)
for _ in range(3):
    g.gemm_f32_packed_ab(pa, pb, C, a_k0=0)
torch.cuda.synchronize()
print("kslice", K, g.last_kernel())
