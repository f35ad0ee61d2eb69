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
    # round 2: 4-bit path, packed operands, general epilogue, the row-panel plan (single rank, K-sliced)
    qa, sfa, _, _ = g.mxf4_quantize(A)
    qb, sfb, _, _ = g.mxf4_quantize(B, transpose=True)
    g.gemm_mxf4(qa, sfa, qb, sfb, m, n, k)
    print(m, n, k, g.last_kernel())
    pa, pb = g.PackedA(A, 5), g.PackedB(B, 5)
    Cc = torch.zeros(m, n, device=dev)
    g.gemm_f32_packed_ab(pa, pb, Cc)
    g.gemm_f32_ex(0.5, A, B, 2.0, Cc, mode=5)
    pa.close(); pb.close()
import importlib
sys.path.insert(0, _libs.ROOT)
rp = importlib.import_module(_libs.PKG + ".rowpanel")
A = torch.rand(300, 512, device=dev); B = torch.rand(512, 392, device=dev); Cc = torch.empty(300, 392, device=dev)
plan = rp.RowPanelPlan(g, 0, 300, 392, 512, 5, [(0, 128), (128, 512)])
plan.run(A, B, Cc)
plan.close()
g.lib.b200_gemm_debug_set_dynamic_sched(1)
g.gemm_f32(A, B, mode=5); g.gemm_bf16(A.bfloat16(), B.bfloat16())
g.lib.b200_gemm_debug_set_dynamic_sched(0)
g.gemm_f32(torch.rand(77, 77, device=dev), torch.rand(77, 77, device=dev), mode=0)
torch.cuda.synchronize()
print("sanitize_run ok")
