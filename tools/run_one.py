"""Launch one kernel family a few times (for ncu captures):  python tools/run_one.py KIND N [REPS]
KIND in {strict, tf32, bf16, bf16_obf16, s8, s8_requant, bf16x3, bf16x2, f16x2, mxf4, generic}."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import _libs

g = _libs.load_pkg()
if os.environ.get("B200_GROUP_ROWS"):            # tuning: rows of A per raster group of the tensor-core kernels
    g.lib.b200_gemm_debug_set_group_rows(int(os.environ["B200_GROUP_ROWS"]))
kind, n = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = "cuda"
if kind == "s8":
    A = torch.randint(-127, 128, (n, n), device=dev, dtype=torch.int8)
    B = torch.randint(-127, 128, (n, n), device=dev, dtype=torch.int8)
    fn = lambda: g.gemm_s8s32(A, B)
elif kind == "mxf4":
    A = torch.rand(n, n, device=dev) - 0.5
    B = torch.rand(n, n, device=dev) - 0.5
    def fn():
        qa, sfa, _, _ = g.mxf4_quantize(A)
        qb, sfb, _, _ = g.mxf4_quantize(B, transpose=True)
        g.gemm_mxf4(qa, sfa, qb, sfb, n, n, n)
elif kind == "generic":          # unaligned pitch: the CUDA-core fallback (strict mode, so no split pre-pass takes it)
    A = (torch.rand(n, n + 1, device=dev) - 0.5)[:, :n]
    B = (torch.rand(n, n + 3, device=dev) - 0.5)[:, :n]
    fn = lambda: g.gemm_f32(A, B, mode=0)
elif kind == "s8_requant":
    A = torch.randint(-127, 128, (n, n), device=dev, dtype=torch.int8)
    B = torch.randint(-127, 128, (n, n), device=dev, dtype=torch.int8)
    sc = torch.rand(n, device=dev) * 1e-4
    bi = torch.rand(n, device=dev)
    fn = lambda: g.gemm_s8s8_requant(A, B, sc, bi)
elif kind.startswith("bf16") and kind not in ("bf16x3", "bf16x2"):
    A = (torch.rand(n, n, device=dev) - 0.5).bfloat16()
    B = (torch.rand(n, n, device=dev) - 0.5).bfloat16()
    od = torch.bfloat16 if kind.endswith("obf16") else torch.float32
    fn = lambda: g.gemm_bf16(A, B, out_dtype=od)
else:
    A = torch.rand(n, n, device=dev) - 0.5
    B = torch.rand(n, n, device=dev) - 0.5
    mode = {"strict": 0, "tf32": 1, "bf16x3": 2, "bf16x2": 3, "f16x2": 5}[kind]
    fn = lambda: g.gemm_f32(A, B, mode=mode)
for _ in range(reps):
    fn()
torch.cuda.synchronize()
print(kind, n, g.last_kernel())
