"""A/B of the epilogue store path (staged vs direct register stores): equality of results, then times."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import _libs

g = _libs.load_pkg()
sete = g.lib.b200_gemm_debug_set_epilogue


def t_us(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    return best


def cases(m, n, k):
    A = torch.rand(m, k, device="cuda") - 0.5
    B = torch.rand(k, n, device="cuda") - 0.5
    Ab, Bb = A.bfloat16(), B.bfloat16()
    A8 = torch.randint(-127, 128, (m, k), device="cuda", dtype=torch.int8)
    B8 = torch.randint(-127, 128, (k, n), device="cuda", dtype=torch.int8)
    return [("bf16_f32out", lambda: g.gemm_bf16(Ab, Bb)),
            ("bf16_bf16out", lambda: g.gemm_bf16(Ab, Bb, out_dtype=torch.bfloat16)),
            ("tf32", lambda: g.gemm_f32(A, B, mode=g.F32_TF32)),
            ("bf16x3", lambda: g.gemm_f32(A, B, mode=g.F32_BF16X3)),
            ("s8", lambda: g.gemm_s8s32(A8, B8))]


for shape in [(300, 528, 208), (1000, 1104, 2048), (2304, 2304, 512), (129, 4100, 64)]:
    for name, fn in cases(*shape):
        sete(0); c0 = fn(); sete(1); c1 = fn()
        print("equal", shape, name, g.last_kernel(), bool(torch.equal(c0, c1)), flush=True)
for n in (4096, 8192):
    for name, fn in cases(n, n, n):
        r = []
        for d in (0, 1, 0, 1):
            sete(d)
            r.append(t_us(fn))
        print(f"N={n} {name:13s} staged {min(r[0], r[2]):8.1f} us  direct {min(r[1], r[3]):8.1f} us  {g.last_kernel()}", flush=True)
sete(0)
