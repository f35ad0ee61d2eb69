"""int8 GEMM with int32 output vs the fused requant epilogue (int8 output): time at N^3."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import _libs

g = _libs.load_pkg()


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


for n in [int(x) for x in sys.argv[1:]] or [4096, 8192]:
    A = torch.randint(-127, 128, (n, n), device="cuda", dtype=torch.int8)
    B = torch.randint(-127, 128, (n, n), device="cuda", dtype=torch.int8)
    C32 = torch.empty(n, n, device="cuda", dtype=torch.int32)
    C8 = torch.empty(n, n, device="cuda", dtype=torch.int8)
    S = torch.rand(n, device="cuda") * 1e-4
    Bi = torch.rand(n, device="cuda")
    a = t_us(lambda: g.gemm_s8s32(A, B, out=C32)); ka = g.last_kernel()
    b = t_us(lambda: g.gemm_s8s8_requant(A, B, S, Bi, out=C8)); kb = g.last_kernel()
    print(f"N={n}: s8->s32 {a:8.1f} us {2.0*n**3/a/1e6:6.0f} TOPS {ka} | s8->s8 requant {b:8.1f} us {2.0*n**3/b/1e6:6.0f} TOPS {kb}")
