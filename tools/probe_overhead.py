"""Fixed vs per-K cost of the bf16 pair kernel (and cuBLAS via torch.matmul beside it):
t(K) = a + b*K at fixed M = N.   python tools/probe_overhead.py [N ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import _libs

g = _libs.load_pkg()
dev = "cuda"


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


S8 = "--s8" in sys.argv
for n in [int(x) for x in sys.argv[1:] if x.isdigit()] or [4096]:
    rows = []
    if S8:
        for k in (512, 1024, 2048, 4096, 8192, 16384):
            A = torch.randint(-127, 128, (n, k), device=dev, dtype=torch.int8)
            B = torch.randint(-127, 128, (k, n), device=dev, dtype=torch.int8)
            C = torch.empty(n, n, device=dev, dtype=torch.int32)
            us = t_us(lambda: g.gemm_s8s32(A, B, out=C))
            cub = t_us(lambda: torch._int_mm(A, B, out=C))
            rows.append((k, us, cub))
            print(f"N={n} K={k:6d}  ours s8 {us:8.1f} us {2.0*n*n*k/us/1e6:7.0f} TOPS | torch._int_mm {cub:8.1f} us {2.0*n*n*k/cub/1e6:7.0f} TOPS  {g.last_kernel()}", flush=True)
        (k0, a0, c0), (k1, a1, c1) = rows[1], rows[-1]
        for name, x0, x1 in (("ours s8", a0, a1), ("int_mm", c0, c1)):
            slope = (x1 - x0) / (k1 - k0)
            print(f"  {name}: per-1024-K {slope*1024:.2f} us, fixed {x0 - slope*k0:.2f} us, asymptotic {2.0*n*n/slope/1e6:.0f} TOPS")
        continue
    for k in (256, 512, 1024, 2048, 4096, 8192, 16384):
        A = (torch.rand(n, k, device=dev) - 0.5).bfloat16()
        B = (torch.rand(k, n, device=dev) - 0.5).bfloat16()
        C = torch.empty(n, n, device=dev, dtype=torch.float32)
        Cb = torch.empty(n, n, device=dev, dtype=torch.bfloat16)
        ours = t_us(lambda: g.gemm_bf16(A, B, out=C))
        ours_b = t_us(lambda: g.gemm_bf16(A, B, out=Cb, out_dtype=torch.bfloat16))
        cub = t_us(lambda: torch.matmul(A, B, out=Cb))
        fl = 2.0 * n * n * k
        rows.append((k, ours, ours_b, cub))
        print(f"N={n} K={k:6d}  ours f32-out {ours:8.1f} us {fl/ours/1e6:7.0f} TF  | bf16-out {ours_b:8.1f} us {fl/ours_b/1e6:7.0f} TF"
              f"  | cuBLAS bf16-out {cub:8.1f} us {fl/cub/1e6:7.0f} TF   {g.last_kernel()}", flush=True)
    (k0, a0, b0, c0), (k1, a1, b1, c1) = rows[2], rows[-1]
    for name, x0, x1 in (("ours f32", a0, a1), ("ours bf16", b0, b1), ("cublas", c0, c1)):
        slope = (x1 - x0) / (k1 - k0)
        print(f"  {name}: per-1024-K {slope*1024:.2f} us, fixed {x0 - slope*k0:.2f} us, asymptotic {2.0*n*n/slope/1e6:.0f} TF")
