"""Tile-shape A/B through the tuning hooks: bf16 (bf16 out) at a few sizes, every (cta_group, BN) the library has.
python tools/probe_tiles.py [N ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import _libs

g = _libs.load_pkg()


def t_us(fn, reps=20, warm=3):
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


for n in [int(x) for x in sys.argv[1:]] or [4096, 2304]:
    A = (torch.rand(n, n, device="cuda") - 0.5).bfloat16()
    B = (torch.rand(n, n, device="cuda") - 0.5).bfloat16()
    C = torch.empty(n, n, device="cuda", dtype=torch.bfloat16)
    Af, Bf, Cf = A.float(), B.float(), torch.empty(n, n, device="cuda")
    for cg, bn in ((2, 256), (1, 256), (1, 192), (1, 128)):
        for tail in (1, 0):
            g.lib.b200_gemm_debug_set_cta_group(cg)
            g.lib.b200_gemm_debug_set_bn(bn)
            g.lib.b200_gemm_debug_set_split_tail(tail)
            us = t_us(lambda: g.gemm_bf16(A, B, out=C))
            k1 = g.last_kernel()
            us3 = t_us(lambda: g.gemm_f32(Af, Bf, out=Cf, mode=g.F32_BF16X3))
            print(f"N={n} cg={cg} bn={bn} tail={tail}: bf16 {us:8.1f} us {2.0*n**3/us/1e6:7.0f} TF {k1:28s} | bf16x3 {us3:8.1f} us "
                  f"{2.0*n**3/us3/1e6:6.1f} TF {g.last_kernel()}", flush=True)
    us = t_us(lambda: torch.matmul(A, B, out=C))
    print(f"N={n} cuBLAS bf16 {us:8.1f} us {2.0*n**3/us/1e6:7.0f} TF")
g.lib.b200_gemm_debug_set_cta_group(0)
g.lib.b200_gemm_debug_set_bn(0)
g.lib.b200_gemm_debug_set_split_tail(1)
