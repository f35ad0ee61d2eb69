"""Per-call split of B vs a pre-split B handle (b200_gemm_f32_pack_b), BF16X3, N^3."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import _libs

g = _libs.load_pkg()
for n in [int(x) for x in sys.argv[1:]] or [4096]:
    A = torch.rand(n, n, device="cuda") - 0.5
    B = torch.rand(n, n, device="cuda") - 0.5
    C = torch.empty(n, n, device="cuda")
    pk = g.PackedB(B, g.F32_BF16X3)
    res = {}
    for name, fn in (("per-call split", lambda: g.gemm_f32(A, B, out=C, mode=g.F32_BF16X3)),
                     ("pre-split B", lambda: g.gemm_f32_packed(A, pk, out=C))):
        for _ in range(5):
            fn()
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(20):
                fn()
            e.record()
            torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e) / 20)
        res[name] = best
        print(f"N={n} {name:15s} {best*1e3:8.1f} us {2*n**3/best/1e9:7.1f} TFLOP/s", flush=True)
