"""GEMM-only time of the F16X2 kernel vs K (M = N = 4096), with and without C += (what the row-panel plan's K-slices cost)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _libs
g = _libs.load_pkg()
dev = "cuda"
N = 4096
R = 3
def timeit(fn, iters=20):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(iters): fn(i)
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
A = [torch.rand(N, N, device=dev) - 0.5 for _ in range(R)]
B = [torch.rand(N, N, device=dev) - 0.5 for _ in range(R)]
C = [torch.empty(N, N, device=dev) for _ in range(R)]
pa = [g.PackedA(a, mode=5) for a in A]
out = []
for K in (256, 512, 1024, 1536, 2048, 3072, 4096):
    pb = [g.PackedB(b[:K], mode=5) for b in B]
    for acc in (False, True):
        for dyn in (0, 1):
            g.lib.b200_gemm_debug_set_dynamic_sched(dyn)
            ms = timeit(lambda i: g.gemm_f32_packed_ab(pa[i % R], pb[i % R], C[i % R], a_k0=0, accumulate=acc))
            out.append({"K": K, "accumulate": acc, "dynamic": dyn, "us": round(ms * 1e3, 1), "tflops": round(2.0 * N * N * K / ms / 1e9, 1)})
            print(out[-1], flush=True)
    for p in pb: p.close()
g.lib.b200_gemm_debug_set_dynamic_sched(0)
# the same for the plain bf16 kernel (no split): how much of the fixed cost is the kernel family's
Ab = [a.bfloat16() for a in A]; Bb = [b.bfloat16() for b in B]; Cb = [torch.empty(N, N, device=dev, dtype=torch.bfloat16) for _ in range(R)]
for K in (512, 1024, 2048, 4096):
    ms = timeit(lambda i: g.gemm_bf16(Ab[i % R][:, :K], Bb[i % R][:K], out=Cb[i % R]))
    out.append({"bf16_K": K, "us": round(ms * 1e3, 1), "tflops": round(2.0 * N * N * K / ms / 1e9, 1)})
    print(out[-1], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_kslice.json"), "w"), indent=1)
