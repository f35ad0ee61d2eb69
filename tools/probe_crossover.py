"""AUTO's size routing: strict (1 launch) vs BF16X3 (2 launches) vs F16X2 (4 launches) per square size."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _libs
g = _libs.load_pkg()
def timeit(fn, iters=20):
    for i in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
out = []
for n in (256, 384, 512, 640, 768, 896, 1024, 1152, 1280, 1536, 1792, 2048, 2304, 2560, 3072):
    a = torch.rand(n, n, device="cuda") - 0.5; b = torch.rand(n, n, device="cuda") - 0.5; c = torch.empty(n, n, device="cuda")
    r = {"n": n}
    for md, name in ((0, "strict"), (2, "bf16x3"), (5, "f16x2")):
        ms = timeit(lambda: g.gemm_f32(a, b, out=c, mode=md))
        r[name] = round(2.0 * n ** 3 / ms / 1e9, 1)
    r["best"] = max(("strict", "bf16x3", "f16x2"), key=lambda k: r[k])
    out.append(r); print(r, flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_crossover.json"), "w"), indent=1)
