cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 300 python - <<'PY'
import sys, os, torch
sys.path.insert(0, "tests"); import _libs
g = _libs.load_pkg()
N = 4096; R = 3
S = [(torch.rand(N, N, device="cuda"), torch.rand(N, N, device="cuda"), torch.empty(N, N, device="cuda")) for _ in range(R)]
Sb = [(a.bfloat16(), b.bfloat16(), torch.empty(N, N, device="cuda", dtype=torch.bfloat16)) for a, b, _ in S]
Si = [(torch.randint(-127, 128, (N, N), device="cuda", dtype=torch.int8), torch.randint(-127, 128, (N, N), device="cuda", dtype=torch.int8), torch.empty(N, N, device="cuda", dtype=torch.int32)) for _ in range(R)]
def t(fn, it=20):
    for i in range(3): fn(i)
    torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); [fn(i) for i in range(it)]; e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / it
for dyn, pdl in ((1, 1), (0, 1), (1, 0), (0, 0), (1, 1), (0, 1)):
    g.lib.b200_gemm_debug_set_dynamic_sched(dyn); g.lib.b200_gemm_debug_set_pdl(pdl)
    print("dynamic", dyn, "pdl", pdl, "f16x2", round(t(lambda i: g.gemm_f32(S[i % R][0], S[i % R][1], out=S[i % R][2], mode=5)), 4),
          "bf16", round(t(lambda i: g.gemm_bf16(Sb[i % R][0], Sb[i % R][1], out=Sb[i % R][2])), 4),
          "s8", round(t(lambda i: g.gemm_s8s32(Si[i % R][0], Si[i % R][1], out=Si[i % R][2])), 4),
          "tf32", round(t(lambda i: g.gemm_f32(S[i % R][0], S[i % R][1], out=S[i % R][2], mode=1)), 4), flush=True)
PY
