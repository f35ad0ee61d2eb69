cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_rowpanel_gpu.py -x -q -m gpu -k "f16 or packed or split or default or alpha or rowpanel or full_size" 2>&1 | tail -3
timeout 300 python - <<'PY'
import sys, os, torch
sys.path.insert(0, "tests"); import _libs
g = _libs.load_pkg()
R = 3
def t(fn, it=20):
    for i in range(3): fn(i)
    torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); [fn(i) for i in range(it)]; e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / it
for N in (4096, 2048, 8192):
    S = [(torch.rand(N, N, device="cuda"), torch.rand(N, N, device="cuda"), torch.empty(N, N, device="cuda")) for _ in range(R)]
    for mask in (1, 3, 1, 3):
        g.lib.b200_gemm_debug_set_pdl(mask)
        print("N", N, "prepass fork", "off" if mask & 2 else "on", "f16x2 step ms", round(t(lambda i: g.gemm_f32(S[i % R][0], S[i % R][1], out=S[i % R][2], mode=5)), 4), flush=True)
    del S
PY
