# 1-GPU batch: the whole -m gpu suite (incl. the 4-bit path and the reference harness binaries), probes
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
timeout 400 python tools/probe_r2.py 4096 2>&1 | tail -24
timeout 200 python - <<'PY'
import sys, os, torch
sys.path.insert(0, "tests"); import _libs
g = _libs.load_pkg()
N = 4096
A = torch.rand(N, N, device="cuda"); B = torch.rand(N, N, device="cuda"); C = torch.empty(N, N, device="cuda")
Ab, Bb = A.bfloat16(), B.bfloat16(); Cb = torch.empty(N, N, device="cuda", dtype=torch.bfloat16)
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); [fn() for _ in range(it)]; e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / it
for dyn in (1, 0, 1):
    g.lib.b200_gemm_debug_set_dynamic_sched(dyn)
    print("dynamic_sched", dyn, "f16x2 ms", round(t(lambda: g.gemm_f32(A, B, out=C, mode=5)), 4), "bf16 ms", round(t(lambda: g.gemm_bf16(Ab, Bb, out=Cb)), 4),
          "tf32 ms", round(t(lambda: g.gemm_f32(A, B, out=C, mode=1)), 4), flush=True)
PY
