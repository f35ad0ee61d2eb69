"""GPU probe (run under gpurun, not pytest): pins the UMMA descriptor semantics and gives first
timings.  Each case runs in its own subprocess with a timeout so a hang or sticky CUDA error in one
kernel cannot take the rest down.  Writes gpurun_out/probe.jsonl.

    python tests/probe_gpu.py            # all cases
    python tests/probe_gpu.py case NAME  # one case (internal)
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _libs  # noqa: E402

OUT = os.path.join(_libs.ROOT, "gpurun_out")


def emit(**kw):
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "probe.jsonl"), "a") as f:
        f.write(json.dumps(kw) + "\n")
    print(json.dumps(kw), flush=True)


def time_call(fn, reps=10, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def case_strict():
    import torch
    o, g = _libs.load_oracle(), _libs.load_pkg()
    for (m, n, k, pad) in [(128, 128, 64, 0), (256, 384, 512, 0), (300, 260, 100, 0), (77, 77, 77, 0),
                           (130, 70, 257, 3), (1024, 1024, 1024, 0)]:
        a = _libs.gen_f32(o, m, k + pad, 11)[:, :k]
        b = _libs.gen_f32(o, k, n + pad, 12)[:, :n]
        ref = _libs.ref_f32_fma(o, np.ascontiguousarray(a), np.ascontiguousarray(b))
        A = torch.from_numpy(np.ascontiguousarray(_libs.gen_f32(o, m, k + pad, 11))).cuda()[:, :k]
        B = torch.from_numpy(np.ascontiguousarray(_libs.gen_f32(o, k, n + pad, 12))).cuda()[:, :n]
        Cg = g.gemm_f32(A, B, mode=g.F32_STRICT).cpu().numpy()
        emit(case="strict", shape=[m, n, k], pad=pad, kernel=g.last_kernel(),
             bit_exact=bool(np.array_equal(Cg, ref)), maxdiff=float(np.abs(Cg - ref).max()))
    # host entry: C += A*B with non-zero C
    m, n, k = 96, 80, 160
    a, b = _libs.gen_f32(o, m, k, 1), _libs.gen_f32(o, k, n, 2)
    c0 = _libs.gen_f32(o, m, n, 3)
    ref = _libs.ref_f32_fma(o, a, b, c0)
    c = c0.copy()
    g.MY_MMult(m, n, k, a, k, b, n, c, n, mode=g.F32_STRICT)
    emit(case="strict_host_accumulate", bit_exact=bool(np.array_equal(c, ref)))
    for N in (2048, 4096):
        A = torch.rand(N, N, device="cuda") - 0.5
        B = torch.rand(N, N, device="cuda") - 0.5
        Cc = torch.empty(N, N, device="cuda")
        ms = time_call(lambda: g.gemm_f32(A, B, out=Cc, mode=g.F32_STRICT))
        emit(case="strict_time", N=N, ms=ms, tflops=2 * N ** 3 / ms / 1e9)


def _tc_inputs(o, kind, m, n, k, pattern=None):
    import torch
    if kind == "s8":
        a, b = _libs.gen_s8(o, m, k, 5), _libs.gen_s8(o, k, n, 6)
        if pattern == "ident":
            a = np.zeros((m, k), np.int8)
            for i in range(min(m, k)):
                a[i, i] = 1
            b = ((np.arange(k)[:, None] * 7 + np.arange(n)[None, :]) % 100).astype(np.int8)
        ref = _libs.ref_s8(o, a, b)
        return torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), ref, a, b
    a, b = _libs.gen_f32(o, m, k, 5), _libs.gen_f32(o, k, n, 6)
    if pattern == "ident":
        a = np.zeros((m, k), np.float32)
        for i in range(min(m, k)):
            a[i, i] = 1
        b = ((np.arange(k)[:, None] * 7 + np.arange(n)[None, :]) % 100).astype(np.float32)
    if kind == "bf16":
        a, b = _libs.round_bf16(o, a), _libs.round_bf16(o, b)
        A, B = torch.from_numpy(a).cuda().bfloat16(), torch.from_numpy(b).cuda().bfloat16()
    else:
        A, B = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    return A, B, _libs.ref_f64(o, a, b), a, b


def _run_tc(g, kind, A, B):
    if kind == "bf16":
        return g.gemm_bf16(A, B).cpu().numpy()
    if kind == "tf32":
        return g.gemm_f32(A, B, mode=g.F32_TF32).cpu().numpy()
    return g.gemm_s8s32(A, B).cpu().numpy()


def case_tc(kind):
    import torch
    o, g = _libs.load_oracle(), _libs.load_pkg()
    box = {"bf16": 8192, "tf32": 4096, "s8": 16384}[kind]
    tol = {"bf16": 1e-3, "tf32": 5e-2, "s8": 0}[kind]
    cands = [(0, 0), (1024, box), (box, 128), (128, box)]
    good = None
    for (lbo, sbo) in cands:
        g.lib.b200_gemm_debug_set_b_desc(lbo, sbo)
        A, B, ref, a, b = _tc_inputs(o, kind, 128, 256, 256, "ident")
        Cg = _run_tc(g, kind, A, B)
        err = float(np.abs(Cg.astype(np.float64) - ref).max())
        emit(case="tc_desc", kind=kind, lbo=lbo, sbo=sbo, maxdiff=err, kernel=g.last_kernel())
        if err <= tol:
            good = (lbo, sbo)
            break
        emit(case="tc_desc_dump", kind=kind, lbo=lbo, sbo=sbo, got=Cg[:3, :20].tolist(), want=ref[:3, :20].tolist())
    if good is None:
        emit(case="tc_desc_FAILED", kind=kind)
        return
    for (m, n, k) in [(128, 256, 64), (128, 128, 512), (256, 512, 1024), (300, 520, 200), (77, 96, 80),
                      (1024, 1024, 1024), (129, 257, 4096)]:
        kk = k if kind != "s8" else (k + 15) // 16 * 16
        nn = n if kind != "s8" else (n + 15) // 16 * 16
        A, B, ref, a, b = _tc_inputs(o, kind, m, nn, kk)
        Cg = _run_tc(g, kind, A, B)
        d = np.abs(Cg.astype(np.float64) - ref)
        emit(case="tc_parity", kind=kind, shape=[m, nn, kk], kernel=g.last_kernel(), maxdiff=float(d.max()),
             maxrel=float(d.max() / max(np.abs(ref).max(), 1e-30)))
    for N in (4096, 8192):
        if kind == "s8":
            A = torch.randint(-127, 128, (N, N), device="cuda", dtype=torch.int8)
            B = torch.randint(-127, 128, (N, N), device="cuda", dtype=torch.int8)
            Cc = torch.empty(N, N, device="cuda", dtype=torch.int32)
            fn = lambda: g.gemm_s8s32(A, B, out=Cc)
        elif kind == "bf16":
            A = (torch.rand(N, N, device="cuda") - 0.5).bfloat16()
            B = (torch.rand(N, N, device="cuda") - 0.5).bfloat16()
            Cc = torch.empty(N, N, device="cuda", dtype=torch.bfloat16)
            fn = lambda: g.gemm_bf16(A, B, out=Cc)
        else:
            A = torch.rand(N, N, device="cuda") - 0.5
            B = torch.rand(N, N, device="cuda") - 0.5
            Cc = torch.empty(N, N, device="cuda")
            fn = lambda: g.gemm_f32(A, B, out=Cc, mode=g.F32_TF32)
        ms = time_call(fn)
        emit(case="tc_time", kind=kind, N=N, ms=ms, tflops=2 * N ** 3 / ms / 1e9)
    if kind == "bf16":
        N = 4096
        A = (torch.rand(N, N, device="cuda") - 0.5).bfloat16()
        B = (torch.rand(N, N, device="cuda") - 0.5).bfloat16()
        ms = time_call(lambda: torch.matmul(A, B))
        emit(case="cublas_bf16_time", N=N, ms=ms, tflops=2 * N ** 3 / ms / 1e9)
    if kind == "tf32":
        N = 4096
        A = torch.rand(N, N, device="cuda") - 0.5
        B = torch.rand(N, N, device="cuda") - 0.5
        ms = time_call(lambda: torch.matmul(A, B))
        emit(case="cublas_sgemm_time", N=N, ms=ms, tflops=2 * N ** 3 / ms / 1e9)


def case_trunc():
    """Does kind::tf32 truncate or round the 13 low mantissa bits of its fp32 operands?"""
    import torch
    g = _libs.load_pkg()
    m, n, k = 128, 256, 32
    a = np.zeros((m, k), np.float32)
    b = np.zeros((k, n), np.float32)
    a[:, 0] = np.float32(1.0 + 2.0 ** -11 + 2.0 ** -12)   # rounds up to 1+2^-10 under RN, truncates to 1.0
    b[0, :] = 1.0
    Cg = g.gemm_f32(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), mode=g.F32_TF32).cpu().numpy()
    emit(case="tf32_rounding", value=float(Cg[0, 0]), truncates=bool(Cg[0, 0] == 1.0))


def case_split():
    import torch
    o, g = _libs.load_oracle(), _libs.load_pkg()
    for mode, name in ((g.F32_BF16X3, "bf16x3"), (g.F32_BF16X2, "bf16x2"), (g.F32_STRICT, "strict"), (g.F32_TF32, "tf32")):
        for (m, n, k) in [(128, 192, 32), (128, 128, 64), (256, 512, 1024), (300, 520, 200), (77, 96, 80), (1000, 1100, 4096)]:
            a, b = _libs.gen_f32(o, m, k, 5), _libs.gen_f32(o, k, n, 6)
            t = _libs.ref_f64(o, a, b)
            Cg = g.gemm_f32(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), mode=mode).cpu().numpy()
            d = np.abs(Cg.astype(np.float64) - t)
            emit(case="split_parity", mode=name, shape=[m, n, k], kernel=g.last_kernel(), maxdiff=float(d.max()),
                 maxrel=float(d.max() / np.abs(t).max()))
    for N in (4096, 8192):
        A = torch.rand(N, N, device="cuda") - 0.5
        B = torch.rand(N, N, device="cuda") - 0.5
        Cc = torch.empty(N, N, device="cuda")
        for mode, name in ((g.F32_BF16X3, "bf16x3"), (g.F32_BF16X2, "bf16x2"), (g.F32_TF32, "tf32")):
            for bn in (0, 128, 192, 256):
                g.lib.b200_gemm_debug_set_bn(bn)
                ms = time_call(lambda: g.gemm_f32(A, B, out=Cc, mode=mode))
                emit(case="split_time", mode=name, N=N, bn=bn, kernel=g.last_kernel(), ms=ms, tflops=2 * N ** 3 / ms / 1e9)
        Ab, Bb = A.bfloat16(), B.bfloat16()
        Cb = torch.empty(N, N, device="cuda", dtype=torch.bfloat16)
        for bn in (0, 128, 192, 256):
            g.lib.b200_gemm_debug_set_bn(bn)
            ms = time_call(lambda: g.gemm_bf16(Ab, Bb, out=Cb))
            emit(case="bf16_time", N=N, bn=bn, kernel=g.last_kernel(), ms=ms, tflops=2 * N ** 3 / ms / 1e9)
        g.lib.b200_gemm_debug_set_bn(0)


def case_chunk():
    """Two-level accumulation: error and time of the split modes vs the K-chunk folded into C."""
    import torch
    o, g = _libs.load_oracle(), _libs.load_pkg()
    m, n, k = 1000, 1100, 4096
    a, b = _libs.gen_f32(o, m, k, 5), _libs.gen_f32(o, k, n, 6)
    t = _libs.ref_f64(o, a, b)
    A, B = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    N = 4096
    A4 = torch.rand(N, N, device="cuda") - 0.5
    B4 = torch.rand(N, N, device="cuda") - 0.5
    C4 = torch.empty(N, N, device="cuda")
    for mode, name in ((g.F32_BF16X3, "bf16x3"), (g.F32_BF16X2, "bf16x2")):
        for ck in (0, 2048, 1024, 512, 256, 128):
            g.lib.b200_gemm_debug_set_split_chunk(ck, ck)
            Cg = g.gemm_f32(A, B, mode=mode).cpu().numpy()
            d = np.abs(Cg.astype(np.float64) - t)
            ms = time_call(lambda: g.gemm_f32(A4, B4, out=C4, mode=mode))
            emit(case="chunk", mode=name, chunk_k=ck, kernel=g.last_kernel(), maxrel=float(d.max() / np.abs(t).max()),
                 ms_4096=ms, tflops_4096=2 * N ** 3 / ms / 1e9)


def case_pair():
    """CTA pairs (tcgen05 cta_group::2): parity of every kind against the oracle, then 1-CTA vs 2-CTA time."""
    import torch
    o, g = _libs.load_oracle(), _libs.load_pkg()
    g.lib.b200_gemm_debug_set_cta_group(2)
    for (m, n, k) in [(256, 256, 64), (256, 256, 256), (384, 512, 1024), (300, 520, 200), (1000, 1100, 2048)]:
        a, b = _libs.gen_f32(o, m, k, 5), _libs.gen_f32(o, k, n, 6)
        t = _libs.ref_f64(o, a, b)
        A, B = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        for mode, name in ((g.F32_TF32, "tf32"), (g.F32_BF16X3, "bf16x3"), (g.F32_BF16X2, "bf16x2")):
            Cg = g.gemm_f32(A, B, mode=mode).cpu().numpy()
            emit(case="pair_parity", kind=name, shape=[m, n, k], kernel=g.last_kernel(),
                 maxrel=float(np.abs(Cg - t).max() / np.abs(t).max()))
        ab, bb = _libs.round_bf16(o, a), _libs.round_bf16(o, b)
        tb = _libs.ref_f64(o, ab, bb)
        Cg = g.gemm_bf16(torch.from_numpy(ab).cuda().bfloat16(), torch.from_numpy(bb).cuda().bfloat16()).cpu().numpy()
        emit(case="pair_parity", kind="bf16", shape=[m, n, k], kernel=g.last_kernel(),
             maxrel=float(np.abs(Cg - tb).max() / np.abs(tb).max()))
        kk, nn = (k + 15) // 16 * 16, (n + 15) // 16 * 16
        a8, b8 = _libs.gen_s8(o, m, kk, 7), _libs.gen_s8(o, kk, nn, 8)
        Cg = g.gemm_s8s32(torch.from_numpy(a8).cuda(), torch.from_numpy(b8).cuda()).cpu().numpy()
        emit(case="pair_parity", kind="s8", shape=[m, nn, kk], kernel=g.last_kernel(),
             exact=bool(np.array_equal(Cg, _libs.ref_s8(o, a8, b8))))
    for N in (4096, 8192):
        A = torch.rand(N, N, device="cuda") - 0.5
        B = torch.rand(N, N, device="cuda") - 0.5
        Cc = torch.empty(N, N, device="cuda")
        Ab, Bb = A.bfloat16(), B.bfloat16()
        Cb = torch.empty(N, N, device="cuda", dtype=torch.bfloat16)
        A8 = torch.randint(-127, 128, (N, N), device="cuda", dtype=torch.int8)
        B8 = torch.randint(-127, 128, (N, N), device="cuda", dtype=torch.int8)
        C8 = torch.empty(N, N, device="cuda", dtype=torch.int32)
        for cg, tail in ((1, 0), (1, 1), (2, 0), (2, 1)):
            g.lib.b200_gemm_debug_set_cta_group(cg)
            g.lib.b200_gemm_debug_set_split_tail(tail)
            for name, fn in (("bf16", lambda: g.gemm_bf16(Ab, Bb, out=Cb)),
                             ("tf32", lambda: g.gemm_f32(A, B, out=Cc, mode=g.F32_TF32)),
                             ("bf16x3", lambda: g.gemm_f32(A, B, out=Cc, mode=g.F32_BF16X3)),
                             ("bf16x2", lambda: g.gemm_f32(A, B, out=Cc, mode=g.F32_BF16X2)),
                             ("s8", lambda: g.gemm_s8s32(A8, B8, out=C8))):
                ms = time_call(fn)
                emit(case="pair_time", kind=name, N=N, cg=cg, tail=tail, kernel=g.last_kernel(), ms=ms, tflops=2 * N ** 3 / ms / 1e9)
    g.lib.b200_gemm_debug_set_cta_group(0)
    g.lib.b200_gemm_debug_set_split_tail(1)


def case_ab():
    """Noise-robust A/B of the scheduling knobs: interleaved trials, min and median of per-trial means."""
    import statistics
    import torch
    g = _libs.load_pkg()
    for N in (4096, 8192):
        A = torch.rand(N, N, device="cuda") - 0.5
        B = torch.rand(N, N, device="cuda") - 0.5
        Cc = torch.empty(N, N, device="cuda")
        Ab, Bb = A.bfloat16(), B.bfloat16()
        Cf = torch.empty(N, N, device="cuda")
        Cb = torch.empty(N, N, device="cuda", dtype=torch.bfloat16)
        kinds = (("bf16_f32out", lambda: g.gemm_bf16(Ab, Bb, out=Cf)), ("bf16_bf16out", lambda: g.gemm_bf16(Ab, Bb, out=Cb)),
                 ("tf32", lambda: g.gemm_f32(A, B, out=Cc, mode=g.F32_TF32)),
                 ("bf16x3", lambda: g.gemm_f32(A, B, out=Cc, mode=g.F32_BF16X3)))
        cfgs = ((2, 0), (2, 1), (2, 2), (1, 0), (1, 1))
        res = {(k, c): [] for k, _ in kinds for c in cfgs}
        for trial in range(5):
            for name, fn in kinds:
                for cfg in cfgs:
                    g.lib.b200_gemm_debug_set_cta_group(cfg[0])
                    g.lib.b200_gemm_debug_set_split_tail(cfg[1])
                    res[(name, cfg)].append(time_call(fn, reps=20, warm=2))
        for (name, cfg), v in res.items():
            emit(case="ab", kind=name, N=N, cg=cfg[0], tail=cfg[1], ms_min=min(v), ms_med=statistics.median(v),
                 tflops_best=2 * N ** 3 / min(v) / 1e9, tflops_med=2 * N ** 3 / statistics.median(v) / 1e9)
        ms = [time_call(lambda: torch.matmul(Ab, Bb), reps=20, warm=2) for _ in range(5)]
        emit(case="ab_cublas_bf16", N=N, tflops_best=2 * N ** 3 / min(ms) / 1e9, tflops_med=2 * N ** 3 / statistics.median(ms) / 1e9)
    g.lib.b200_gemm_debug_set_cta_group(0)
    g.lib.b200_gemm_debug_set_split_tail(1)


CASES = {"ab": case_ab, "pair": case_pair,"chunk": case_chunk,"split": case_split,"strict": case_strict, "bf16": lambda: case_tc("bf16"), "tf32": lambda: case_tc("tf32"),
         "s8": lambda: case_tc("s8"), "trunc": case_trunc}

if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "case":
        CASES[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(CASES)
    for name in names:
        t = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "case", name], timeout=240)
            emit(case="done", name=name, rc=r.returncode, secs=round(time.time() - t, 1))
        except subprocess.TimeoutExpired:
            emit(case="TIMEOUT", name=name, secs=round(time.time() - t, 1))
