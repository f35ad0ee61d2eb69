#!/usr/bin/env python
"""bench.py — the driver-facing benchmark of the GEMM hot path (contract: see DESIGN.md §Measurement).

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path on the host cores

Workload (BASELINE.json configs[1], headline point): fp32 SGEMM, row-major, M = 4096*N_gpus,
N = K = 4096.  At 1 GPU this is the 4096^3 point the reference quotes (cuda/output_MMult_cuda_12.m:29);
at N GPUs C is sharded by row panels (one 4096-row panel per rank, per-GPU work fixed => "weak"),
B lives on rank 0 and is broadcast over NVLink inside the timed region.  A "step" is one such GEMM.  value = 2*M*N*K / max-over-ranks time.

The JSON line also carries: modes (every fp32 precision mode at the same size with its measured
error against the oracle), sweep (the GFLOP/s-vs-N curve, also written in the reference's
output_*.m format under profiles/), roofline, cpu_baseline, e2e, clocks, gpu_launches.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
# stdout must carry exactly one JSON line.  NCCL prints its version banner with a C-level printf to
# fd 1 on the first communicator (seen on the 2-GPU box even with NCCL_DEBUG_FILE set), so the real
# stdout is set aside at start-up, fd 1 is pointed at stderr for everything else this process or its
# libraries print, and the JSON line alone is written to the saved descriptor.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
_JSON_FD = None


def _claim_stdout():
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(obj):
    sys.stdout.flush()
    data = (json.dumps(obj) + "\n").encode()
    fd = _JSON_FD if _JSON_FD is not None else 1
    while data:
        data = data[os.write(fd, data):]

sys.path.insert(0, os.path.join(ROOT, "tests"))

N0 = 4096                     # headline size
BCAST_CHUNKS = 2              # K-slices of the B broadcast / GEMM pipeline (measured on 2 GPUs: 2 -> 0.680 ms, 4 -> 0.722, 8 -> 0.883, no pipeline 0.742; GEMM alone 0.583)
MODE_NAMES = {0: "strict_ffma", 1: "tf32", 2: "bf16x3", 3: "bf16x2", 5: "f16x2_scaled"}
MODE_DTYPE = {0: "f32", 1: "tf32", 2: "bf16x3(split-f32)", 3: "bf16x2(split-f32)", 5: "f16x2(scaled split-f32)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, f"/tmp/b200_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "10"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.03)
        self.proc.terminate()
        self.proc.wait()
        self.f.close()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in open(self.path):
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], c[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU implementation of the path: cuda/REF_MMult.cpp (cblas_sgemm of the
    vendored OpenBLAS-0.2.20) from oracle/_ref/libref.so, all host threads; falls back to the
    oracle port when libref.so is absent."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import _libs
    cores = os.cpu_count() or 1
    M = N0 * args.gpus
    rng = np.random.default_rng(0)
    a = rng.uniform(-1, 1, (M, N0)).astype(np.float32)
    b = rng.uniform(-1, 1, (N0, N0)).astype(np.float32)
    c = np.zeros((M, N0), np.float32)
    if _libs.have_ref():
        r = _libs.load_ref()
        threads = min(cores, 128)
        r.openblas_set_num_threads(threads)
        fn = lambda: r.cuda_REF_MMult(M, N0, N0, _libs.P(a), N0, _libs.P(b), N0, _libs.P(c), N0)
        kind, what = "reference", "cuda/REF_MMult.cpp -> cblas_sgemm (vendored OpenBLAS-0.2.20, HASWELL kernels)"
    else:
        o = _libs.load_oracle()
        threads = o.oracle_get_threads()
        def fn():
            c[:] = 0
            o.oracle_ref_mmult_f32_fma_fast(M, N0, N0, _libs.P(a), N0, _libs.P(b), N0, _libs.P(c), N0)
        kind, what = "port", "oracle_ref_mmult_f32_fma_fast (naive REF_MMult arithmetic, row-parallel)"
    for _ in range(max(1, min(args.warmup, 3))):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = (time.perf_counter() - t0) / args.steps
    gf = 2.0 * M * N0 * N0 / dt / 1e9
    sample = f"{args.steps} full SGEMMs M={M} N=K={N0} ({what}), {threads} threads"
    _emit({
        "impl": "reference", "metric": "SGEMM GFLOP/s (square N=4096 point of the 256..4096 sweep)", "value": gf,
        "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"fp32 SGEMM row-major M={M} N=K={N0} (BASELINE configs[1], N=4096 point)"},
        "cpu_baseline": {"value": gf, "unit": "GFLOP/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ------------------------------------------------------------------------------------------------
def cpu_baseline(o, budget_s=12.0):
    """Timed beside the GPU number on this box's host cores (rank 0, N=1): the reference's OpenBLAS
    REF_MMult on the full 4096^3 problem, and the naive REF_MMult arithmetic on a row subset."""
    import numpy as np
    import _libs
    rng = np.random.default_rng(1)
    a = rng.uniform(-1, 1, (N0, N0)).astype(np.float32)
    b = rng.uniform(-1, 1, (N0, N0)).astype(np.float32)
    c = np.zeros((N0, N0), np.float32)
    out = {}
    cores = os.cpu_count() or 1
    if _libs.have_ref():
        r = _libs.load_ref()
        threads = min(cores, 128)
        r.openblas_set_num_threads(threads)
        r.cuda_REF_MMult(N0, N0, N0, _libs.P(a), N0, _libs.P(b), N0, _libs.P(c), N0)
        reps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < budget_s * 0.6 and reps < 50:
            r.cuda_REF_MMult(N0, N0, N0, _libs.P(a), N0, _libs.P(b), N0, _libs.P(c), N0)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        out = {"value": 2.0 * N0 ** 3 / dt / 1e9, "unit": "GFLOP/s", "cores": threads, "kind": "reference",
               "sample": f"{reps} full 4096^3 cblas_sgemm calls via cuda/REF_MMult.cpp (OpenBLAS-0.2.20)"}
    rows = 64
    cs = np.zeros((rows, N0), np.float32)
    t0 = time.perf_counter()
    o.oracle_ref_mmult_f32_fma(rows, N0, N0, _libs.P(a), N0, _libs.P(b), N0, _libs.P(cs), N0)
    dt = time.perf_counter() - t0
    naive = {"value": 2.0 * rows * N0 * N0 / dt / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": "port",
             "sample": f"naive REF_MMult loop nest (aarch64/REF_MMult.cpp:18-28) on {rows} of 4096 rows, extrapolated"}
    if not out:
        out = dict(naive)
    out["naive_ref_mmult"] = naive
    return out


_T0 = time.time()


def _phase(name):
    if os.environ.get("B200_BENCH_TRACE"):
        print(f"[bench +{time.time() - _T0:6.1f}s] {name}", file=sys.stderr, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--mode", type=int, default=-1, help="fp32 precision mode of the headline (default: library default)")
    ap.add_argument("--no-extras", action="store_true", help="skip sweep / modes / cpu_baseline (quick runs)")
    ap.add_argument("--workload", default="headline", choices=["headline", "c5"],
                    help="headline: M=4096*gpus, N=K=4096 (weak).  c5: BASELINE configs[4], M=N=K=16384 "
                         "row-panel sharded over the ranks (strong); not the driver's default")
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import _libs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _phase("torch/nccl up")
    g = _libs.load_pkg()            # raises if libb200gemm.so is missing: no fallback
    mode = args.mode if args.mode >= 0 else g.lib.b200_gemm_default_f32_mode()
    dev = torch.device("cuda", local)
    K = N = N0
    Mloc = N0
    scaling = "weak"
    if args.workload == "c5":
        K = N = 16384
        Mloc = 16384 // world
        scaling = "strong"

    # ---- inputs resident in HBM: R rotating sets so consecutive steps never hit a warm L2 --------
    R = 3 if args.workload == "headline" else 1     # c5 operands (1 GiB each) exceed L2 on their own
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    sets = []
    for _ in range(R):
        A = torch.rand((Mloc, K), device=dev, generator=gen) * 2 - 1
        B = torch.rand((K, N), device=dev, generator=gen) * 2 - 1 if (rank == 0 or world == 1) else torch.empty((K, N), device=dev)
        Cm = torch.empty((Mloc, N), device=dev)
        sets.append((A, B, Cm, None))
    rp = None
    if world > 1:
        rowpanel = __import__("importlib").import_module(_libs.PKG + ".rowpanel")
        slices = BCAST_CHUNKS if args.workload == "headline" else 4
        if os.environ.get("B200_BCAST_SLICES"):        # tuning: "3" = three balanced slices, "1,3,4" = weighted
            v = [int(x) for x in os.environ["B200_BCAST_SLICES"].split(",")]
            slices = v[0] if len(v) == 1 else tuple(v)
        rp = rowpanel.RowPanelGemm(lambda a, b, out, acc: g.gemm_f32(a, b, out=out, mode=mode, accumulate=acc), dist, rank, world,
                                   K, N, slices, dev, torch.float32)

    def step(i):
        A, B, Cm, _ = sets[i % R]
        if world == 1:
            g.gemm_f32(A, B, out=Cm, mode=mode)
        else:
            rp.run(A, B, Cm)        # NCCL broadcast of B in K-slices (in place) pipelined with C (+)= A[:,ks] * B[ks,:]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()              # nvidia-smi needs ~0.1 s to start: launch it ahead of the warm-up
    for i in range(max(args.warmup, 3)):
        step(i)
    if rank == 0:
        time.sleep(0.15)             # let the sampler come up; BEFORE the barrier so all ranks start together
    barrier()
    l0 = g.launch_count()
    g.lib.b200_gemm_debug_kernel_timing(1)      # event pair around every dominant-kernel launch, same stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = g.launch_count() - l0
    kern_ms_sum, kern_launches = g.kernel_time_ms()
    g.lib.b200_gemm_debug_kernel_timing(0)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms = ms_total / args.steps
    flops_step = 2.0 * (Mloc * world) * N * K
    value = flops_step / (ms * 1e-3) / 1e9
    kernel_name = g.last_kernel()

    _phase("timed region done")
    # ---- e2e: the host-pointer plug-in call (9-arg MY_MMult contract, C += A*B), copies inside ----
    e2e_steps = max(3, min(args.steps, 8)) if args.workload == "headline" else 2
    hA = torch.empty((Mloc, K), dtype=torch.float32).pin_memory().uniform_(-1, 1)
    hB = torch.empty((K, N), dtype=torch.float32).pin_memory().uniform_(-1, 1)
    hC = torch.zeros((Mloc, N), dtype=torch.float32).pin_memory()
    def e2e_step():
        rc = g.lib.b200_gemm_f32_host(Mloc, N, K, hA.data_ptr(), K, hB.data_ptr(), N, hC.data_ptr(), N, mode)
        assert rc == 0, rc
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()                      # synchronous: returns when C is back in host memory
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e = {"value": flops_step / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s",
           "h2d_bytes_per_step": (Mloc * K + K * N + Mloc * N) * 4, "d2h_bytes_per_step": Mloc * N * 4,
           "ms_per_step": e2e_ms, "api": "b200_gemm_f32_host (9-arg MY_MMult contract, pinned host buffers, per rank)"}

    if world > 1:
        dist.destroy_process_group()        # every rank, right after the last collective
    if rank != 0:
        return

    _phase("e2e done")
    pk = peaks()
    out = {
        "metric": "SGEMM GFLOP/s (square N=4096 point of the 256..4096 sweep)", "value": value, "unit": "GFLOP/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": MODE_DTYPE.get(mode, str(mode)), "data": "synthetic",
        "config": {"workload": f"fp32 SGEMM row-major M={Mloc * world} N=K={N} (BASELINE configs[{1 if args.workload == 'headline' else 4}]); "
                               f"C row-panel sharded, B broadcast from rank 0 inside every step as {len(rp.chunks) if rp else 0} K-slices {[k1 - k0 for k0, k1 in rp.chunks] if rp else ""} (NCCL, in place) pipelined with the K-sliced GEMM" if world > 1 else
                               f"fp32 SGEMM row-major M=N=K={N0} (BASELINE configs[1], N=4096 point)",
                   "precision_mode": MODE_NAMES.get(mode, str(mode)), "kernel": kernel_name,
                   "l2": f"{R} rotating input/output sets of {3 * N0 * N0 * 4 / 1e6:.0f} MB each (> 126 MB L2 between reuses)",
                   "inputs": "uniform(-1,1), row-major, lda=k ldb=n ldc=n (cuda/test_MMult.cpp:62)"},
        "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e,
        "published_reference": {"MMult_cuda_12 @4096 on RTX 3090": 21410.87, "note": "other hardware; BASELINE.json.published is {}"},
    }
    # roofline of the dominant kernel: its own launch durations (CUDA events on the launching stream,
    # recorded inside the timed region); algorithmic flops = 2*M*N*K, no credit for the 6 split passes
    if world == 1 and args.workload == "headline":
        kern_ms = kern_ms_sum / max(kern_launches, 1)
        achieved = 2.0 * N0 ** 3 / (kern_ms * 1e-3) / 1e12
        out["roofline"] = {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                           "frac": achieved / pk["bf16_tflops"], "traffic": None,
                           "kernel_ms": kern_ms, "kernel_launches_timed": kern_launches,
                           "kernel_share_of_step": kern_ms / ms,
                           "tensor_pipe_flops_per_launch": 2.0 * N0 ** 3 * {2: 6, 3: 3, 5: 3}.get(mode, 1),
                           "tensor_pipe_frac": achieved * {2: 6, 3: 3, 5: 3}.get(mode, 1) / ({1: 0.5}.get(mode, 1.0) * pk["bf16_tflops"]),
                           "peak_source": pk["source"] + ", burst bf16; sustained " + str(pk["bf16_tflops_sustained"]),
                           "frac_of_sustained": achieved / pk["bf16_tflops_sustained"] if pk["bf16_tflops_sustained"] else None,
                           "algorithmic_flops_per_launch": 2.0 * N0 ** 3,
                           "algorithmic_bytes_per_launch": 3 * N0 * N0 * 4,
                           "achieved_hbm_gbs": 3 * N0 * N0 * 4 / (ms * 1e-3) / 1e9,
                           "kernel": kernel_name}
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                # bytes per launch (dram read+write) of this kernel at this size from the committed
                # `ncu --set full` capture (profiles/, tools/summarize_ncu.py); null if not captured
                out["roofline"]["traffic"] = json.load(open(tp)).get(f"{kernel_name}@{N0}")
            except Exception:
                pass

    if not args.no_extras and world == 1 and args.workload == "headline":
        o = _libs.load_oracle()
        A, B, Cm, _ = sets[0]
        # ---- every precision mode at the headline size, with its error against the oracle --------
        rows = torch.arange(0, N0, 67, device=dev)[:48]
        a_np, b_np = A[rows].cpu().numpy(), B.cpu().numpy()
        truth = _libs.ref_f64(o, a_np, b_np)
        ref_naive = _libs.ref_f32_fma(o, a_np, b_np)
        modes = {}
        for md, name in MODE_NAMES.items():
            try:
                g.gemm_f32(A, B, out=Cm, mode=md)
            except g.B200GemmError:
                continue
            kn = g.last_kernel()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for i in range(10):
                A2, B2, C2, _ = sets[i % R]
                g.gemm_f32(A2, B2, out=C2, mode=md)
            e.record()
            torch.cuda.synchronize()
            t_ms = s.elapsed_time(e) / 10
            g.gemm_f32(A, B, out=Cm, mode=md)
            got = Cm[rows].cpu().numpy()
            modes[name] = {"gflops": 2.0 * N0 ** 3 / t_ms / 1e6, "ms": t_ms, "kernel": kn,
                           "max_abs_err_vs_f64": float(np.abs(got - truth).max()),
                           "max_rel_err_vs_maxabs": float(np.abs(got - truth).max() / np.abs(truth).max()),
                           "max_abs_diff_vs_REF_MMult_naive": float(np.abs(got - ref_naive).max()),
                           "bit_exact_vs_REF_MMult_naive": bool(np.array_equal(got, ref_naive)),
                           "frac_of_bf16_peak": 2.0 * N0 ** 3 / t_ms / 1e9 / pk["bf16_tflops"]}
        out["modes"] = modes
        fp32_peak = 2 * 128 * torch.cuda.get_device_properties(dev).multi_processor_count * (clocks["sm_max_mhz"] or 1965.0) * 1e6 / 1e12
        out["fp32_cuda_core_peak_tflops"] = fp32_peak
        if "strict_ffma" in modes:
            modes["strict_ffma"]["frac_of_fp32_cuda_core_peak"] = modes["strict_ffma"]["gflops"] / 1e3 / fp32_peak
        # ---- GFLOP/s-vs-N curve in the reference's output_*.m format ------------------------------
        sweep, sweep_kernels = [], []
        sweep_mode = args.mode if args.mode >= 0 else g.F32_AUTO    # the library default, size heuristic included
        for n in range(256, 4097, 256):
            a = torch.rand((n, n), device=dev) * 2 - 1
            b = torch.rand((n, n), device=dev) * 2 - 1
            c = torch.empty((n, n), device=dev)
            for _ in range(3):
                g.gemm_f32(a, b, out=c, mode=sweep_mode)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(20):            # NREPEATS = 20 back-to-back launches (cuda/parameters.h:24)
                g.gemm_f32(a, b, out=c, mode=sweep_mode)
            e.record()
            torch.cuda.synchronize()
            sweep.append([n, round(2.0 * n ** 3 / (s.elapsed_time(e) / 20) / 1e6, 2)])
            sweep_kernels.append(g.last_kernel())
        out["sweep"] = sweep
        out["sweep_kernels"] = sweep_kernels       # AUTO takes the single-launch strict kernel up to ~512^3
        out["cpu_baseline"] = cpu_baseline(o)
    _phase("extras done")
    _emit(out)


if __name__ == "__main__":
    main()
