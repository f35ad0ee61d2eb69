"""2-rank probe of the bench.py step components."""
import os, sys, importlib
import torch, torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import _libs
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
g = _libs.load_pkg()
rowpanel = importlib.import_module(_libs.PKG + ".rowpanel")
N = 4096
A = torch.rand((N, N), device=dev) - 0.5
B = torch.rand((N, N), device=dev) - 0.5
C = torch.empty((N, N), device=dev)
def timeit(name, fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    if rank == 0: print(f"{name}: {s.elapsed_time(e) / reps:.3f} ms", flush=True)
for mode in (2, 1):
    timeit(f"gemm only mode {mode}", lambda: g.gemm_f32(A, B, out=C, mode=mode))
    timeit("broadcast only", lambda: dist.broadcast(B, src=0))
    timeit(f"broadcast + gemm same stream mode {mode}", lambda: (dist.broadcast(B, src=0), g.gemm_f32(A, B, out=C, mode=mode)))
    for chunks, pipe in ((4, False), (2, True), (4, True), (8, True)):
        rp = rowpanel.RowPanelGemm(lambda a, b, out, acc: g.gemm_f32(a, b, out=out, mode=mode, accumulate=acc), dist, rank, world, N, N, chunks, dev, torch.float32, pipeline=pipe)
        timeit(f"RowPanelGemm.run mode {mode} chunks {chunks} pipeline {pipe}", lambda: rp.run(A, B, C))
dist.barrier(); dist.destroy_process_group()
