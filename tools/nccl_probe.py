"""2-rank probe: NCCL broadcast bandwidth for the shapes bench.py uses, with the transport NCCL picked."""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
def bw(name, fn, nbytes, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    if rank == 0: print(f"{name}: {ms:.3f} ms  {nbytes / ms / 1e6:.1f} GB/s", flush=True)
B = torch.rand((4096, 4096), device=dev)
P = [torch.empty((4096, 1024), device=dev) for _ in range(4)]
bw("broadcast whole B 64MB", lambda: dist.broadcast(B, src=0), B.numel() * 4)
bw("broadcast 4 row chunks in place", lambda: [dist.broadcast(B[i * 1024:(i + 1) * 1024], src=0) for i in range(4)], B.numel() * 4)
bw("broadcast 4 separate 16MB buffers", lambda: [dist.broadcast(p, src=0) for p in P], B.numel() * 4)
comm = torch.cuda.Stream(device=dev)
def side():
    cur = torch.cuda.current_stream(); comm.wait_stream(cur)
    with torch.cuda.stream(comm):
        for i in range(4): dist.broadcast(B[i * 1024:(i + 1) * 1024], src=0)
    cur.wait_stream(comm)
bw("4 row chunks on a side stream", side, B.numel() * 4)
big = torch.empty((16384, 16384), device=dev)
bw("broadcast 1 GiB", lambda: dist.broadcast(big, src=0), big.numel() * 4, reps=3)
if rank == 0:
    print("can_device_access_peer(0,1):", torch.cuda.can_device_access_peer(0, 1))
dist.barrier(); dist.destroy_process_group()
