"""2-rank timeline of one row-panel step (torchrun --nproc-per-node 2 tools/trace_rowpanel.py [slices] [reserve])."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _libs

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = _libs.load_pkg()
rp = __import__("importlib").import_module(_libs.PKG + ".rowpanel")
comm = rp.nccl_comm_ptr(dist, dev)
N = 4096
sl = None
if len(sys.argv) > 1 and sys.argv[1] != "default":
    e = [0]
    for x in sys.argv[1].split(","):
        e.append(e[-1] + int(x))
    sl = list(zip(e[:-1], e[1:]))
plan = rp.RowPanelPlan(g, comm, N, N, N, 5, sl)
if len(sys.argv) > 2:
    g.lib.b200_rowpanel_set_reserve_sms(plan.handle, int(sys.argv[2]))
sets = [(torch.rand(N, N, device=dev), torch.rand(N, N, device=dev), torch.empty(N, N, device=dev)) for _ in range(3)]
for i in range(6):
    plan.run(*sets[i % 3])
torch.cuda.synchronize()
dist.barrier()
for mode in ("isolated", "back_to_back"):
    g.lib.b200_rowpanel_trace(plan.handle, 1)
    if mode == "back_to_back":
        g.lib.b200_rowpanel_trace(plan.handle, 0)
        for i in range(5):
            plan.run(*sets[i % 3])
        g.lib.b200_rowpanel_trace(plan.handle, 1)
    plan.run(*sets[0])
    buf = (C.c_float * 64)()
    n = g.lib.b200_rowpanel_trace_dump(plan.handle, buf, 64)
    v = [round(buf[i] * 1e3) for i in range(n)]
    out = [f"A_split_done={v[0]}us"]
    for j in range((n - 1) // 5):
        b = v[1 + 5 * j: 6 + 5 * j]
        out.append(f"slice{j}: bcast {b[0]}->{b[1]} visible {b[2]} split_done {b[3]} gemm_done {b[4]}")
    dist.barrier()
    for r in range(2):
        if r == rank:
            print(f"[rank {rank}] {mode} slices={plan.chunks}: " + " | ".join(out), flush=True)
        dist.barrier()
plan.close()
dist.destroy_process_group()
