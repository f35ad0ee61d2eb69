"""The N>1 path on CPU: world_size-2 `gloo` run of the row-panel shard + row-chunk broadcast of B
(how-to-optimize-gemm_b200/rowpanel.py).  The local kernel is a host stand-in (the oracle's
REF_MMult arithmetic) — this checks partitioning and the exchange step, not the CUDA kernels."""
import os
import socket
import sys

import numpy as np
import pytest

import _libs

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, M, N, K, panel, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import importlib
    import _libs as L
    sys.path.insert(0, L.ROOT)
    rowpanel = importlib.import_module(L.PKG + ".rowpanel")   # pure-python module: does not need the GPU library
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = L.load_oracle()
    o.oracle_set_threads(1)

    def host_gemm(a, b, out, accumulate):
        c0 = np.ascontiguousarray(out.numpy()) if accumulate else None     # C += A*B: chain continues from C
        c = L.ref_f32_fma(o, np.ascontiguousarray(a.numpy()), np.ascontiguousarray(b.numpy()), c0)
        out.copy_(torch.from_numpy(c))

    A = torch.from_numpy(L.gen_f32(o, M, K, 100))
    B = torch.from_numpy(L.gen_f32(o, K, N, 200)) if rank == 0 else torch.full((K, N), float("nan"))
    r0, r1 = rowpanel.row_panel(rank, world, M)
    rp = rowpanel.RowPanelGemm(host_gemm, dist, rank, world, K, N, panel)
    C = torch.full((r1 - r0, N), float("nan"))
    rp.run(A[r0:r1], B, C)
    np.save(os.path.join(out_dir, f"c_{rank}.npy"), C.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("M,N,K,panel", [(37, 50, 29, 3), (64, 96, 40, 4)])
def test_rowpanel_world2(tmp_path, oracle, M, N, K, panel):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), M, N, K, panel, str(tmp_path)), nprocs=world, join=True)
    C = np.concatenate([np.load(tmp_path / f"c_{r}.npy") for r in range(world)], axis=0)
    a, b = _libs.gen_f32(oracle, M, K, 100), _libs.gen_f32(oracle, K, N, 200)
    assert np.array_equal(C, _libs.ref_f32_fma(oracle, a, b))


@pytest.mark.parametrize("world,M,N,K,panel", [(4, 41, 24, 200, (1, 3, 4)), (3, 10, 16, 130, 2), (2, 33, 40, 1100, None)])
def test_rowpanel_world_gt2(tmp_path, oracle, world, M, N, K, panel):
    """More ranks than the GPU validation had (the driver's scaling run goes to 8), weighted K-slices."""
    mp.spawn(_worker, args=(world, _free_port(), M, N, K, panel, str(tmp_path)), nprocs=world, join=True)
    C = np.concatenate([np.load(tmp_path / f"c_{r}.npy") for r in range(world)], axis=0)
    a, b = _libs.gen_f32(oracle, M, K, 100), _libs.gen_f32(oracle, K, N, 200)
    assert np.array_equal(C, _libs.ref_f32_fma(oracle, a, b))


def test_partition_helpers():
    import importlib
    sys.path.insert(0, _libs.ROOT)
    rowpanel = importlib.import_module(_libs.PKG + ".rowpanel")
    for M in (1, 7, 16384, 4097):
        for world in (1, 2, 3, 4, 8):
            spans = [rowpanel.row_panel(r, world, M) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == M
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert rowpanel.row_chunks(4096, 4) == [(0, 1024), (1024, 2048), (2048, 3072), (3072, 4096)]
    assert rowpanel.row_chunks(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert rowpanel.row_chunks(2, 8) == [(0, 1), (1, 2)]
    assert rowpanel.row_chunks(4096, (1, 3, 4)) == [(0, 512), (512, 2048), (2048, 4096)]
    assert rowpanel.row_chunks(200, (1, 3, 4)) == [(0, 128), (128, 200)]        # 25 rows round to no block
    # the C++ plan's default schedule (b200_rowpanel_create), mirrored by default_slices
    assert rowpanel.default_slices(4096, 1) == [(0, 4096)]
    assert rowpanel.default_slices(4096, 2) == [(0, 1024), (1024, 4096)]
    assert rowpanel.default_slices(16384, 8) == [(0, 4096), (4096, 8192), (8192, 12288), (12288, 16384)]     # 1 GiB: 4 x 256 MB
    assert rowpanel.default_slices(8192, 2) == [(0, 2048), (2048, 8192)]                                     # 256 MB: still two slices
    assert len(rowpanel.default_slices(65536, 4, 16384)) == 8
    assert rowpanel.default_slices(1000, 4) == [(0, 1000)]
    assert rowpanel.default_slices(1100, 2) == [(0, 320), (320, 1100)]
    for K in (1, 63, 64, 200, 4096, 16384):
        for w in ((1, 3, 4), (1, 1), (5,), (1, 2, 2, 3)):
            ch = rowpanel.row_chunks(K, w)
            assert ch[0][0] == 0 and ch[-1][1] == K and all(a < b for a, b in ch)
            assert all(ch[i][1] == ch[i + 1][0] for i in range(len(ch) - 1))
