"""Row-panel sharding of C = A*B across ranks with one exchange step: the broadcast of B.

Rows of C are independent: rank i owns A[r0:r1, :] and C[r0:r1, :]; B (K x N, row-major) lives on
`root` and is broadcast to every rank inside the step.  Row-major B is contiguous by rows, so it is
sent as K-slices (row blocks B[k0:k1, :]) straight out of / into the operand buffers — no packing copy
and no staging — back to back on a side stream (NCCL over NVLink/NVSwitch).  The broadcast is hidden
behind the math by slicing K the same way: A_i is split into its planes while slice 0 travels, and as
soon as slice j has landed the local kernel runs C (+)= A[:, k0:k1] * B[k0:k1, :] while slices j+1..
are still in flight.

The product path is C++ behind the C ABI (`b200_rowpanel_create` / `b200_gemm_f32_rowpanel` /
`b200_gemm_f32_rowpanel_host`, csrc/rowpanel.cuh); `RowPanelPlan` below is its ctypes binding and is what
bench.py and the GPU tests drive.  `RowPanelGemm` is the host-side MODEL of the same schedule with an
injected local kernel: the CPU `gloo` test (tests/test_rowpanel_gloo.py) runs it with a stand-in to
check the partition / exchange logic where no GPU exists.
"""
import ctypes as C


def row_panel(rank, world, M):
    """Contiguous, balanced split of M rows: the first M % world ranks get one extra row."""
    base, extra = divmod(M, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def default_slices(K, world, N=None):
    """The C++ plan's default K-slices (b200_rowpanel_create): one slice on a single rank or for a short K; up to 256 MB
    of B two slices weighted 1 : 3 (a shorter first slice shortens the only part of the exchange the math cannot hide
    behind; more slices cost a GEMM launch and an NCCL call each); beyond that equal slices of ~256 MB (at most 8).
    Boundaries are rounded up to 64 rows.  N defaults to K (square B)."""
    if world == 1 or K < 1024:
        return [(0, K)]
    N = K if N is None else N
    ns = max(2, min(8, -(-(K * N * 4) // (256 << 20))))
    if ns == 2:
        edges = [0, (K // 4 + 63) // 64 * 64, K]
    else:
        edges = [0] + [(K * j // ns + 63) // 64 * 64 for j in range(1, ns)] + [K]
    return list(zip(edges[:-1], edges[1:]))


def row_chunks(K, chunks, align=64):
    """Split of the K rows of B into contiguous blocks: an int gives that many balanced blocks, a sequence
    gives blocks proportional to its weights with boundaries rounded to `align` rows (whole k-blocks)."""
    if isinstance(chunks, int):
        chunks = max(1, min(chunks, K))
        return [row_panel(i, chunks, K) for i in range(chunks)]
    w = [float(x) for x in chunks if x > 0]
    tot, acc, edges = sum(w), 0.0, [0]
    for x in w[:-1]:
        acc += x
        e = int(round(K * acc / tot / align)) * align
        e = min(max(e, edges[-1]), K)
        if e > edges[-1]:
            edges.append(e)
    if edges[-1] < K:
        edges.append(K)
    return list(zip(edges[:-1], edges[1:]))


def nccl_comm_ptr(dist, device):
    """ncclComm_t of torch.distributed's default NCCL process group on `device`, as an int (the C ABI takes
    it as void*).  A tiny collective first makes sure the communicator exists."""
    import torch
    t = torch.zeros(1, device=device)
    dist.all_reduce(t)
    torch.cuda.synchronize(device)
    pg = dist.distributed_c10d._get_default_group()
    return int(pg._get_backend(device)._comm_ptr())


class RowPanelPlan:
    """ctypes binding of the C-ABI plan (include/b200gemm.h, multi-GPU section)."""

    def __init__(self, pkg, comm_ptr, m_local_max, n, k, mode, slices=None):
        self.pkg, self.lib = pkg, pkg.lib
        self.n, self.k = n, k
        self.handle = C.c_void_p()
        arr, cnt = None, 0
        if slices:
            rows = [k1 - k0 for k0, k1 in slices]
            arr, cnt = (C.c_int * len(rows))(*rows), len(rows)
        pkg._check(self.lib.b200_rowpanel_create(C.byref(self.handle), C.c_void_p(comm_ptr), m_local_max, n, k, mode, arr, cnt))
        import os
        if os.environ.get("B200_RESERVE_SMS"):      # tuning: SMs left to the exchange while a later slice is in flight
            pkg._check(self.lib.b200_rowpanel_set_reserve_sms(self.handle, int(os.environ["B200_RESERVE_SMS"])))
        b = (C.c_int * 17)()
        ns = self.lib.b200_rowpanel_slices(self.handle, b, 17)
        self.chunks = [(b[j], b[j + 1]) for j in range(ns)]

    def run(self, A_local, B, C_local, root=0, stream=None):
        """C_local = A_local * B.  B: the operand on `root`, the receive buffer elsewhere (valid everywhere after)."""
        pkg = self.pkg
        assert B.is_contiguous() and B.shape == (self.k, self.n)
        m = A_local.shape[0]
        pkg._check(self.lib.b200_gemm_f32_rowpanel(self.handle, m, self.n, self.k, A_local.data_ptr(), pkg._ld(A_local),
                                                   B.data_ptr(), pkg._ld(B), C_local.data_ptr(), pkg._ld(C_local), root,
                                                   pkg._stream_ptr(stream)))
        return C_local

    def run_host(self, A_local, B, C_local, root=0):
        """C_local += A_local * B with HOST tensors (pinned for full PCIe rate); B is read on `root` only."""
        pkg = self.pkg
        m = A_local.shape[0]
        pkg._check(self.lib.b200_gemm_f32_rowpanel_host(self.handle, m, self.n, self.k, A_local.data_ptr(), A_local.stride(0),
                                                        B.data_ptr() if B is not None else None,
                                                        B.stride(0) if B is not None else self.n,
                                                        C_local.data_ptr(), C_local.stride(0), root))
        return C_local

    def close(self):
        if getattr(self, "handle", None):
            self.lib.b200_rowpanel_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RowPanelGemm:
    """Host-side model of the plan's schedule (CPU tensors, any torch.distributed backend)."""

    def __init__(self, gemm, dist, rank, world, K, N, chunks=None, src=0, pipeline=True):
        self.gemm, self.dist = gemm, dist
        self.rank, self.world, self.src = rank, world, src
        self.chunks = default_slices(K, world, N) if chunks is None else row_chunks(K, chunks)
        self.pipeline = pipeline

    def run(self, A_local, B, C_local):
        """One step.  `B` is the K x N operand buffer: the data on `src`, the receive buffer elsewhere."""
        dist = self.dist
        assert B.is_contiguous() and not B.is_cuda, "the CUDA path is the C-ABI plan (RowPanelPlan)"
        for j, (k0, k1) in enumerate(self.chunks):
            dist.broadcast(B[k0:k1], src=self.src)
            if self.pipeline:
                self.gemm(A_local[:, k0:k1], B[k0:k1], C_local, j > 0)
        if not self.pipeline:
            self.gemm(A_local, B, C_local, False)
        return C_local
