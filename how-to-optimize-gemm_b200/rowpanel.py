"""Row-panel sharding of C = A*B across ranks with one exchange step: the broadcast of B.

Rows of C are independent: rank i owns A[r0:r1, :] and C[r0:r1, :]; B (K x N, row-major) lives on
`src` and is broadcast to every rank inside the step.  Row-major B is contiguous by rows, so it is
sent as `chunks` row blocks B[k0:k1, :] straight out of / into the operand buffers — no packing copy
and no staging — back to back on a side stream (NCCL over NVLink/NVSwitch).

The broadcast is hidden behind the math by slicing K the same way: as soon as row block j has
landed, the local kernel runs C (+)= A[:, k0:k1] * B[k0:k1, :] (C-ABI `b200_gemm_f32_acc` for
j > 0) while blocks j+1.. are still in flight.  K-slicing costs nothing extra in the split-precision
modes (each slice splits its own part of A and B; two-level accumulation folds partial sums into C
anyway) and keeps every launch a full-size 2-D tile grid, unlike slicing N.

Measured on 2 x B200 (round 1, M = 8192, N = K = 4096, per-step ms): column-panel pipeline with a
packed copy and one GEMM per panel 0.888; broadcast then one GEMM 0.751; this K-sliced pipeline:
see profiles/.

The local kernel is injected (`gemm(A, B, out, accumulate)`): bench.py passes the C-ABI GEMM; the
CPU `gloo` test passes a host stand-in to check the partition / exchange plumbing only.
"""


def row_panel(rank, world, M):
    """Contiguous, balanced split of M rows: the first M % world ranks get one extra row."""
    base, extra = divmod(M, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def row_chunks(K, chunks, align=64):
    """Split of the K rows of B into contiguous blocks: an int gives that many balanced blocks, a sequence
    gives blocks proportional to its weights with boundaries rounded to `align` rows (whole k-blocks).  A small
    first block shortens the only part of the broadcast the math cannot hide behind: (1, 3, 4) on K = 4096
    is 512 / 1536 / 2048 rows."""
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


class RowPanelGemm:
    def __init__(self, gemm, dist, rank, world, K, N, chunks, device, dtype, src=0, pipeline=True):
        import torch
        self.torch, self.gemm, self.dist = torch, gemm, dist
        self.rank, self.world, self.src = rank, world, src
        self.chunks = row_chunks(K, chunks)
        self.cuda = device.type == "cuda"
        self.pipeline = pipeline
        self.comm = torch.cuda.Stream(device=device) if self.cuda else None

    def run(self, A_local, B, C_local):
        """One step.  `B` is the K x N operand buffer: the data on `src`, the receive buffer elsewhere."""
        torch, dist = self.torch, self.dist
        assert B.is_contiguous()
        if not self.cuda:
            for j, (k0, k1) in enumerate(self.chunks):
                dist.broadcast(B[k0:k1], src=self.src)
                if self.pipeline:
                    self.gemm(A_local[:, k0:k1], B[k0:k1], C_local, j > 0)
            if not self.pipeline:
                self.gemm(A_local, B, C_local, False)
            return C_local
        cur = torch.cuda.current_stream()
        self.comm.wait_stream(cur)                     # B (src) / the previous consumer of B (others) is ready
        events = []
        with torch.cuda.stream(self.comm):
            for k0, k1 in self.chunks:
                dist.broadcast(B[k0:k1], src=self.src)
                e = torch.cuda.Event()
                e.record(self.comm)
                events.append(e)
        if not self.pipeline:
            cur.wait_stream(self.comm)
            self.gemm(A_local, B, C_local, False)
            return C_local
        for j, ((k0, k1), e) in enumerate(zip(self.chunks, events)):
            cur.wait_event(e)                          # K-slice j of B is here; later slices still in flight
            self.gemm(A_local[:, k0:k1], B[k0:k1], C_local, j > 0)
        return C_local
