"""Row-panel sharding of C = A*B across ranks with one exchange step: the broadcast of B.

Rows of C are independent: rank i owns A[r0:r1, :] and C[r0:r1, :]; B (K x N, row-major) lives on
`src` and is broadcast to every rank inside the step.  Row-major B is contiguous by rows, so it is
sent as `chunks` row blocks B[k0:k1, :] straight out of / into the operand buffers — no packing copy
and no staging.  The chunks go out back to back on a side stream (NCCL pipelines them over
NVLink/NVSwitch); the local GEMM is ONE call on the full panel once the last chunk has landed.

Measured on 2 x B200 (round 1): a finer pipeline (GEMM per column panel as it lands) lost more to
repeated fp32->bf16 splitting of A and to wave quantisation of the small per-panel GEMMs (64 pair-tiles
on 74 pairs) than the ~0.1 ms of broadcast it hid; see DESIGN.md §7.

The local kernel is injected (`gemm(A, B, out)`): bench.py passes the C-ABI GEMM; the CPU `gloo` test
passes a host stand-in to check the partition / broadcast plumbing only.
"""


def row_panel(rank, world, M):
    """Contiguous, balanced split of M rows: the first M % world ranks get one extra row."""
    base, extra = divmod(M, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def row_chunks(K, chunks):
    """Balanced split of the K rows of B into `chunks` contiguous blocks."""
    chunks = max(1, min(chunks, K))
    return [row_panel(i, chunks, K) for i in range(chunks)]


class RowPanelGemm:
    def __init__(self, gemm, dist, rank, world, K, N, chunks, device, dtype, src=0):
        import torch
        self.torch, self.gemm, self.dist = torch, gemm, dist
        self.rank, self.world, self.src = rank, world, src
        self.chunks = row_chunks(K, chunks)
        self.cuda = device.type == "cuda"
        self.comm = torch.cuda.Stream(device=device) if self.cuda else None

    def run(self, A_local, B, C_local):
        """One step.  `B` is the K x N operand buffer: the data on `src`, the receive buffer elsewhere."""
        torch, dist = self.torch, self.dist
        assert B.is_contiguous()
        if not self.cuda:
            for k0, k1 in self.chunks:
                dist.broadcast(B[k0:k1], src=self.src)
            self.gemm(A_local, B, C_local)
            return C_local
        cur = torch.cuda.current_stream()
        self.comm.wait_stream(cur)                     # B (src) / the previous consumer of B (others) is ready
        with torch.cuda.stream(self.comm):
            for k0, k1 in self.chunks:
                dist.broadcast(B[k0:k1], src=self.src)
        cur.wait_stream(self.comm)
        self.gemm(A_local, B, C_local)
        return C_local
