"""Row-panel sharding of C = A*B across ranks with one exchange step: the broadcast of B.

Rows of C are independent: rank i owns A[r0:r1, :] and C[r0:r1, :]; B (K x N) lives on `src` and is
broadcast in column panels so the local GEMM on panel j overlaps the transfer of panels j+1...
Column panels keep every output tile's K reduction local (no partial-sum exchange).

The local kernel is injected (`gemm(A, Bpanel, out)`): bench.py passes the C-ABI GEMM; the CPU `gloo`
test passes a host stand-in to check the partition / broadcast plumbing only.
"""


def row_panel(rank, world, M):
    """Contiguous, balanced split of M rows: the first M % world ranks get one extra row."""
    base, extra = divmod(M, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def column_panels(N, width):
    return [(c0, min(c0 + width, N)) for c0 in range(0, N, width)]


class RowPanelGemm:
    def __init__(self, gemm, dist, rank, world, K, N, panel, device, dtype, src=0):
        import torch
        self.torch, self.gemm, self.dist = torch, gemm, dist
        self.rank, self.world, self.src = rank, world, src
        self.panels = column_panels(N, panel)
        self.cuda = device.type == "cuda"
        # panel staging (src) / receive (others) buffers: contiguous K x width blocks NCCL can send
        self.bufs = [torch.empty((K, c1 - c0), device=device, dtype=dtype) for c0, c1 in self.panels]
        self.comm = torch.cuda.Stream(device=device) if self.cuda else None

    def run(self, A_local, B, C_local):
        """One step.  `B` is the full K x N matrix on `src` (ignored elsewhere)."""
        torch, dist = self.torch, self.dist
        if not self.cuda:
            for (c0, c1), buf in zip(self.panels, self.bufs):
                if self.rank == self.src:
                    buf.copy_(B[:, c0:c1])
                dist.broadcast(buf, src=self.src)
                self.gemm(A_local, buf, C_local[:, c0:c1])
            return C_local
        cur = torch.cuda.current_stream()
        self.comm.wait_stream(cur)
        events = []
        with torch.cuda.stream(self.comm):
            for (c0, c1), buf in zip(self.panels, self.bufs):
                if self.rank == self.src:
                    buf.copy_(B[:, c0:c1])            # pack the column panel (counted in the step)
                dist.broadcast(buf, src=self.src)
                e = torch.cuda.Event()
                e.record(self.comm)
                events.append(e)
        for (c0, c1), buf, e in zip(self.panels, self.bufs, events):
            cur.wait_event(e)                          # GEMM on panel j while j+1.. are in flight
            self.gemm(A_local, buf, C_local[:, c0:c1])
        return C_local
