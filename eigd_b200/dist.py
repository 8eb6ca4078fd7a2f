"""Multi-GPU sharding of the stages that partition naturally.

One process per GPU (``torch.distributed``, NCCL over NVLink; gloo on CPU for the host-logic
tests).  The eigensolve is a sequential Krylov recurrence on one factor and is replicated on every
rank (deterministic: same start vector, same kernels, same sums).  What is sharded:

  * per-mode adjoint solves -- mode i goes to rank i mod world (round robin balances the
    per-mode iteration counts); every rank needs the full Phi for the projector but only its own
    columns of Phib / psi.  The solved columns are exchanged with one all-gather of dense
    (n x ceil(N/world)) fp64 slabs; the N x N coupling matrix G with a second, tiny one.
  * element ranges of the df/dx reduction -- contiguous element blocks, one all-gather of the
    per-element results.

There is no collective inside a Krylov iteration: the reference's modes are uncoupled
(eigd/eigenvector_derivatives.py:1189-1217 with update_guess=False), so ranks only meet at the
two gathers.
"""
import numpy as np
import torch
import torch.distributed as dist


class ModeSharding:
    def __init__(self, group=None):
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("ModeSharding needs an initialised torch.distributed process group")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    # ---- partitions --------------------------------------------------------------------------
    def my_cols(self, N, rank=None):
        r = self.rank if rank is None else rank
        return np.arange(r, N, self.world)

    def max_cols(self, N):
        return (N + self.world - 1) // self.world

    def my_range(self, count, rank=None):
        """Contiguous block [lo, hi) of ``count`` items (element ranges)."""
        r = self.rank if rank is None else rank
        base, extra = divmod(count, self.world)
        lo = r * base + min(r, extra)
        return lo, lo + base + (1 if r < extra else 0)

    # ---- collectives ---------------------------------------------------------------------------
    def allgather_cols(self, X_sub, N, transpose=None):
        """Reassemble the (n, N) row-major matrix whose columns ``my_cols(N)`` are the columns of
        ``X_sub`` (n, Ns) on each rank.  ``transpose(src2d, dst2d)`` is the strided copy used for
        the layout change (device.copy2d on CUDA; torch on CPU)."""
        n = X_sub.shape[0]
        nmax = self.max_cols(N)
        send = torch.zeros((nmax, n), dtype=X_sub.dtype, device=X_sub.device)      # vector-major slab
        ns = X_sub.shape[1]
        if ns:
            _copy(X_sub, send[:ns].T, transpose)
        recv = torch.empty((self.world, nmax, n), dtype=X_sub.dtype, device=X_sub.device)
        dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=self.group)
        out = torch.empty((n, N), dtype=X_sub.dtype, device=X_sub.device)
        for r in range(self.world):
            cols = self.my_cols(N, r)
            if len(cols):
                # columns r, r+world, ... of out <- rows of the slab
                _copy(recv[r, :len(cols)].T, out[:, r::self.world], transpose)
        return out

    def allgather_ranges(self, x_part, count):
        """Concatenate contiguous per-rank blocks of a length-``count`` vector."""
        sizes = [self.my_range(count, r) for r in range(self.world)]
        nmax = max(hi - lo for lo, hi in sizes)
        send = torch.zeros(nmax, dtype=x_part.dtype, device=x_part.device)
        send[: x_part.shape[0]].copy_(x_part)
        recv = torch.empty((self.world, nmax), dtype=x_part.dtype, device=x_part.device)
        dist.all_gather_into_tensor(recv.view(-1), send, group=self.group)
        out = torch.empty(count, dtype=x_part.dtype, device=x_part.device)
        for r, (lo, hi) in enumerate(sizes):
            out[lo:hi].copy_(recv[r, : hi - lo])
        return out

    def allgather_object(self, obj):
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def merge_cols_host(self, parts, N):
        """Host (numpy) counterpart for small per-column arrays: parts[r] has shape (..., len(my_cols(N, r)))."""
        first = np.asarray(parts[0])
        out = np.zeros(first.shape[:-1] + (N,), dtype=first.dtype)
        for r, p in enumerate(parts):
            cols = self.my_cols(N, r)
            if len(cols):
                out[..., cols] = np.asarray(p)
        return out


def _copy(src, dst, transpose):
    if transpose is not None and src.is_cuda:
        transpose(src, dst)
    else:
        dst.copy_(src)
