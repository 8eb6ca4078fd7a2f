"""Multi-GPU sharding of the stages that partition naturally.

One process per GPU (``torch.distributed``, NCCL over NVLink; gloo on CPU for the host-logic
tests).  The eigensolve is a sequential Krylov recurrence on one factor and is replicated on every
rank (deterministic: same start vector, same kernels, same sums).  What is sharded:

  * per-mode adjoint solves -- mode i goes to rank i mod world (round robin balances the
    per-mode iteration counts); every rank needs the full Phi for the projector but only its own
    columns of Phib / psi.  The solved columns are exchanged with ONE packed fp64 all-gather: dense
    (ceil(N/world) x n) slabs followed by the per-mode scalars (columns of the N x N coupling matrix G,
    convergence flags, residual histories) -- no pickled objects on the path.
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
        self._events = []

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
    def allgather_cols(self, X_sub, N, transpose=None, extra=None):
        """Reassemble the (n, N) row-major matrix whose columns ``my_cols(N)`` are the columns of
        ``X_sub`` (n, Ns) on each rank.  ``transpose(src2d, dst2d)`` is the strided copy used for
        the layout change (device.copy2d on CUDA; torch on CPU).

        ``extra``: a 1-D float64 host array of the SAME length on every rank (per-mode scalars: columns of G,
        convergence flags, residual histories).  It rides in the same slab, so the sharded adjoint meets in ONE
        packed fp64 all-gather; returns (out, extras) with extras[r] = rank r's array."""
        n = X_sub.shape[0]
        nmax = self.max_cols(N)
        L = 0 if extra is None else int(len(extra))
        slab = nmax * n + L
        send = torch.zeros(slab, dtype=X_sub.dtype, device=X_sub.device)           # vector-major slab + scalars
        ns = X_sub.shape[1]
        if ns:
            _copy(X_sub, send[: nmax * n].view(nmax, n)[:ns].T, transpose)
        if L:
            send[nmax * n:].copy_(torch.as_tensor(np.ascontiguousarray(extra, dtype=np.float64)), non_blocking=False)
        recv = torch.empty((self.world, slab), dtype=X_sub.dtype, device=X_sub.device)
        self._timed_allgather(recv, send)
        out = torch.empty((n, N), dtype=X_sub.dtype, device=X_sub.device)
        for r in range(self.world):
            cols = self.my_cols(N, r)
            if len(cols):
                # columns r, r+world, ... of out <- rows of the slab
                _copy(recv[r, : nmax * n].view(nmax, n)[:len(cols)].T, out[:, r::self.world], transpose)
        if extra is None:
            return out
        return out, recv[:, nmax * n:].cpu().numpy()

    def _timed_allgather(self, recv, send):
        """all_gather_into_tensor with CUDA events around it (device time of the collective, resolved lazily by
        ``collective_stats``) and a byte count -- bench.py reports both."""
        nbytes = recv.numel() * recv.element_size()
        if send.is_cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=self.group)
            e1.record()
            self._events.append((e0, e1, nbytes))
        else:
            dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=self.group)
            self._events.append((None, None, nbytes))

    def collective_stats(self, reset=True):
        """-> {"calls", "ms", "bytes"} of the all-gathers issued since the last reset (device time, this rank)."""
        ms = 0.0
        for e0, e1, _ in self._events:
            if e0 is not None:
                e1.synchronize()
                ms += e0.elapsed_time(e1)
        out = {"calls": len(self._events), "ms": ms, "bytes": int(sum(b for _, _, b in self._events))}
        if reset:
            self._events = []
        return out

    def allgather_ranges(self, x_part, count):
        """Concatenate contiguous per-rank blocks of a length-``count`` vector."""
        sizes = [self.my_range(count, r) for r in range(self.world)]
        nmax = max(hi - lo for lo, hi in sizes)
        send = torch.zeros(nmax, dtype=x_part.dtype, device=x_part.device)
        send[: x_part.shape[0]].copy_(x_part)
        recv = torch.empty((self.world, nmax), dtype=x_part.dtype, device=x_part.device)
        self._timed_allgather(recv, send)
        out = torch.empty(count, dtype=x_part.dtype, device=x_part.device)
        for r, (lo, hi) in enumerate(sizes):
            out[lo:hi].copy_(recv[r, : hi - lo])
        return out

    # ---- per-mode host scalars of the Krylov solvers, packed for allgather_cols(extra=...) -------------
    HIST = 128          # residual-history slots per mode (sibk / pgmres stop at maxiter = 50, pcpg at 100)

    def pack_mode_scalars(self, N, G, info, hist):
        """G: (N, ns) columns of the coupling matrix of this rank's modes; info, hist: per local mode."""
        nmax, H = self.max_cols(N), self.HIST
        buf = np.zeros(nmax * (N + 2 + H))
        ns = 0 if G is None else G.shape[1]
        g = buf[: nmax * N].reshape(nmax, N)
        if ns:
            g[:ns] = np.asarray(G, dtype=float).T
        meta = buf[nmax * N:].reshape(nmax, 2 + H)
        for c in range(min(ns, len(info))):
            h = [float(v) for v in (hist[c] if c < len(hist) else [])][:H]
            meta[c, 0] = float(info[c])
            meta[c, 1] = len(h)
            meta[c, 2: 2 + len(h)] = h
        return buf

    def unpack_mode_scalars(self, N, extras, info_type=int):
        """-> (G (N, N), info[N], hist[N]) from the gathered per-rank buffers of pack_mode_scalars."""
        nmax, H = self.max_cols(N), self.HIST
        G = np.zeros((N, N))
        info, hist = [0] * N, [[] for _ in range(N)]
        for r in range(self.world):
            buf = np.asarray(extras[r])
            g = buf[: nmax * N].reshape(nmax, N)
            meta = buf[nmax * N:].reshape(nmax, 2 + H)
            for c, i in enumerate(self.my_cols(N, r)):
                G[:, i] = g[c]
                info[i] = info_type(meta[c, 0])
                hist[i] = [float(v) for v in meta[c, 2: 2 + int(meta[c, 1])]]
        return G, info, hist

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
