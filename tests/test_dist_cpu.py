"""World-size-2 gloo tests (CPU) of the multi-GPU host logic in eigd_b200/dist.py: the per-mode column
shards, the contiguous element ranges and the one packed all-gather of the sharded adjoint (columns of psi
plus the per-mode host scalars).  The sharded solvers themselves run on GPUs (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, N, nelems, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("eigd_dist", os.path.join(ROOT, "eigd_b200", "dist.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)                      # dist.py alone: no CUDA library needed
        sh = mod.ModeSharding()
        assert sh.rank == rank and sh.world == world
        full = torch.arange(n * N, dtype=torch.float64).reshape(n, N) * 0.5 + 1.0       # the "true" psi
        cols = sh.my_cols(N)
        assert list(cols) == list(range(rank, N, world))
        mine = full[:, cols].contiguous()                  # what this rank solved
        got = sh.allgather_cols(mine, N)
        assert got.shape == (n, N) and torch.equal(got, full)
        # element ranges: contiguous, disjoint, covering
        lo, hi = sh.my_range(nelems)
        ranges = [sh.my_range(nelems, r) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == nelems
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        assert max(h - l for l, h in ranges) - min(h - l for l, h in ranges) <= 1
        vec = torch.arange(nelems, dtype=torch.float64) ** 2
        assert torch.equal(sh.allgather_ranges(vec[lo:hi].clone(), nelems), vec)
        # per-mode host data (G columns, info, residual histories) ride in the SAME all-gather as the columns
        G = np.arange(N * N, dtype=float).reshape(N, N) + 0.25
        info = [int(3 + c) for c in cols]
        hist = [[1.0 / (k + 1 + c) for k in range(2 + int(c) % 3)] for c in cols]
        packed = sh.pack_mode_scalars(N, G[:, cols], info, hist)
        got2, extras = sh.allgather_cols(mine, N, extra=packed)
        assert torch.equal(got2, full)
        G2, info2, hist2 = sh.unpack_mode_scalars(N, extras)
        assert np.array_equal(G2, G)
        assert info2 == [3 + i for i in range(N)]
        assert hist2 == [[1.0 / (k + 1 + i) for k in range(2 + i % 3)] for i in range(N)]
        st = sh.collective_stats()
        assert st["calls"] == 3 and st["bytes"] > 0 and sh.collective_stats()["calls"] == 0
        # max over ranks of a per-rank timing, as bench.py does it
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t) == float(world)
        out[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,N,nelems", [(37, 10, 101), (16, 3, 7), (9, 1, 2)])
def test_mode_sharding_world2(n, N, nelems):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, N, nelems, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert sorted(out.keys()) == list(range(world))
