"""Sharded vs unsharded parity on real GPUs (needs >= 2 devices; skipped on a 1-GPU box).

Two NCCL ranks (torch.multiprocessing, one process per GPU) run the same design twice through the device drivers
of eigd_b200.topo: once unsharded, once with ``dist.ModeSharding`` (eigensolve replicated and seeded, per-mode
adjoint solves on rank i mod world, element ranges of df/dx per rank, ONE packed all-gather per adjoint solve).
Gradient and adjoint vectors must agree to 1e-12 relative for every adjoint method and for the normal and the
buckling (indefinite, refined) operators.  SURVEY.md section 8e; eigd_b200/dist.py."""
import os
import socket
import sys
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-12


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        from eigd_b200 import device as D, topo as T
        from eigd_b200.dist import ModeSharding
        D.init("cuda:%d" % rank)
        shard = ModeSharding()
        rel = lambda a, b: float((a - b).abs().max() / b.abs().max())     # noqa: E731

        def grad(model, seed_fn, sh):
            model.sharding = sh
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                model.initialize()
            model.initialize_adjoint()
            seed_fn(model)
            model.finalize_adjoint()
            return model.xb.clone(), (model.psi if hasattr(model, "psi") else model.psir).clone()

        res = {}
        vec = np.random.default_rng(1).uniform(size=(61 * 61))
        th = T.make_thermal_model(nx=60, ny=60, N=7, m=40, sigma=-0.1, adjoint_options={"lanczos_guess": True}, rtol=1e-12,
                                  seed=0)
        th.x = np.random.default_rng(0).uniform(0.3, 1.0, th.nnodes)
        seed_th = lambda m: m.add_thermal_compliance_derivative(1.0, vec)       # noqa: E731
        for meth in ("sibk", "pcpg", "pgmres", "laa"):
            th.adjoint_method = meth
            a, pa = grad(th, seed_th, None)
            b, pb = grad(th, seed_th, shard)
            res["thermal_" + meth] = (rel(b, a), rel(pb, pa))
        st = shard.collective_stats()
        res["collectives"] = (st["calls"], st["bytes"])
        for sigma in (3.0, 6.0):                # 6.0: shift inside the spectrum -> indefinite factor, refined solves
            bk = T.make_buckling_model(nx=24, ny=48, N=7, m=40, sigma=sigma, solver_type="IRAM", adjoint_method="sibk",
                                       adjoint_options={"lanczos_guess": True}, rtol=1e-12)
            node = 2 * (bk.nnodes // 2) + 1
            seed_bk = lambda m: m.add_eigenvector_aggregate_derivative(1.0, 100.0, node, mode="tanh")   # noqa: E731
            a, pa = grad(bk, seed_bk, None)
            b, pb = grad(bk, seed_bk, shard)
            res["buckling_sigma%g" % sigma] = (rel(b, a), rel(pb, pa))
        # every rank holds the same gathered result
        t = b.clone()
        dist.broadcast(t, src=0)
        res["ranks_agree"] = (float((t - b).abs().max()), 0.0)
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_sharded_gradient_equals_unsharded_world2():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert sorted(out.keys()) == list(range(world))
    for r in range(world):
        for name, (e_xb, e_psi) in out[r].items():
            if name == "collectives":
                assert e_xb > 0 and e_psi > 0
                continue
            assert e_xb <= TOL and e_psi <= TOL, (r, name, e_xb, e_psi)
