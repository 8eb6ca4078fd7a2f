"""torchrun check: the per-mode sharded adjoint / element-range sharded df/dx reproduce the unsharded gradient."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from eigd_b200 import device as D, topo as T
from eigd_b200.dist import ModeSharding
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
D.init("cuda:%d" % local)
dist.init_process_group("nccl")
rank = dist.get_rank()
shard = ModeSharding()
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())

def grad(model, seed_fn, sh):
    model.sharding = sh
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.initialize()
    model.initialize_adjoint()
    seed_fn(model)
    model.finalize_adjoint()
    return model.xb.clone(), (model.psi if hasattr(model, "psi") else model.psir).clone()

vec = np.random.default_rng(1).uniform(size=(61 * 61))
th = T.make_thermal_model(nx=60, ny=60, N=7, m=40, sigma=-0.1, adjoint_options={"lanczos_guess": True}, rtol=1e-12, seed=0)
th.x = np.random.default_rng(0).uniform(0.3, 1.0, th.nnodes)
a, pa = grad(th, lambda m: m.add_thermal_compliance_derivative(1.0, vec), None)
b, pb = grad(th, lambda m: m.add_thermal_compliance_derivative(1.0, vec), shard)
if rank == 0: print("thermal  xb rel %.2e  psi rel %.2e" % (rel(b, a), rel(pb, pa)))
bk = T.make_buckling_model(nx=24, ny=48, N=7, m=40, sigma=3.0, solver_type="IRAM", adjoint_method="sibk",
                           adjoint_options={"lanczos_guess": True}, rtol=1e-12)
node = 2 * (bk.nnodes // 2) + 1
a, pa = grad(bk, lambda m: m.add_eigenvector_aggregate_derivative(1.0, 100.0, node, mode="tanh"), None)
b, pb = grad(bk, lambda m: m.add_eigenvector_aggregate_derivative(1.0, 100.0, node, mode="tanh"), shard)
if rank == 0: print("buckling xb rel %.2e  psi rel %.2e  BLF %s" % (rel(b, a), rel(pb, pa), bk.BLF[:3]))
bk2 = T.make_buckling_model(nx=24, ny=48, N=7, m=40, sigma=6.0, solver_type="IRAM", adjoint_method="sibk",
                            adjoint_options={"lanczos_guess": True}, rtol=1e-12)
a, pa = grad(bk2, lambda m: m.add_eigenvector_aggregate_derivative(1.0, 100.0, node, mode="tanh"), None)
ra = bk2.eig_solver.eval_adjoint_residual_norm(bk2.Qrb, bk2.psir, b_ortho=True)[0].max()
b, pb = grad(bk2, lambda m: m.add_eigenvector_aggregate_derivative(1.0, 100.0, node, mode="tanh"), shard)
rb = bk2.eig_solver.eval_adjoint_residual_norm(bk2.Qrb, bk2.psir, b_ortho=True)[0].max()
if rank == 0: print("buckling sigma=6 (indefinite, refine=%d, info %s) xb rel %.2e  psi rel %.2e  adjoint res %.2e / %.2e  BLF %s" % (bk2.factor.refine, bk2.factor.info, rel(b, a), rel(pb, pa), ra, rb, bk2.BLF))
for meth in ("pcpg", "pgmres", "laa"):
    th.adjoint_method = meth
    a, pa = grad(th, lambda m: m.add_thermal_compliance_derivative(1.0, vec), None)
    b, pb = grad(th, lambda m: m.add_thermal_compliance_derivative(1.0, vec), shard)
    if rank == 0: print("thermal %s xb rel %.2e  psi rel %.2e" % (meth, rel(b, a), rel(pb, pa)))
dist.destroy_process_group()
