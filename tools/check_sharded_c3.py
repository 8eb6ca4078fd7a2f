"""torchrun check at C3-like settings (N = 20, indefinite shifted matrix, refinement on): sharded vs unsharded."""
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
nx = int(os.environ.get("NX", "120"))
bk = T.make_buckling_model(nx=nx, ny=2 * nx, N=20, m=60, sigma=float(os.environ.get("SIGMA", "5.0")), solver_type="IRAM",
                           adjoint_method="sibk", adjoint_options={"lanczos_guess": True}, rtol=1e-10)
node = 2 * (bk.nnodes // 2) + 1
out = []
for sh in (None, shard, None, shard):
    bk.sharding = sh
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bk.initialize()
    bk.initialize_adjoint()
    h = bk.add_eigenvector_aggregate_derivative(1.0, 100.0, node, mode="tanh")
    bk.finalize_adjoint()
    res = bk.eig_solver.eval_adjoint_residual_norm(bk.Qrb, bk.psir, b_ortho=True)[0]
    out.append((bk.xb.clone(), res.max()))
    if rank == 0:
        print("sharded" if sh else "single ", "xb norm %.10e  h %.6e  adjoint res max %.2e  refine %d info %s iters %s" % (
            bk.xb.norm().item(), h, res.max(), bk.factor.refine, bk.factor.info, bk.profile["adjoint iterations"]))
if rank == 0:
    r = lambda a, b: float((a - b).abs().max() / b.abs().max())
    print("single vs single %.2e   sharded vs single %.2e   sharded vs sharded %.2e" % (r(out[2][0], out[0][0]), r(out[1][0], out[0][0]), r(out[3][0], out[1][0])))
dist.destroy_process_group()
