import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from eigd_b200 import device as D, topo as T
D.init()
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 500
model = T.make_thermal_model(nx=nx, ny=nx, N=10, m=60, sigma=-0.1, adjoint_options={"lanczos_guess": True}, seed=0)
rng = np.random.default_rng(0)
x_d = D.to_device(rng.uniform(0.3, 1.0, model.nnodes)); vec_d = D.to_device(rng.uniform(size=model.nnodes))
def step():
    model.initialize(x=x_d); model.initialize_adjoint(); model.add_thermal_compliance_derivative(1.0, vec_d); model.finalize_adjoint()
for i in range(3):
    t0 = time.perf_counter(); step(); torch.cuda.synchronize(); print("step", i, time.perf_counter() - t0, {k: round(v, 4) for k, v in model.profile.items() if k.endswith("time")})
pr = cProfile.Profile(); pr.enable()
for i in range(3): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
