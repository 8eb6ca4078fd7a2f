"""Developer check: wall time of consecutive numpy-API (end-to-end) gradient steps, one number per step."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import eigd_b200 as E
from eigd_b200 import device as D, topo as T
D.init()
N, SIGMA = 10, -0.1
model = T.make_thermal_model(nx=500, ny=500, N=N, m=60, sigma=SIGMA, solver_type="IRAM", adjoint_method="sibk",
                             adjoint_options={"lanczos_guess": True}, rtol=1e-10, deriv_type="tensor", seed=0)
x_d = D.to_device(np.random.default_rng(0).uniform(0.3, 1.0, model.nnodes))
vec_h = np.random.default_rng(12345).uniform(size=model.nnodes)
model.initialize(x=x_d)
K_h, M_h = model.K.to_scipy(), model.M.to_scipy()
mat_h = (K_h - SIGMA * M_h).tocsc()
for A_h in (K_h, M_h, mat_h):
    vals = D.pinned_empty(A_h.data.shape); vals[...] = A_h.data; A_h.data = vals
prob = model.prob
Phib_h = D.pinned_empty((model.nnodes, N))
Phib_t, vec_t = torch.from_numpy(Phib_h), torch.from_numpy(vec_h)


def step():
    t = [time.perf_counter()]
    f = E.SpLuOperator(mat_h, coords=model.X, dof_per_node=1); t.append(time.perf_counter())
    s = E.IRAM(N=N, m=60); s.seed = 0
    lam, Phi = s.solve(K_h, M_h, f, SIGMA); t.append(time.perf_counter())
    c = vec_h @ Phi
    coef = 2.0 * c / lam; coef[0] = 0.0
    torch.outer(vec_t, torch.from_numpy(coef), out=Phib_t)
    lamb = -(c * c) / lam**2; lamb[0] = 0.0; t.append(time.perf_counter())
    psi, data = s.solve_adjoint(Phib_h, method="sibk", rtol=1e-10, lanczos_guess=True); t.append(time.perf_counter())
    dfdx = np.zeros(prob.nelems)
    s.add_total_derivative(lamb, Phib_h, psi, prob.dAdx, prob.dBdx, dfdx, adj_corr_data=data, deriv_type="tensor")
    t.append(time.perf_counter())
    return [1e3 * (b - a) for a, b in zip(t, t[1:])]

for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 16):
    p = step()
    print("step %2d: %6.1f ms   SpLu %5.1f  solve %5.1f  seeds %4.1f  adjoint %5.1f  dfdx %4.1f" % (i, sum(p), *p), flush=True)
