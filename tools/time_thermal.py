"""Developer timing of the thermal time-to-gradient path (not the bench contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import eigd_b200 as E
from eigd_b200 import device as D, fe

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 500
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10
m = 60
sigma = -0.1
def sync(): torch.cuda.synchronize()
def T():
    sync(); return time.perf_counter()
D.init()
t0 = T()
conn, X = fe.grid_mesh(nx, nx, 1.0, 1.0)
prob = fe.Q4Problem(conn, X, "thermal")
t1 = T(); print("problem setup %.3f s" % (t1 - t0))
flt = fe.NodeFilter(conn, X, r0=4.0 / nx)
t2 = T(); print("filter setup %.3f s" % (t2 - t1))
rng = np.random.default_rng(0)
x = rng.uniform(0.3, 1.0, prob.nnodes)
vec = rng.uniform(size=prob.nnodes)
for rep in range(3):
    l0 = D.launch_count()
    ta = T()
    x_d = D.to_device(x)
    rho = flt.apply(x_d)
    prob.set_density(rho=rho)
    K, M = prob.assemble()
    tb = T()
    vals = D.axpby(1.0, K.data, -sigma, M.data)
    f = E.SpLuOperator(K.with_values(vals), coords=X, dof_per_node=1)
    tc = T()
    s = E.IRAM(N=N, m=m); s.seed = 0
    lam, Phi = s.solve(K, M, f, sigma)
    td = T()
    nsolve_eig = f.count; f.count = 0
    Phi_d = s._Phi_d
    vec_d = D.to_device(vec)
    c = D.gemm_tn(Phi_d, vec_d).cpu().numpy().ravel()
    coef = 2.0 * c / lam; coef[0] = 0.0
    lamb = -(c * c) / lam**2; lamb[0] = 0.0
    Phib = D.zeros(prob.ndof, N)
    D.gemm_nn(vec_d.unsqueeze(1), D.to_device(coef[None, :]), Phib)
    te = T()
    psi, data = s.solve_adjoint(Phib, method="sibk", rtol=1e-10, lanczos_guess=True)
    tf = T()
    dfdx = D.zeros(prob.nelems)
    s.add_total_derivative(lamb, Phib, psi, prob.dAdx, prob.dBdx, dfdx, adj_corr_data=data, deriv_type="tensor")
    xb = flt.apply_gradient(prob.scatter_to_nodes(dfdx))
    xb_h = xb.cpu().numpy()
    tg = T()
    print("rep %d: assemble %.4f factor %.4f (info %s) eig %.4f (%d solves, %d cycles) seeds %.4f adjoint %.4f (%d solves, its %s) dfdx %.4f | ttg %.4f  launches %d" % (
        rep, tb - ta, tc - tb, f.info, td - tc, nsolve_eig, s.lanczos_state.ncycles, te - td, tf - te, f.count, max(s.adjoint_info), tg - tf, tg - tb, D.launch_count() - l0))
print("lam", lam)
res, orth = s.eval_adjoint_residual_norm(Phib, psi, b_ortho=True)
print("adjoint res", res.max(), "ortho", orth.max())
# solve accuracy
b = torch.as_tensor(rng.normal(size=(prob.ndof, 10)), device="cuda")
xx = f.solve_dev(b)
r = K.with_values(vals).spmm(xx) - b
print("solve rel resid", float(r.abs().max() / b.abs().max()))
for k in (1, 10, 20):
    b = torch.as_tensor(rng.normal(size=(prob.ndof, k)), device="cuda").contiguous()
    f.solve_dev(b); ta = T()
    for _ in range(10): f.solve_dev(b)
    tb = T(); print("solve k=%d: %.3f ms" % (k, (tb - ta) * 100))
ta = T()
for _ in range(5): f.lu.numeric(vals, f._amap)
tb = T(); print("numeric factor: %.3f ms" % ((tb - ta) * 200))
