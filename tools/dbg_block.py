"""Developer check: block vs single-vector Lanczos on a buckling model (gradient, orthonormality, T exactness)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from eigd_b200 import device as D, topo as T, arpack
D.init()
nx, ny, N, sigma = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])) if len(sys.argv) > 4 else (24, 48, 7, 3.0)
res = {}
for P in (1, 2):
    arpack.BLOCK_SIZE = P
    bk = T.make_buckling_model(nx=nx, ny=ny, N=N, m=60, sigma=sigma, solver_type="IRAM", adjoint_method="sibk",
                               adjoint_options={"lanczos_guess": True}, rtol=1e-12)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bk.initialize()
    bk.initialize_adjoint()
    node = int(np.argmax(np.abs(bk.prob.full_vector(bk.Qr[:, 0].contiguous()).cpu().numpy())))
    if P == 1:
        node0 = node
    bk.add_eigenvector_aggregate_derivative(1.0, 100.0, node0, mode="tanh")
    bk.finalize_adjoint()
    es = bk.eig_solver
    V = es._V_d
    KV = bk.Kr.spmm(V.contiguous())
    G = D.gemm_tn(V.contiguous(), KV).cpu().numpy()
    OPV = bk.factor.solve_dev(KV)
    Tm = D.gemm_tn(KV, OPV).cpu().numpy()
    Phi = bk.Qr
    KP = bk.Kr.spmm(Phi)
    Gp = D.gemm_tn(Phi, KP).cpu().numpy()
    ar, ao = es.eval_adjoint_residual_norm(bk.Qrb, bk.psir, b_ortho=True)
    print("P=%d m=%d refine=%d info=%s: |V^T K V - I| %.2e  |V^T K OP V - T| %.2e (|T| %.2e)  |Phi^T K Phi - I| %.2e  adj res %.2e  eig solves %d adj solves %d"
          % (P, V.shape[1], bk.factor.refine, bk.factor.info, np.abs(G - np.eye(G.shape[0])).max(), np.abs(Tm - es.T).max(), np.abs(es.T).max(),
             np.abs(Gp - np.eye(N)).max(), float(np.max(ar)), bk.profile["solve preconditioner count"], bk.profile["adjoint preconditioner count"]))
    res[P] = (bk.xb.clone(), bk.psir.clone(), np.asarray(bk.BLF).copy(), Phi.clone())
a, b = res[1], res[2]
rel = lambda x, y: float((x - y).abs().max() / y.abs().max())
sg = torch.sign((a[3] * b[3]).sum(dim=0))
print("xb rel %.2e  psi rel %.2e  BLF rel %.2e  Phi rel %.2e" % (rel(b[0], a[0]), rel(b[1] * sg, a[1]), np.abs(a[2] - b[2]).max() / np.abs(a[2]).max(), rel(b[3] * sg, a[3])))
# ---- details of the last two solvers
import eigd_b200.eigenvector_derivatives as ED
for P in (1, 2):
    arpack.BLOCK_SIZE = P
    bk = T.make_buckling_model(nx=nx, ny=ny, N=N, m=60, sigma=sigma, solver_type="IRAM", adjoint_method="sibk",
                               adjoint_options={"lanczos_guess": True}, rtol=1e-12)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bk.initialize()
    es = bk.eig_solver
    st = es.lanczos_state
    th, idx = es.theta, es.indices
    first, rest = idx[:N], idx[N:]
    gap = np.abs(th[first][None, :] - th[rest][:, None])
    eigs = sigma * th / (th - 1.0)
    print("P=%d ncycles %d nops %d resid max %.2e  min |theta_first - theta_rest| %.3e (|theta| first range %.3e..%.3e)  lam[N-1] %.8f lam[N] %.8f"
          % (P, st.ncycles, st.nops, np.max(st.resid / np.abs(st.theta)), gap.min(), np.abs(th[first]).min(), np.abs(th[first]).max(),
             eigs[idx[N - 1]], eigs[idx[N]]))
    bk.initialize_adjoint()
    bk.add_eigenvector_aggregate_derivative(1.0, 100.0, node0, mode="tanh")
    psi0 = ED._laa_dev(bk.Qrb, es._Bd, es.factor, es.sigma, np.asarray(es.lam), es._V_d, es.Y, es.theta, es.indices, True, es.mode)
    print("   |psi_laa| col norms", (psi0 ** 2).sum(dim=0).sqrt().cpu().numpy()[:6], " |Qrb|", float(bk.Qrb.norm()))
    for lg in (True, False):
        bk.adjoint_options = {"lanczos_guess": lg}
        bk.xb = D.zeros(bk.xb.shape[0]); bk.rhoEb = D.zeros(bk.nelems)
        bk.finalize_adjoint()
        ar, ao = es.eval_adjoint_residual_norm(bk.Qrb, bk.psir, b_ortho=False)
        print("   lanczos_guess=%s: pert.xb %.10f  adj res (b_ortho=False) %.2e  iterations %d  |psi| %.4e" % (
            lg, float((bk.xb * torch.as_tensor(np.random.default_rng(777).uniform(size=bk.xb.shape[0]), device=bk.xb.device)).sum()),
            float(np.max(ar)), bk.profile["adjoint preconditioner count"], float(bk.psir.norm())))
