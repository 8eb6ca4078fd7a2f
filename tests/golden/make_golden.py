"""Regenerates the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference is loaded through oracle/ref_loader.py (scipy>=1.15 compatibility shim that
leaves eigd/eigenvector_derivatives.py untouched).  Each fixture freezes the inputs (CSR A, B,
shift, adjoint right-hand sides) and the reference's outputs (eigenpairs, adjoints, gradients)
of one small seeded case.  Krylov bases of the ARPACK path are not reproducible (random start
vector), so for IRAM only converged quantities are stored; the seeded BasicLanczos path also
stores its basis and tridiagonal coefficients.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import ref_loader as rl  # noqa: E402

SIBK = {"lanczos_guess": True, "update_guess": False, "bs_target": 1}


def csr_dict(prefix, A):
    A = A.tocsr()
    A.sort_indices()
    return {prefix + "_indptr": A.indptr.astype(np.int32), prefix + "_indices": A.indices.astype(np.int32),
            prefix + "_data": A.data.copy(), prefix + "_shape": np.array(A.shape)}


def corr_to_array(data):
    rows = [(i, j, xi, eta) for i, items in sorted(data.items()) for (j, xi, eta) in items]
    return np.array(rows, dtype=float).reshape(-1, 4)


def thermal_case(solver_type, methods, nx=24, ny=20, N=6, m=30, seed=3):
    th = rl.load_example("thermal")
    out = {}
    for method in methods:
        opts = dict(SIBK) if method == "sibk" else ({} if method == "laa" else {"lanczos_guess": method != "dl"})
        np.random.seed(seed)
        topo = th.make_model(nx=nx, ny=ny, Lx=1.0, Ly=0.8, N=N, m=m, solver_type=solver_type, adjoint_method=method,
                             adjoint_options=opts, rtol=1e-12, tol=1e-14 if solver_type != "IRAM" else 0.0)
        topo.x[:] = np.random.uniform(0.3, 1.0, topo.x.shape)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            topo.initialize()
            topo.initialize_adjoint()
            vec = np.random.uniform(size=topo.nnodes)
            topo.add_thermal_compliance_derivative(1.0, vec)
            topo.finalize_adjoint()
        if not out:
            out.update(csr_dict("A", topo.K))
            out.update(csr_dict("B", topo.M))
            out.update(dict(sigma=topo.sigma, N=N, m=topo.eig_solver.m, x=topo.x.copy(), rhoE=topo.rhoE.copy(), vec=vec,
                            conn=topo.conn, X=topo.X, r0=topo.fltr.r0, lam=topo.lam.copy(), Phi=topo.Q.copy(),
                            Phib=topo.Qb.copy(), lamb=topo.lamb.copy(), nx=nx, ny=ny))
            es = topo.eig_solver
            if solver_type != "IRAM":
                out.update(dict(m_max=es.m_max, alpha=es.alpha.copy(), beta=es.beta.copy(), V=es.V[:, :es.m].copy(), theta=es.theta.copy(),
                                indices=es.indices.copy(), eig_res=es.eig_res.copy()))
        out["psi_" + method] = topo.psi.copy()
        out["corr_" + method] = corr_to_array(topo.profile["adjoint correction data"])
        out["dfdx_" + method] = topo.rhoEb.copy()
        out["xb_" + method] = topo.xb.copy()
        res, orth = topo.eig_solver.eval_adjoint_residual_norm(topo.Qb, topo.psi, b_ortho=False)
        out["res_" + method] = res
        out["nsolves_" + method] = topo.profile["adjoint preconditioner count"]
    if solver_type != "IRAM":
        # the coupled variants of the reference's free function sibk (block Arnoldi, residual recycling), zero guess
        es = topo.eig_solver
        for tag, kw in (("bs2", dict(bs_target=2)), ("ug", dict(update_guess=True)), ("bs3ug", dict(bs_target=3, update_guess=True))):
            psi_v, _, info_v = rl.load_reference().sibk(topo.Qb, topo.K, topo.M, topo.lam, topo.Q, sigma=topo.sigma,
                                            factor=topo.factor, rtol=1e-12, **kw)
            out["psi_sibk_free_" + tag] = psi_v.copy()
            out["info_sibk_free_" + tag] = np.array(info_v)
        # complex-step run of the example (thermal.py:652-661): design x + i h pert -> complex K, M -> the reference's
        # BasicLanczos treats the imaginary parts as forward derivatives; stored as tangents (imag / h)
        hcs = 1e-30
        x0 = topo.x.copy()
        pert = np.random.default_rng(21).uniform(size=x0.shape)
        topo.x = x0.astype(complex)
        topo.x.imag += hcs * pert
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            topo.initialize()
        Kc, Mc = topo.K.tocsr(), topo.M.tocsr()
        Kc.sort_indices()
        Mc.sort_indices()
        out.update(dict(cs_pert=pert, cs_A_tan=Kc.data.imag / hcs, cs_B_tan=Mc.data.imag / hcs,
                        cs_A_real=Kc.data.real.copy(), cs_lam=np.asarray(topo.lam).real.copy(),
                        cs_lam_tan=np.asarray(topo.lam).imag / hcs, cs_Phi=topo.Q.real.copy(), cs_Phi_tan=topo.Q.imag / hcs,
                        cs_m=topo.eig_solver.m))
        topo.x = x0
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            topo.initialize()
        # BasicLanczos with selective orthogonalisation (eigd/eigenvector_derivatives.py:1553-1605) on the same pencil
        sel = rl.load_reference().BasicLanczos(N=N, m=es.m_max, tol=1e-14, ortho_type="selective")
        fsel = rl.load_reference().SpLuOperator((topo.K - topo.sigma * topo.M).tocsc())
        lam_s, Phi_s = sel.solve(topo.K, topo.M, fsel, topo.sigma)
        out.update(dict(sel_lam=lam_s.copy(), sel_Phi=Phi_s.copy(), sel_m=sel.m, sel_alpha=sel.alpha.copy(),
                        sel_beta=sel.beta.copy(), sel_eig_res=sel.eig_res.copy()))
    return out


def nf_case(nx=48, ny=24, N=6, seed=0):
    nf = rl.load_example("natural_frequency")
    np.random.seed(seed)
    topo = nf.make_model(nx=nx, ny=ny, Lx=2.0, Ly=1.0, N=N, solver_type="IRAM", adjoint_method="sibk",
                         adjoint_options=dict(SIBK), rtol=1e-12, deriv_type="tensor")
    topo.x[:] = np.random.uniform(0.3, 1.0, topo.x.shape)
    opt = nf.MinFreqOpt(topo, ks_param=1.0, fixed_mass=1.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        topo.initialize()
        topo.initialize_adjoint()
        # seed of the adjoint: derivative of a smooth function of the flexible modes
        w = np.random.uniform(size=topo.Q.shape)
        topo.Qb[:] = w * 0.0
        lam = topo.lam
        for i in range(topo.N):
            val = topo.Q[:, i] @ w[:, i]
            topo.Qb[:, i] += 2.0 * val * w[:, i]
            topo.lamb[i] += 0.3 * (i + 1)
        topo.finalize_adjoint()
    out = {}
    out.update(csr_dict("A", topo.K))
    out.update(csr_dict("B", topo.M))
    es = topo.eig_solver
    out.update(dict(sigma=topo.sigma, N=topo.N, Ncomp=len(es.lam), m=es.m, x=topo.x.copy(), rhoE=topo.rhoE.copy(),
                    conn=topo.conn, X=topo.X, lam_all=es.lam.copy(), Phi_all=es.Phi.copy(), lam=topo.lam.copy(),
                    Phi=topo.Q.copy(), Phib=topo.Qb.copy(), lamb=topo.lamb.copy(), psi=topo.psi.copy(),
                    dfdx=topo.rhoEb.copy(), xb=topo.xb.copy(), nx=nx, ny=ny, w=w,
                    corr=corr_to_array(topo.profile["adjoint correction data"])))
    return out


def buckling_case(solver_type="BasicLanczos", methods=("sibk", "pcpg"), nx=16, ny=32, N=5, seed=0):
    bk = rl.load_example("buckling")
    out = {}
    for method in methods:
        opts = dict(SIBK) if method == "sibk" else {"lanczos_guess": True}
        np.random.seed(seed)
        topo = bk.make_model(nx=nx, ny=ny, N=N, m=60, sigma=3.0, solver_type=solver_type, adjoint_method=method,
                             adjoint_options=opts, rtol=1e-12, deriv_type="tensor")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            topo.initialize()
            topo.initialize_adjoint()
            node = int(np.argmax(np.abs(topo.Q[:, 0])))
            topo.add_eigenvector_aggregate_derivative(1.0, 100.0, node, mode="tanh")
            topo.finalize_adjoint()
        if not out:
            out.update(csr_dict("A", topo.Gr))
            out.update(csr_dict("B", topo.Kr))
            es = topo.eig_solver
            out.update(dict(sigma=topo.sigma, N=N, m=es.m, m_max=es.m_max, lam=topo.lam.copy(), Phi=topo.Qr.copy(), Phib=topo.Qrb.copy(),
                            lamb=topo.lamb.copy(), reduced=np.array(topo.reduced), u=topo.u.copy(), rhoE=topo.rhoE.copy(),
                            conn=topo.conn, X=topo.X, node=node, nx=nx, ny=ny,
                            f=topo.f.copy(), material=np.array([topo.E, topo.nu, topo.p, topo.rho0_K, topo.rho0_G]),
                            BLF=np.asarray(topo.BLF).copy()))
            # element callbacks of the example on seeded operands (isolates the sensitivity kernels from the solvers)
            rng = np.random.default_rng(7)
            Wr, Vr = rng.normal(size=topo.Qr.shape), rng.normal(size=topo.Qr.shape)
            Wf, Vf = topo.full_vector(Wr), topo.full_vector(Vr)
            dfds = topo.intital_stress_stiffness_matrix_deriv(topo.rhoE, topo.Te, topo.detJ, Wf, Vf)
            out.update(dict(cb_W=Wr, cb_V=Vr,
                            cb_dGdu=topo.get_stress_stiffness_matrix_uderiv_tensor(dfds, topo.Be),
                            cb_dGdx=topo.get_stress_stiffness_matrix_xderiv_tensor(topo.rhoE, topo.u, dfds, topo.Be),
                            cb_dKdx=topo.get_stiffness_matrix_deriv(topo.rhoE, Wf, Vf)))
        out["psi_" + method] = topo.psir.copy()
        out["corr_" + method] = corr_to_array(topo.profile["adjoint correction data"])
        out["xb_" + method] = topo.xb.copy()
        out["rhob_" + method] = topo.rhob.copy()          # nodal gradient before the filter transpose
    return out


def objectives_case(seed=1):
    """The objective-side pieces of the examples, run with the unmodified reference drivers:
      * MinFreqOpt of examples/natural_frequency.py (:693-803): KS minimum frequency with point masses, full gradient;
      * eval_ks_buckling / eval_ks_buckling_derivative of examples/buckling.py (:641-700);
      * the Helmholtz (PDE) filter of examples/node_filter.py (:90-162, 164-217) with a design-variable map."""
    out = {}
    nf = rl.load_example("natural_frequency")
    np.random.seed(seed)
    nx, ny, N = 40, 20, 5
    topo = nf.make_model(nx=nx, ny=ny, Lx=2.0, Ly=1.0, N=N, solver_type="IRAM", adjoint_method="sibk",
                         adjoint_options=dict(SIBK), rtol=1e-12, deriv_type="tensor")
    topo.x[:] = np.random.uniform(0.3, 1.0, topo.x.shape)
    opt = nf.MinFreqOpt(topo, ks_param=30.0, fixed_mass=10.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        opt.initialize()
        opt.initialize_adjoint()
        opt.finalize_adjoint()
    out.update(dict(mf_nx=nx, mf_ny=ny, mf_N=N, mf_x=topo.x.copy(), mf_ks=opt.get_min_frequency(), mf_omega=opt.omega.copy(),
                    mf_omegab=opt.omegab.copy(), mf_xb=topo.xb.copy(), mf_lamb=topo.lamb.copy()))
    bk = rl.load_example("buckling")
    np.random.seed(seed)
    bnx, bny, bN = 16, 32, 5
    topo = bk.make_model(nx=bnx, ny=bny, N=bN, m=60, sigma=3.0, solver_type="IRAM", adjoint_method="sibk",
                         adjoint_options=dict(SIBK), rtol=1e-12, deriv_type="tensor")
    topo.x[:] = np.random.uniform(0.3, 1.0, topo.x.shape)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        topo.initialize()
        ks = topo.eval_ks_buckling(ks_rho=160.0)
        dks = topo.eval_ks_buckling_derivative(ks_rho=160.0)
    out.update(dict(ks_nx=bnx, ks_ny=bny, ks_N=bN, ks_x=topo.x.copy(), ks_value=ks, ks_grad=np.asarray(dks).copy(),
                    ks_BLF=np.asarray(topo.BLF).copy()))
    nfl = rl.load_example("node_filter")
    conn, X = topo.conn, topo.X
    dvmap, ndv = topo.fltr.dvmap, topo.fltr.num_design_vars
    hf = nfl.NodeFilter(conn, X, r0=0.12, ftype="helmholtz", dvmap=dvmap, num_design_vars=ndv, projection=True, beta=6.0, eta=0.4)
    xh = np.random.default_rng(seed).uniform(0.2, 1.0, ndv)
    gh = np.random.default_rng(seed + 1).normal(size=X.shape[0])
    out.update(dict(hf_conn=conn, hf_X=X, hf_dvmap=np.asarray(dvmap), hf_ndv=ndv, hf_r0=0.12, hf_x=xh, hf_g=gh,
                    hf_rho=hf.apply(xh.copy()), hf_grad=hf.apply_gradient(gh, xh.copy())))
    hf0 = nfl.NodeFilter(conn, X, r0=0.12, ftype="helmholtz")
    x0 = np.random.default_rng(seed + 2).uniform(0.2, 1.0, X.shape[0])
    out.update(dict(hf0_x=x0, hf0_rho=hf0.apply(x0.copy()), hf0_grad=hf0.apply_gradient(gh, x0.copy())))
    return out


def shell_case(nx=16, ny=14, ncx=4, ncy=3, N=6, m=30, omega0=10.0, seed=5):
    """The flow of examples/crm.py (:212-370) -- IRAM, sibk, modal compliance with f[1::6] = 1, add_eig_total_derivative in
    its default per-mode "vector" form -- with the UNMODIFIED reference solvers on host matrices of the synthetic shell
    (oracle/shell_oracle.py stands in for TACS)."""
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    import shell_oracle as so
    from eigd_b200.shell import cylindrical_panel, shell_unit_matrices       # host-only numpy functions
    ref = rl.load_reference()
    conn, X, P = cylindrical_panel(nx, ny, 1.0, 0.9, 2.0)
    E1, E3, F1, F3 = shell_unit_matrices(conn, X)
    ei, ej = np.arange(nx * ny) % nx, np.arange(nx * ny) // nx
    comp = (ei * ncx // nx) + ncx * (ej * ncy // ny)
    nodes = np.arange((nx + 1) * (ny + 1)).reshape(nx + 1, ny + 1)
    orc = so.ShellOracle(conn, X, comp, nodes[:, 0], E1, E3, F1, F3)
    x = np.random.default_rng(seed).uniform(0.6, 1.4, orc.ncomp)
    Kr, Mr = orc.assemble(x)
    sigma = omega0 ** 2
    factor = ref.SpLuOperator((Kr - sigma * Mr).tocsc())
    es = ref.IRAM(N=N, m=m, eig_atol=1e-5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lam, Q = es.solve(Kr, Mr, factor, sigma)
    f = np.zeros(orc.ndof_full)
    f[1::6] = 1.0
    fr = f[orc.reduced]
    Qb, lamb = np.zeros(Q.shape), np.zeros(N)
    comp_val = 0.0
    for i in range(N):                                   # crm.py:267-293
        val = Q[:, i].dot(fr)
        comp_val += val * val / lam[i]
        Qb[:, i] += 2.0 * val * fr / lam[i]
        lamb[i] -= (val * val) / lam[i] ** 2
    psi, corr = es.solve_adjoint(Qb, rtol=1e-12, method="sibk", **SIBK)
    grad = np.zeros(orc.ncomp)
    grad = ref.add_eig_total_derivative(lam, Q, lamb, Qb, psi, lambda w, v: orc.dK(x, w, v), lambda w, v: orc.dM(x, w, v), grad,
                                        adj_corr_data=corr)
    out = {}
    out.update(csr_dict("A", Kr))
    out.update(csr_dict("B", Mr))
    out.update(dict(nx=nx, ny=ny, ncx=ncx, ncy=ncy, N=N, m=es.m, omega0=omega0, sigma=sigma, x=x, lam=lam.copy(), Phi=Q.copy(),
                    Phib=Qb.copy(), lamb=lamb.copy(), psi=psi.copy(), grad=grad.copy(), compliance=comp_val,
                    corr=corr_to_array(corr), comp=comp, conn=conn, X=X))
    return out


def transient_heat_funcs(tfinal):
    """The heat loads of the example's own transient run (reference examples/thermal.py:1665-1690)."""
    beta = 50 / tfinal
    H = lambda t: 0.5 + 0.5 * np.tanh(beta * t)                                   # noqa: E731
    interval = lambda t, t0, t1: (H(t - t0) + H(t1 - t) - 1.0)                    # noqa: E731
    interval0 = lambda t, t0, t1: interval(t, t0, t1) - interval(0, t0, t1)      # noqa: E731
    f = {"center": lambda t: 10 * interval0(t, 0.1 * tfinal, 1.5 * tfinal)}
    for k in range(4):
        f["corner%d" % k] = lambda t: -2.5 * interval0(t, 0.1 * tfinal, 1.5 * tfinal)
    return {"test": f}


def transient_flow(th, nx=32, N=8, m=40, nsteps=100, tfinal=25.0, ks_rho=10.0, seed=2):
    """Transient thermal KS objective and its gradient (reference examples/thermal.py: ThermalOpt :997-1321 over
    ThermalTopologyAnalysis built by make_opt_model :1512).  Shared by the golden generator (th = the example loaded
    against the reference) and by tests/test_dropin_examples_gpu.py (th = the same file loaded against the alias)."""
    element_sets = {"center": [], "corner0": [], "corner1": [], "corner2": [], "corner3": []}
    topo = th.make_opt_model(nx=nx, rfact=4.0, N=N, m=m, p=3, epsilon=1e-5, solver_type="IRAM", adjoint_method="sibk",
                             adjoint_options=dict(SIBK), element_sets=element_sets, eig_atol=1e-5, rtol=1e-12,
                             deriv_type="tensor")
    np.random.seed(seed)
    topo.x[:] = np.random.uniform(0.3, 1.0, topo.x.shape)
    opt = th.ThermalOpt(topo, transient_heat_funcs(tfinal), nsteps=nsteps, tfinal=tfinal)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        opt.initialize()
        ks = opt.eval_ks_functions(ks_rho)
        opt.initialize_adjoint()
        opt.add_ks_derivative(ks_rho, {"test": 1.0})
        opt.finalize_adjoint()
    return {"nx": nx, "N": N, "m": m, "nsteps": nsteps, "tfinal": tfinal, "ks_rho": ks_rho, "seed": seed, "x": topo.x.copy(),
            "lam": np.asarray(topo.lam).copy(), "ks": float(np.real(ks["test"])), "xi": np.asarray(opt.xi["test"]).copy(),
            "lamb": np.asarray(topo.lamb).copy(), "xb": np.asarray(topo.xb).copy()}


def transient_case():
    return transient_flow(rl.load_example("thermal"))


if __name__ == "__main__":
    if not rl.reference_available():
        raise SystemExit("reference tree not available; fixtures cannot be regenerated here")
    cases = {
        "thermal_basiclanczos": lambda: thermal_case("BasicLanczos", ["sibk", "laa", "dl", "pcpg", "pgmres"]),
        "thermal_iram": lambda: thermal_case("IRAM", ["sibk"]),
        "nf_iram": nf_case,
        "buckling_basiclanczos": buckling_case,
        "shell_iram": shell_case,
        "objectives": objectives_case,
        "thermal_transient": transient_case,
    }
    only = set(sys.argv[1:])
    for name, fn in cases.items():
        if only and name not in only:
            continue
        d = fn()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024))
