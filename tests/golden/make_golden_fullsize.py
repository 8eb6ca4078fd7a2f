"""Full-size golden fixtures from the UNMODIFIED reference (BASELINE.json configs[1], [2], [4] at the named sizes).

Run in the build container only (needs /root/reference; minutes of CPU per case):
    python tests/golden/make_golden_fullsize.py c2 c5 c3

The reference runs through oracle/ref_loader.py exactly as the example drivers run it (make_model ->
initialize -> initialize_adjoint -> objective seeds -> finalize_adjoint).  Per case the fixture stores what a
parity test needs and nothing of size O(n * N): the design x, the objective's inputs, the eigenvalues, the
final design gradient xb (one vector), a fixed random direction `pert` and pert . xb, the reference's
operation counts and its own stage timers on this container's host cores (8 cores, scipy SuperLU sequential).

    c2  examples/thermal.py            make_model(nx=500, ny=500, N=10, m=60, sigma=-0.1), x ~ U(0.3, 1) seed 0,
                                       thermal-compliance seeds with vec ~ default_rng(12345)   [the bench.py line]
    c5  examples/natural_frequency.py  make_model(nx=448, ny=224, Lx=2, Ly=1, N=6), x ~ U(0.3, 1) default_rng(b), b = 0,
                                       smooth modal function f = sum_i (phi_i . w_i)^2
    c3  examples/buckling.py           make_model(nx=352, ny=704, N=20, m=60, sigma=3), x = 0.5, eigenvector aggregate
                                       (hb=1, rho=100, tanh) at the dof of largest |phi_1|
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import ref_loader as rl  # noqa: E402

SIBK = {"lanczos_guess": True, "update_guess": False, "bs_target": 1}
RTOL = 1e-12


def _timers(topo):
    p = topo.profile
    keys = ("matrix assembly time", "eigenvalue solve time", "adjoint solution time", "total derivative time",
            "solve preconditioner count", "adjoint preconditioner count", "adjoint iterations")
    return {("ref_" + k.replace(" ", "_")): np.float64(p[k]) for k in keys if k in p}


def c2(nx=500, N=10, m=60, sigma=-0.1):
    th = rl.load_example("thermal")
    t0 = time.perf_counter()
    topo = th.make_model(nx=nx, ny=nx, N=N, m=m, sigma=sigma, solver_type="IRAM", adjoint_method="sibk",
                         adjoint_options=dict(SIBK), rtol=RTOL, deriv_type="tensor")
    t_setup = time.perf_counter() - t0
    x = np.random.default_rng(0).uniform(0.3, 1.0, topo.nnodes)
    vec = np.random.default_rng(12345).uniform(size=topo.nnodes)
    topo.x[:] = x
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        topo.initialize()
        topo.initialize_adjoint()
        topo.add_thermal_compliance_derivative(1.0, vec)
        topo.finalize_adjoint()
    pert = np.random.default_rng(777).uniform(size=topo.x.shape)
    out = dict(nx=nx, ny=nx, N=N, m=m, sigma=sigma, rtol=RTOL, lam=np.asarray(topo.lam).copy(), xb=topo.xb.copy(),
               pert_dot_xb=float(pert @ topo.xb), compliance=float(topo.get_thermal_compliance(vec)),
               ref_setup_s=t_setup, ref_cpu_count=os.cpu_count())
    out.update(_timers(topo))
    return out


def c5(nx=448, ny=224, N=6, design=0):
    nf = rl.load_example("natural_frequency")
    t0 = time.perf_counter()
    topo = nf.make_model(nx=nx, ny=ny, Lx=2.0, Ly=1.0, N=N, solver_type="IRAM", adjoint_method="sibk",
                         adjoint_options=dict(SIBK), rtol=RTOL, deriv_type="tensor")
    t_setup = time.perf_counter() - t0
    x = np.random.default_rng(design).uniform(0.3, 1.0, topo.x.shape)
    topo.x[:] = x
    w = np.random.default_rng(99).normal(size=(topo.nvars, N))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        topo.initialize()
        topo.initialize_adjoint()
        fval = 0.0
        for i in range(topo.N):
            val = topo.Q[:, i] @ w[:, i]
            topo.Qb[:, i] += 2.0 * val * w[:, i]
            fval += val * val
        topo.finalize_adjoint()
    pert = np.random.default_rng(777).uniform(size=topo.x.shape)
    out = dict(nx=nx, ny=ny, N=N, m=topo.eig_solver.m, sigma=topo.sigma, rtol=RTOL, design=design, ndv=len(topo.x),
               lam=np.asarray(topo.lam).copy(), lam_all=np.asarray(topo.eig_solver.lam).copy(), xb=topo.xb.copy(),
               pert_dot_xb=float(pert @ topo.xb), fval=fval, dvmap=np.asarray(topo.fltr.dvmap).astype(np.int32),
               ref_setup_s=t_setup, ref_cpu_count=os.cpu_count())
    out.update(_timers(topo))
    return out


def c3(nx=352, ny=704, N=20, m=60, sigma=3.0, lanczos_guess=True):
    """lanczos_guess=False (fixture fullsize_c3_noguess.npz): at this configuration the shift lies INSIDE the wanted
    spectrum and the reference's IRAM pairs the returned eigenvectors with ``indices[:N]`` = the N smallest Ritz
    values (eigd/eigenvector_derivatives.py:1960-1965), which is not the set ARPACK returned (the N largest |theta|);
    its Lanczos-adjoint guess is then not B-orthogonal to Phi and the gradient it returns (fullsize_c3.npz) disagrees
    with finite differences (-11.90 vs -3.654 along `pert`).  From a zero guess the reference is correct."""
    bk = rl.load_example("buckling")
    t0 = time.perf_counter()
    opts = dict(SIBK)
    opts["lanczos_guess"] = bool(lanczos_guess)
    topo = bk.make_model(nx=nx, ny=ny, N=N, m=m, sigma=sigma, solver_type="IRAM", adjoint_method="sibk",
                         adjoint_options=opts, rtol=RTOL, deriv_type="tensor")
    t_setup = time.perf_counter() - t0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        topo.initialize()
        topo.initialize_adjoint()
        node = int(np.argmax(np.abs(topo.Q[:, 0])))
        h = topo.add_eigenvector_aggregate_derivative(1.0, 100.0, node, mode="tanh")
        topo.finalize_adjoint()
    pert = np.random.default_rng(777).uniform(size=topo.x.shape)
    out = dict(nx=nx, ny=ny, N=N, m=m, sigma=sigma, rtol=RTOL, node=node, aggregate_h=float(h) if h is not None else np.nan,
               BLF=np.asarray(topo.BLF).copy(), lam=np.asarray(topo.lam).copy(), xb=topo.xb.copy(),
               qnode=topo.Q[node, :].copy(), pert_dot_xb=float(pert @ topo.xb), compliance=float(topo.f @ topo.u),
               ref_setup_s=t_setup, ref_cpu_count=os.cpu_count())
    out.update(_timers(topo))
    return out


if __name__ == "__main__":
    if not rl.reference_available():
        raise SystemExit("reference tree not available; fixtures cannot be regenerated here")
    cases = {"c2": c2, "c5": c5, "c3": c3, "c3_noguess": lambda: c3(lanczos_guess=False)}
    for name in sys.argv[1:]:
        t0 = time.perf_counter()
        d = cases[name]()
        path = os.path.join(HERE, "fullsize_%s.npz" % name)
        np.savez_compressed(path, **d)
        print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024), "%.0f s" % (time.perf_counter() - t0),
              {k: (float(v) if np.ndim(v) == 0 else None) for k, v in d.items() if k.startswith("ref_")}, flush=True)
