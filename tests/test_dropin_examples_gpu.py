"""The drop-in boundary, executed: the reference's UNMODIFIED example drivers (examples/thermal.py,
natural_frequency.py, buckling.py, loaded from the offline install baseline/_ref) run with ``import eigd``
resolving to this repository's alias package (eigd/ -> eigd_b200), i.e. their own host-side assembly, their own
numpy dA/dx and dB/dx callbacks, scipy matrices in and numpy arrays out, ``SpLuOperator(mat)`` without any extra
keyword -- and must reproduce the frozen outputs of the same drivers run against the reference itself
(tests/golden/*.npz, written by tests/golden/make_golden.py).

The bodies below mirror make_golden.py line for line and import nothing from eigd_b200.

Tolerances (BASELINE.json north_star): eigenvalues 1e-10, sign-aligned eigenvectors / adjoints and gradients 1e-8."""
import os
import sys
import warnings

import numpy as np
import pytest

from conftest import ROOT, load_golden, align_signs

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

SIBK = {"lanczos_guess": True, "update_guess": False, "bs_target": 1}


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def rl():
    if not os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "examples", "thermal.py")):
        pytest.skip("baseline/_ref (offline install of the reference, baseline/install_reference.py) is not present")
    os.environ["EIGD_REFERENCE_ROOT"] = os.path.join(ROOT, "baseline", "_ref")
    import ref_loader
    ref_loader.REF_ROOT = os.environ["EIGD_REFERENCE_ROOT"]
    return ref_loader


def test_alias_surface_matches_reference_signatures(rl):
    """Every public name of the reference package exists in the alias with the same parameters (the alias may
    add optional keyword arguments after them)."""
    import inspect
    import eigd as alias
    assert "eigd_b200" in alias.IRAM.__module__
    ref = rl.load_reference()
    try:
        names = [n for n in ("SpLuOperator", "IRAM", "BasicLanczos", "add_eig_total_derivative", "eval_adjoint_residual_norm",
                             "are_eigenvalues_repeated", "generate_adjoint_correction", "laa", "dl", "pcpg", "pgmres", "sibk")]
        for n in names:
            r, a = getattr(ref, n), getattr(alias, n)
            rp = list(inspect.signature(r.__init__ if inspect.isclass(r) else r).parameters.values())
            ap = list(inspect.signature(a.__init__ if inspect.isclass(a) else a).parameters.values())
            assert [p.name for p in ap[: len(rp)]] == [p.name for p in rp], n
            for p, q in zip(rp, ap):
                if p.default is not inspect._empty and not isinstance(p.default, (dict,)):
                    assert p.default == q.default, (n, p.name)
            assert all(q.default is not inspect._empty for q in ap[len(rp):]), n
        from eigd.arpack import eigsh_mod  # noqa: F401
    finally:
        import importlib
        for k in ("eigd", "eigd.arpack", "eigd.eigenvector_derivatives"):
            sys.modules.pop(k, None)
        importlib.import_module("eigd")


def test_thermal_example_unmodified(rl):
    """examples/thermal.py ThermalTopologyAnalysis (:14) through make_model (:1475), IRAM + sibk -- the flow of
    make_golden.thermal_case("IRAM", ["sibk"])."""
    g = load_golden("thermal_iram")
    th = rl.load_example("thermal", against="alias")
    assert "eigd_b200" in th.IRAM.__module__ and "eigd_b200" in th.SpLuOperator.__module__
    np.random.seed(3)
    topo = th.make_model(nx=int(g["nx"]), ny=int(g["ny"]), Lx=1.0, Ly=0.8, N=int(g["N"]), m=int(g["m"]), solver_type="IRAM",
                         adjoint_method="sibk", adjoint_options=dict(SIBK), rtol=1e-12, tol=0.0)
    topo.x[:] = np.random.uniform(0.3, 1.0, topo.x.shape)
    assert np.array_equal(topo.x, g["x"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        topo.initialize()
        topo.initialize_adjoint()
        vec = np.random.uniform(size=topo.nnodes)
        topo.add_thermal_compliance_derivative(1.0, vec)
        topo.finalize_adjoint()
    assert isinstance(topo.Q, np.ndarray) and isinstance(topo.psi, np.ndarray) and isinstance(topo.xb, np.ndarray)
    assert rel(topo.lam, g["lam"]) < 1e-10
    Qa, s = align_signs(topo.Q, g["Phi"])
    assert rel(Qa[:, 1:], g["Phi"][:, 1:]) < 1e-8
    assert rel(topo.psi * s, g["psi_sibk"]) < 1e-8
    assert rel(topo.rhoEb, g["dfdx_sibk"]) < 1e-8
    assert rel(topo.xb, g["xb_sibk"]) < 1e-8
    assert topo.profile["solve preconditioner count"] > 0 and topo.profile["adjoint preconditioner count"] > 0
    res, _ = topo.eig_solver.eval_adjoint_residual_norm(topo.Qb, topo.psi, b_ortho=False)
    assert np.all(res < 1e-9 * max(np.linalg.norm(topo.Qb, axis=0).max(), 1.0))


def test_natural_frequency_example_unmodified(rl):
    """examples/natural_frequency.py TopologyAnalysis (IRAM asks for N + 3 modes, three rigid-body modes dropped,
    in-place sign flips of the returned eigenvectors at :383-390) -- the flow of make_golden.nf_case()."""
    g = load_golden("nf_iram")
    nf = rl.load_example("natural_frequency", against="alias")
    np.random.seed(0)
    topo = nf.make_model(nx=int(g["nx"]), ny=int(g["ny"]), Lx=2.0, Ly=1.0, N=int(g["N"]), solver_type="IRAM",
                         adjoint_method="sibk", adjoint_options=dict(SIBK), rtol=1e-12, deriv_type="tensor")
    topo.x[:] = np.random.uniform(0.3, 1.0, topo.x.shape)
    nf.MinFreqOpt(topo, ks_param=1.0, fixed_mass=1.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        topo.initialize()
        topo.initialize_adjoint()
        w = np.random.uniform(size=topo.Q.shape)
        topo.Qb[:] = w * 0.0
        for i in range(topo.N):
            val = topo.Q[:, i] @ w[:, i]
            topo.Qb[:, i] += 2.0 * val * w[:, i]
            topo.lamb[i] += 0.3 * (i + 1)
        topo.finalize_adjoint()
    assert np.array_equal(w, g["w"])
    assert np.abs(topo.lam - g["lam"]).max() < 1e-10 * np.abs(g["lam"]).max()
    _, s = align_signs(topo.Q, g["Phi"])
    assert rel(topo.Q * s, g["Phi"]) < 1e-8
    assert rel(topo.psi * s, g["psi"]) < 1e-8
    assert rel(topo.rhoEb, g["dfdx"]) < 1e-8
    assert rel(topo.xb, g["xb"]) < 1e-8


def test_buckling_example_unmodified(rl):
    """examples/buckling.py TopologyAnalysis (mode="buckling", BasicLanczos, fundamental-path solve with the
    example's own scipy splu) -- the flow of make_golden.buckling_case()."""
    g = load_golden("buckling_basiclanczos")
    bk = rl.load_example("buckling", against="alias")
    for method in ("sibk", "pcpg"):
        opts = dict(SIBK) if method == "sibk" else {"lanczos_guess": True}
        np.random.seed(0)
        topo = bk.make_model(nx=int(g["nx"]), ny=int(g["ny"]), N=int(g["N"]), m=60, sigma=3.0, solver_type="BasicLanczos",
                             adjoint_method=method, adjoint_options=opts, rtol=1e-12, deriv_type="tensor")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            topo.initialize()
            topo.initialize_adjoint()
            node = int(g["node"])
            topo.add_eigenvector_aggregate_derivative(1.0, 100.0, node, mode="tanh")
            topo.finalize_adjoint()
        assert rel(topo.lam, g["lam"]) < 1e-10
        _, s = align_signs(topo.Qr, g["Phi"])
        assert rel(topo.Qr * s, g["Phi"]) < 1e-8
        assert rel(topo.psir * s, g["psi_" + method]) < 1e-8, method
        assert rel(topo.rhob, g["rhob_" + method]) < 1e-8, method
        assert rel(topo.xb, g["xb_" + method]) < 1e-8, method


def test_transient_thermal_ks_example_unmodified(rl):
    """examples/thermal.py ThermalOpt (:997-1321: modal transient heat equations, KS aggregate of the mean temperatures
    and its adjoint) over ThermalTopologyAnalysis built by make_opt_model (:1512), both unmodified, on the alias -- the
    flow of make_golden.transient_flow, shared with the golden generator."""
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "thermal_transient.npz")))
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    th = rl.load_example("thermal", against="alias")
    assert "eigd_b200" in th.IRAM.__module__
    out = make_golden.transient_flow(th, nx=int(g["nx"]), N=int(g["N"]), m=int(g["m"]), nsteps=int(g["nsteps"]),
                                     tfinal=float(g["tfinal"]), ks_rho=float(g["ks_rho"]), seed=int(g["seed"]))
    assert np.array_equal(out["x"], g["x"])
    assert rel(out["lam"], g["lam"]) < 1e-10
    assert abs(out["ks"] - float(g["ks"])) < 1e-8 * abs(float(g["ks"]))
    assert rel(out["lamb"], g["lamb"]) < 1e-8
    assert rel(out["xb"], g["xb"]) < 1e-8
