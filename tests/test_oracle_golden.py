"""The CPU oracle (oracle/eigd_oracle.py, oracle/fe_oracle.py) against the frozen outputs of the
unmodified reference (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import numpy as np
import pytest

import eigd_oracle as eo
import fe_oracle as fo
from conftest import load_golden, corr_from_array, align_signs


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def th():
    return load_golden("thermal_basiclanczos")


def _basic(g, mode="normal"):
    A, B = g["A"], g["B"]
    sigma = float(g["sigma"])
    shifted = (A - sigma * B) if mode == "normal" else (B + sigma * A)
    f = eo.SpLu(shifted)
    s = eo.BasicLanczosOracle(N=int(g["N"]), m=int(g["m_max"]), tol=1e-14, mode=mode)
    s.solve(A, B, f, sigma)
    return s


def test_basic_lanczos_matches_reference(th):
    s = _basic(th)
    assert rel(s.lam0, th["lam"]) < 1e-10
    assert rel(s.alpha, th["alpha"]) < 1e-8
    Phi, _ = align_signs(s.Phi, th["Phi"])
    assert rel(Phi, th["Phi"]) < 1e-8
    assert (s.indices == th["indices"]).all()


@pytest.mark.parametrize("method", ["sibk", "laa", "dl", "pcpg", "pgmres"])
def test_adjoint_methods_match_reference(th, method):
    s = _basic(th)
    # the golden Phi sign convention is the reference's; BasicLanczos is deterministic so equal
    kw = {} if method in ("laa", "dl") else {"rtol": 1e-12}
    psi, data = s.solve_adjoint(th["Phib"].copy(), method=method, **kw)
    tol = 1e-8
    assert rel(psi, th["psi_" + method]) < tol
    ref = corr_from_array(th["corr_" + method])
    assert set(ref) == set(data)


def test_total_derivative_matches_reference(th):
    s = _basic(th)
    conn, X = th["conn"], th["X"]
    mdl = fo.Q4Model(conn, X, "thermal")
    rhoE = th["rhoE"]
    psi = th["psi_sibk"]
    data = corr_from_array(th["corr_sibk"])
    dfdx = np.zeros(mdl.nelems)
    eo.add_total_derivative(th["lam"], th["Phi"], th["lamb"], th["Phib"], psi,
                            lambda w, v: mdl.dK(rhoE, w, v), lambda w, v: mdl.dM(rhoE, w, v), dfdx, data, "normal")
    assert rel(dfdx, th["dfdx_sibk"]) < 1e-10
    flt = fo.ConicFilter(X, float(th["r0"]))
    xb = flt.apply_gradient(mdl.scatter_to_nodes(dfdx))
    assert rel(xb, th["xb_sibk"]) < 1e-10
    # assembly parity (bit-level pattern, values to rounding)
    K, M = mdl.assemble(rhoE)
    assert (K.indptr == th["A_indptr"]).all() and (K.indices == th["A_indices"]).all()
    assert rel(K.data, th["A_data"]) < 1e-13 and rel(M.data, th["B_data"]) < 1e-13


def test_iram_matches_reference():
    g = load_golden("thermal_iram")
    A, B, sigma = g["A"], g["B"], float(g["sigma"])
    f = eo.SpLu(A - sigma * B)
    s = eo.IRAMOracle(N=int(g["N"]), m=int(g["m"]))
    lam, Phi = s.solve(A, B, f, sigma, rng=1)
    scale = np.abs(g["lam"]).max()
    assert np.abs(lam - g["lam"]).max() < 1e-10 * scale
    # the reference's Phi sign is ARPACK's; align before comparing (non-repeated modes here)
    Phi_a, sgn = align_signs(Phi, g["Phi"])
    assert rel(Phi_a[:, 1:], g["Phi"][:, 1:]) < 1e-7
    # adjoint with the reference's own Phi/Phib (sign convention frozen in the fixture)
    s.Phi = g["Phi"].copy()
    s.Y[:, s.indices[:s.N]] *= sgn
    psi, data = s.solve_adjoint(g["Phib"].copy(), method="sibk", rtol=1e-12)
    assert rel(psi, g["psi_sibk"]) < 1e-7


def test_buckling_matches_reference():
    g = load_golden("buckling_basiclanczos")
    s = _basic(g, mode="buckling")
    assert rel(s.lam0, g["lam"]) < 1e-10
    Phi, _ = align_signs(s.Phi, g["Phi"])
    assert rel(Phi, g["Phi"]) < 1e-8
    for method in ("sibk", "pcpg"):
        psi, data = s.solve_adjoint(g["Phib"].copy(), method=method, rtol=1e-12)
        assert rel(psi, g["psi_" + method]) < 1e-7


def test_nf_iram_gradient():
    g = load_golden("nf_iram")
    A, B, sigma = g["A"], g["B"], float(g["sigma"])
    f = eo.SpLu(A - sigma * B)
    Ncomp = int(g["Ncomp"])
    s = eo.IRAMOracle(N=Ncomp, m=int(g["m"]))
    lam, Phi = s.solve(A, B, f, sigma, rng=5)
    scale = np.abs(g["lam_all"]).max()
    assert np.abs(lam - g["lam_all"]).max() < 1e-9 * scale
    # flexible modes are isolated: sign-aligned comparison; rigid-body triple as a subspace
    Pa, _ = align_signs(Phi[:, 3:], g["Phi_all"][:, 3:])
    assert rel(Pa, g["Phi_all"][:, 3:]) < 1e-6
    R0, R1 = Phi[:, :3], g["Phi_all"][:, :3]
    assert np.abs(R0 @ (R0.T @ (B @ R1)) - R1).max() < 1e-6 * np.abs(R1).max()


def test_buckling_fe_oracle_matches_reference_example():
    """oracle/buckling_oracle.py against the frozen outputs of examples/buckling.py: reduced K, fundamental path,
    reduced G(u, x), and the three sensitivity callbacks on seeded operands."""
    import buckling_oracle as bo
    g = load_golden("buckling_basiclanczos")
    E, nu, p, rho0_K, rho0_G = (float(v) for v in g["material"])
    o = bo.BucklingOracle(g["conn"], g["X"], g["reduced"], g["f"], E=E, nu=nu, p=p, rho0_K=rho0_K, rho0_G=rho0_G)
    rhoE = g["rhoE"]
    u, Kr, _ = o.fundamental_path(rhoE)
    Kr.sort_indices()
    assert np.array_equal(Kr.indptr, g["B_indptr"]) and np.array_equal(Kr.indices, g["B_indices"])
    assert rel(Kr.data, g["B_data"]) < 1e-13
    assert rel(u, g["u"]) < 1e-11
    Gr = o.reduce_matrix(o.stress_stiffness(rhoE, g["u"]))
    Gr.sort_indices()
    assert np.array_equal(Gr.indices, g["A_indices"])
    assert rel(Gr.data, g["A_data"]) < 1e-13
    W, V = o.full_vector(g["cb_W"]), o.full_vector(g["cb_V"])
    assert rel(o.dG_du(rhoE, W, V), g["cb_dGdu"]) < 1e-12
    assert rel(o.scatter(o.dG_dx(rhoE, g["u"], W, V)), g["cb_dGdx"]) < 1e-12
    assert rel(o.scatter(o.dK(rhoE, W, V)), g["cb_dKdx"]) < 1e-12
