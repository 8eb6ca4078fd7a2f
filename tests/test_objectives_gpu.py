"""The objective-side pieces of the reference examples on the device drivers, against the frozen
output of the unmodified reference (tests/golden/objectives.npz, written by make_golden.objectives_case):

  * MinFreqOpt (examples/natural_frequency.py:693-803): KS minimum frequency with point masses at the node sets -> ks value
    1e-10, design gradient 1e-8;
  * eval_ks_buckling / eval_ks_buckling_derivative (examples/buckling.py:641-700) -> value 1e-10, gradient 1e-8;
  * Helmholtz filter (examples/node_filter.py:90-162, 164-217), with and without the design-variable map and the tanh
    projection -> filtered field and gradient 1e-10."""
import warnings

import numpy as np
import pytest

from conftest import load_golden  # noqa: F401
import os
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def g():
    from eigd_b200 import device as D
    D.init()
    return dict(np.load(os.path.join(GOLDEN, "objectives.npz"), allow_pickle=False))


def test_min_frequency_objective_vs_reference(g):
    from eigd_b200 import topo as T
    nx, ny, N = int(g["mf_nx"]), int(g["mf_ny"]), int(g["mf_N"])
    model = T.make_natural_frequency_model(nx=nx, ny=ny, Lx=2.0, Ly=1.0, N=N, solver_type="IRAM", adjoint_method="sibk",
                                           adjoint_options={"lanczos_guess": True}, rtol=1e-12, deriv_type="tensor")
    assert set(model.node_sets) == {"node[%d,%d]" % (i, j) for i in range(3) for j in range(3)}
    model.x = np.array(g["mf_x"])
    opt = T.MinFreqOpt(model, ks_param=30.0, fixed_mass=10.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        opt.initialize()
        opt.initialize_adjoint()
        opt.finalize_adjoint()
    assert rel(opt.omega, g["mf_omega"]) < 1e-9
    assert abs(opt.get_min_frequency() - float(g["mf_ks"])) < 1e-9 * abs(float(g["mf_ks"]))
    assert rel(opt.omegab, g["mf_omegab"]) < 1e-7
    assert rel(model.lamb, g["mf_lamb"]) < 1e-7
    assert rel(model.xb.cpu().numpy(), g["mf_xb"]) < 1e-8


def test_ks_buckling_objective_vs_reference(g):
    from eigd_b200 import topo as T
    model = T.make_buckling_model(nx=int(g["ks_nx"]), ny=int(g["ks_ny"]), N=int(g["ks_N"]), m=60, sigma=3.0, solver_type="IRAM",
                                  adjoint_method="sibk", adjoint_options={"lanczos_guess": True}, rtol=1e-12, deriv_type="tensor")
    model.x = np.array(g["ks_x"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.initialize()
    assert rel(model.BLF, g["ks_BLF"]) < 1e-10
    assert abs(model.eval_ks_buckling(160.0) - float(g["ks_value"])) < 1e-10 * abs(float(g["ks_value"]))
    grad = model.eval_ks_buckling_derivative(160.0)
    assert isinstance(grad, np.ndarray) and rel(grad, g["ks_grad"]) < 1e-8


def test_helmholtz_filter_vs_reference(g):
    from eigd_b200 import fe
    f = fe.NodeFilter(g["hf_conn"], g["hf_X"], r0=float(g["hf_r0"]), ftype="helmholtz", dvmap=g["hf_dvmap"],
                      num_design_vars=int(g["hf_ndv"]), projection=True, beta=6.0, eta=0.4)
    assert rel(f.apply(np.array(g["hf_x"])), g["hf_rho"]) < 1e-10
    assert rel(f.apply_gradient(np.array(g["hf_g"]), np.array(g["hf_x"])), g["hf_grad"]) < 1e-10
    f0 = fe.NodeFilter(g["hf_conn"], g["hf_X"], r0=float(g["hf_r0"]), ftype="helmholtz")
    assert rel(f0.apply(np.array(g["hf0_x"])), g["hf0_rho"]) < 1e-10
    assert rel(f0.apply_gradient(np.array(g["hf_g"]), np.array(g["hf0_x"])), g["hf0_grad"]) < 1e-10
    with pytest.raises(ValueError):
        fe.NodeFilter(g["hf_conn"], g["hf_X"], r0=0.1, ftype="nope")


def test_complex_step_design_through_the_device_driver():
    """examples/thermal.py:652-661: the design is perturbed by i h p and the eigenpairs' imaginary parts are read as
    directional derivatives.  The device driver carries the tangent through filter, material law, assembly, LDL^T and the
    Lanczos recurrence as dual numbers (exact to first order, any h); compared with the reference's own complex run
    (h = 1e-30) frozen in tests/golden/thermal_basiclanczos.npz."""
    from conftest import load_golden
    from eigd_b200 import device as D, topo as T
    D.init()
    g = load_golden("thermal_basiclanczos")
    model = T.make_thermal_model(nx=int(g["nx"]), ny=int(g["ny"]), Lx=1.0, Ly=0.8, N=int(g["N"]), m=int(g["m_max"]),
                                 solver_type="BasicLanczos", tol=1e-14, adjoint_method="sibk", rtol=1e-12)
    h = 1e-6
    model.x = np.array(g["x"]).astype(complex)
    model.x.imag += h * g["cs_pert"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.initialize()
    lam, Q = np.asarray(model.lam), np.asarray(model.Q)
    assert np.iscomplexobj(lam) and np.iscomplexobj(Q)
    scale = np.abs(g["cs_lam"]).max()
    assert np.abs(lam.real - g["cs_lam"]).max() < 1e-10 * scale
    assert np.abs(lam.imag / h - g["cs_lam_tan"]).max() < 1e-8 * np.abs(g["cs_lam_tan"]).max()
    sgn = np.sign(np.einsum("ij,ij->j", Q.real, g["cs_Phi"]))
    assert rel(Q.real * sgn, g["cs_Phi"]) < 1e-8
    assert rel(Q.imag / h * sgn, g["cs_Phi_tan"]) < 1e-6
    # the objective of the example in complex arithmetic: its imaginary part is the directional derivative
    vec = np.random.default_rng(2).uniform(size=model.nnodes)
    c = model.get_thermal_compliance(vec)
    model.x = np.array(g["x"]).real.copy()
    model.adjoint_options = {"lanczos_guess": True}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.initialize()
        model.initialize_adjoint()
        model.add_thermal_compliance_derivative(1.0, vec)
        model.finalize_adjoint()
    ans = float(g["cs_pert"] @ model.xb.cpu().numpy())
    assert abs(c.imag / h - ans) <= 1e-7 * abs(ans), (c.imag / h, ans)
    with pytest.raises(NotImplementedError):
        m2 = T.make_thermal_model(nx=8, ny=8, N=3, solver_type="IRAM")
        m2.x = m2.x.astype(complex)
        m2.initialize()
