"""Device-resident counterparts of the reference's example drivers (``TopologyAnalysis`` classes).

They keep the method names, the ``profile`` timer keys and the data flow of

  examples/thermal.py            :268-342 (solve_eigenvalue_problem), :344-372 (initialize /
                                 initialize_adjoint), :428-442 (thermal compliance), :560-623 (finalize_adjoint)
  examples/natural_frequency.py  :317-392, :394-440, :442-519 (three rigid-body modes computed and dropped)

but every array of length n or nelems lives in HBM between the stages: design x -> filter ->
element densities -> K, M values (gather-form assembly) -> shifted matrix -> numeric LDL^T ->
restarted Lanczos -> adjoint right-hand sides -> lock-step sibk -> fused sensitivity kernel ->
node gather -> filter transpose -> xb.  Host code only sequences kernels and does the small
dense algebra the reference also does on the host.
"""
import time

import numpy as np
import torch

from . import device as D
from . import fe
from ._hostdev import is_dev, small_to_dev, to_dev, to_host
from .eigenvector_derivatives import IRAM, BasicLanczos, SpLuOperator


def _now():
    torch.cuda.synchronize()
    return time.perf_counter()


class _Q4Analysis:
    kind = "thermal"
    nrigid = 0            # leading modes computed but dropped (rigid-body modes of the free structure)

    def __init__(self, fltr, conn, X, sigma, N=10, m=None, Ntarget=None, solver_type="IRAM", tol=0.0, rtol=1e-10,
                 eig_atol=1e-5, adjoint_method="sibk", adjoint_options=None, deriv_type="tensor", seed=0, **material):
        self.fltr = fltr
        self.conn, self.X = np.asarray(conn), np.asarray(X)
        self.prob = fe.Q4Problem(self.conn, self.X, self.kind, **material)
        self.nelems, self.nnodes, self.nvars = self.prob.nelems, self.prob.nnodes, self.prob.ndof
        self.sigma, self.N, self.m, self.Ntarget = sigma, N, m, Ntarget
        self.solver_type, self.tol, self.rtol, self.eig_atol = solver_type, tol, rtol, eig_atol
        if adjoint_method == "shift-invert":          # the examples' historical alias (thermal.py:37)
            adjoint_method = "sibk"
        self.adjoint_method = adjoint_method
        self.adjoint_options = dict(adjoint_options or {})
        self.deriv_type = deriv_type
        self.seed = seed
        self.x = 0.95 * np.ones(self.fltr.num_design_vars)      # thermal.py:71, natural_frequency.py:75
        self.Q = self.lam = None
        self.symbolic = None
        self.sharding = None          # dist.ModeSharding for the per-mode adjoint / element-range shards
        self.profile = {"nnodes": self.nnodes, "nelems": self.nelems, "solver_type": solver_type,
                        "adjoint_method": adjoint_method, "N": N}

    # ---- forward ------------------------------------------------------------------------------
    def solve_eigenvalue_problem(self, store=False):
        t0 = _now()
        K, M = self.prob.assemble()
        t1 = _now()
        self.profile["matrix assembly time"] = t1 - t0
        vals = D.axpby(1.0, K.data, -float(self.sigma), M.data)           # K - sigma M on the shared pattern
        shifted = K.with_values(vals)
        coords, dofpn = self.prob.dof_coords()
        if self.symbolic is None:                                         # once per sparsity pattern
            ts = _now()
            sym = D.Symbolic(self.prob.indptr, self.prob.indices, self.nvars, coords=coords, dof_per_node=dofpn)
            self.symbolic = (sym, sym.assembly_map_device(K.indptr, K.indices))
            self.profile["symbolic analysis time"] = _now() - ts
            t1 = _now()
        self.factor = SpLuOperator(shifted, symbolic=self.symbolic)
        self.K, self.M = K, M
        self.factor.count = 0
        ncomp = self.N + self.nrigid
        if self.solver_type == "IRAM":
            if self.m is None:
                self.m = max(2 * ncomp + 1, 60)
            self.eig_solver = IRAM(N=ncomp, m=self.m, eig_atol=self.eig_atol, tol=self.tol)
            self.eig_solver.seed = self.seed
        else:
            if self.m is None:
                self.m = max(3 * ncomp + 1, 60)
            self.eig_solver = BasicLanczos(N=ncomp, m=self.m, eig_atol=self.eig_atol, tol=self.tol, Ntarget=self.Ntarget)
        self.eig_solver.sharding = self.sharding
        self.prob.sharding = self.sharding
        lam, _ = self.eig_solver.solve(K, M, self.factor, self.sigma)
        t2 = _now()
        self.profile["solve preconditioner count"] = self.factor.count
        self.profile["eigenvalue solve time"] = t2 - t1
        self.profile["m"] = self.m
        self.profile["eig_solver.m"] = str(self.eig_solver.m)
        self.lam0 = np.asarray(lam)
        self.Q0 = self.eig_solver._Phi_d                                   # (n, N + nrigid) device
        return self.lam0[self.nrigid:], self.Q0[:, self.nrigid:]

    def initialize(self, store=False, x=None):
        if x is not None:
            self.x = x
        self.x_d = to_dev(self.x)
        self.rho = self.fltr.apply(self.x_d)
        self.rhoE = self.prob.set_density(rho=self.rho)
        self.lam, self.Q = self.solve_eigenvalue_problem(store)
        self.N = len(self.lam)
        return

    def initialize_adjoint(self):
        self.xb = D.zeros(self.x_d.shape[0])
        self.rhoEb = D.zeros(self.nelems)
        self.lamb = np.zeros(self.N)
        self.Q0b = D.zeros(self.nvars, self.N + self.nrigid)
        self.Qb = self.Q0b[:, self.nrigid:]

    # ---- reverse ------------------------------------------------------------------------------
    def finalize_adjoint(self):
        res_list = []
        self.factor.count = 0
        t0 = _now()
        psi0, corr_data = self.eig_solver.solve_adjoint(self.Q0b, rtol=self.rtol, method=self.adjoint_method,
                                                        callback=res_list.append, **self.adjoint_options)
        t1 = _now()
        self.psi0 = psi0
        self.psi = psi0[:, self.nrigid:]
        self.profile["adjoint preconditioner count"] = self.factor.count
        self.profile["adjoint solution time"] = t1 - t0
        self.profile["adjoint residuals"] = res_list
        self.profile["adjoint iterations"] = len(res_list)
        self.profile["adjoint correction data"] = corr_data
        nr = self.nrigid
        if nr:                                                            # natural_frequency.py:484-494
            data0 = {}
            for i, items in corr_data.items():
                keep = [(j, xi, eta) for (j, xi, eta) in items if j >= nr]
                if i >= nr and keep:
                    data0[i] = keep
            corr_data = data0
        lamb0 = np.zeros(self.N + nr)
        lamb0[nr:] = self.lamb
        self.eig_solver.add_total_derivative(lamb0, self.Q0b, psi0, self.prob.dAdx, self.prob.dBdx, self.rhoEb,
                                             adj_corr_data=corr_data, deriv_type=self.deriv_type)
        rhob = self.prob.scatter_to_nodes(self.rhoEb)
        g = self.fltr.apply_gradient(rhob, self.x_d)
        D.axpby(1.0, self.xb, 1.0, g, out=self.xb)
        t2 = _now()
        self.profile["total derivative time"] = t2 - t1
        return

    def time_to_gradient(self):
        p = self.profile
        return p["eigenvalue solve time"] + p["adjoint solution time"] + p["total derivative time"]


class ThermalTopologyAnalysis(_Q4Analysis):
    """examples/thermal.py ``ThermalTopologyAnalysis`` (scalar heat conduction, one zero mode)."""

    kind = "thermal"

    def __init__(self, fltr, conn, X, kappa=1.0, density=1.0, heat_capacity=1.0, p=3, beta=1e-6, sigma=-0.1, N=10,
                 **kw):
        super().__init__(fltr, conn, X, sigma, N=N, kappa=kappa, density=density, heat_capacity=heat_capacity,
                         p=float(p), beta=beta, **kw)

    def get_thermal_compliance(self, vec):
        """sum_{i>=1} (phi_i . vec)^2 / lam_i   (thermal.py:428-434)"""
        c = to_host(D.gemm_tn(self.Q, to_dev(vec))).ravel()
        return float(np.sum(c[1:] ** 2 / self.lam[1:]))

    def add_thermal_compliance_derivative(self, compb, vec):
        """thermal.py:436-442: Qb_i += 2 compb (phi_i . vec) vec / lam_i ; lamb_i -= compb (phi_i . vec)^2 / lam_i^2"""
        vec_d = to_dev(vec)
        c = to_host(D.gemm_tn(self.Q, vec_d)).ravel()
        coef = 2.0 * compb * c / self.lam
        coef[0] = 0.0
        D.gemm_nn(vec_d.unsqueeze(1), small_to_dev(coef[None, :]), self.Qb, alpha=1.0, beta=1.0)
        dl = compb * (c * c) / self.lam ** 2
        dl[0] = 0.0
        self.lamb -= dl


class NaturalFrequencyAnalysis(_Q4Analysis):
    """examples/natural_frequency.py ``TopologyAnalysis``: free plane-stress structure, N + 3 modes
    computed with sigma < 0 and the three rigid-body modes dropped (:348, :383-384, :457-458)."""

    kind = "plane_stress"
    nrigid = 3

    def __init__(self, fltr, conn, X, E=1.0, nu=0.3, density=1.0, p=3, rho0_K=1e-6, ptype_K="simp", q=5.0,
                 sigma=-10.0, N=6, **kw):
        super().__init__(fltr, conn, X, sigma, N=N, E=E, nu=nu, density=density, p=float(p), rho0_K=rho0_K,
                         ptype_K=ptype_K, q=q, **kw)

    def add_modal_function_derivative(self, w):
        """Seed of a smooth function f = sum_i (phi_i . w_i)^2 of the flexible modes (used by the tests and
        the bench; the KS minimum-frequency objective of the example is host-side numpy)."""
        w_d = to_dev(w)
        val = to_host(D.col_dot(self.Q, w_d))
        D.col_axpy(self.Qb, small_to_dev(2.0 * val), w_d, sign=1.0)
        return float(np.sum(val**2))


def make_thermal_model(nx=128, ny=128, Lx=1.0, Ly=1.0, rfact=4.0, **kwargs):
    """examples/thermal.py ``make_model`` (:1475-1509)."""
    conn, X = fe.grid_mesh(nx, ny, Lx, Ly)
    fltr = fe.NodeFilter(conn, X, r0=rfact * (Ly / ny))
    return ThermalTopologyAnalysis(fltr, conn, X, **kwargs)


def make_natural_frequency_model(nx=128, ny=64, Lx=2.0, Ly=1.0, rfact=4.0, **kwargs):
    """examples/natural_frequency.py ``make_model`` (:850-990) without symmetry map / point masses."""
    conn, X = fe.grid_mesh(nx, ny, Lx, Ly)
    fltr = fe.NodeFilter(conn, X, r0=rfact * (Ly / ny))
    return NaturalFrequencyAnalysis(fltr, conn, X, **kwargs)
