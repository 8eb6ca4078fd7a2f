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
from ._hostdev import is_dev, like_input, small_to_dev, to_dev, to_host
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
        if not is_dev(self.x) and np.iscomplexobj(self.x):
            return self._initialize_dual(store)
        self.x_d = to_dev(self.x)
        self.rho = self.fltr.apply(self.x_d)
        self.rhoE = self.prob.set_density(rho=self.rho)
        self.lam, self.Q = self.solve_eigenvalue_problem(store)
        self.N = len(self.lam)
        return

    def _initialize_dual(self, store=False):
        """Complex design x + i h p (the complex-step check of the examples, thermal.py:652-661): the imaginary part is
        carried as a first-order tangent through the real fp64 kernels -- filter, material law, assembly, LDL^T, Lanczos
        (eigd_b200/dual.py) -- and ``lam`` / ``Q`` come back as complex host arrays (value + i tangent), which is what
        the reference's complex run returns.  As in the reference, only ``BasicLanczos`` supports it."""
        if self.solver_type == "IRAM":
            raise NotImplementedError("complex-step designs need solver_type='BasicLanczos' (as in the reference: ARPACK is real)")
        x = np.asarray(self.x)
        xr, xt = np.ascontiguousarray(x.real), np.ascontiguousarray(x.imag)
        self.x_d = to_dev(xr)
        self.rho = self.fltr.apply(self.x_d)
        rho_t = self.fltr.apply_tangent(self.x_d, to_dev(xt))
        self.rhoE = self.prob.set_density(rho=self.rho)
        K, M = self.prob.assemble()
        Kt, Mt = self.prob.assemble_tangent(rho_t)
        coords, dofpn = self.prob.dof_coords()
        if self.symbolic is None:
            sym = D.Symbolic(self.prob.indptr, self.prob.indices, self.nvars, coords=coords, dof_per_node=dofpn)
            self.symbolic = (sym, sym.assembly_map_device(K.indptr, K.indices))
        shifted = K.with_values(D.axpby(1.0, K.data, -float(self.sigma), M.data))
        self.factor = SpLuOperator(shifted, symbolic=self.symbolic)
        self.factor.tangent = K.with_values(D.axpby(1.0, Kt.data, -float(self.sigma), Mt.data))
        self.factor.dtype = np.dtype(np.complex128)
        self.K, self.M = K, M
        ncomp = self.N + self.nrigid
        if self.m is None:
            self.m = max(3 * ncomp + 1, 60)
        self.eig_solver = BasicLanczos(N=ncomp, m=self.m, eig_atol=self.eig_atol, tol=self.tol, Ntarget=self.Ntarget)
        lam, Phi = self.eig_solver.solve((K, Kt), (M, Mt), self.factor, self.sigma)
        self.lam0, self.Q0 = np.asarray(lam), np.asarray(Phi)               # complex host arrays
        self.lam, self.Q = self.lam0[self.nrigid:], self.Q0[:, self.nrigid:]
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
        if not is_dev(self.Q):                      # dual-number mode: complex host eigenpairs, N-sized host algebra
            c = np.asarray(self.Q).T @ (to_host(vec) if is_dev(vec) else np.asarray(vec))
            return np.sum(c[1:] ** 2 / self.lam[1:])
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

    # ---- objective helpers of examples/natural_frequency.py:521-563 -------------------------------------
    def get_frequencies(self):
        return np.sqrt(np.asarray(self.lam))

    def add_frequency_derivatives(self, omegab):
        """lamb_i += omegab_i / (2 sqrt(lam_i))   (:524-529)"""
        self.lamb += 0.5 * np.asarray(omegab) / np.sqrt(np.asarray(self.lam))

    def _set_rows(self, name):
        if name not in self.node_sets:
            raise ValueError("Unrecognized point name")
        nodes = np.asarray(self.node_sets[name], dtype=np.int64)
        key = ("rows", name)
        if key not in self.__dict__.setdefault("_cache", {}):
            self._cache[key] = (torch.as_tensor(2 * nodes, device=D.dev()), torch.as_tensor(2 * nodes + 1, device=D.dev()))
        return nodes, self._cache[key]

    def get_point_coefficients(self, name):
        """Mean position of a node set and the mean modal displacement over it, (3, N) (:531-551)."""
        nodes, (rx, ry) = self._set_rows(name)
        weight = 1.0 / len(nodes)
        x0 = np.zeros(3)
        x0[0], x0[1] = weight * np.sum(self.X[nodes, 0]), weight * np.sum(self.X[nodes, 1])
        xcoef = None
        if self.Q is not None:
            xcoef = np.zeros((3, self.N))
            xcoef[0] = weight * to_host(self.Q.index_select(0, rx).sum(dim=0))
            xcoef[1] = weight * to_host(self.Q.index_select(0, ry).sum(dim=0))
        return x0, xcoef

    def add_point_derivative(self, name, x0b, xcoefb):
        """Qb[2 nodes, i] += w xcoefb[0, i], Qb[2 nodes + 1, i] += w xcoefb[1, i]   (:553-563)"""
        if xcoefb is None:
            return
        nodes, (rx, ry) = self._set_rows(name)
        weight = 1.0 / len(nodes)
        self.Qb.index_add_(0, rx, small_to_dev(weight * np.asarray(xcoefb)[0][None, :]).expand(len(nodes), self.N).contiguous())
        self.Qb.index_add_(0, ry, small_to_dev(weight * np.asarray(xcoefb)[1][None, :]).expand(len(nodes), self.N).contiguous())

    def add_modal_function_derivative(self, w):
        """Seed of a smooth function f = sum_i (phi_i . w_i)^2 of the flexible modes (used by the tests and
        the bench; the KS minimum-frequency objective of the example is host-side numpy)."""
        w_d = to_dev(w)
        val = to_host(D.col_dot(self.Q, w_d))
        D.col_axpy(self.Qb, small_to_dev(2.0 * val), w_d, sign=1.0)
        return float(np.sum(val**2))


class MinFreqOpt:
    """KS-aggregated minimum natural frequency of the structure with a point mass attached at each node set in turn
    (examples/natural_frequency.py ``MinFreqOpt`` :693-803).  For node set s with modal coefficients c (3 x N) the
    frequencies of structure + mass are those of the N x N pencil (diag(omega^2), I + m c^T c); the objective is the
    soft minimum (KS, parameter ks_param) over all sets and modes.  Only N x N host algebra: the n-sized work is the
    eigensolve and the adjoint of ``topo``."""

    def __init__(self, topo, ks_param=1.0, fixed_mass=1.0):
        self.topo, self.ks_param, self.fixed_mass = topo, ks_param, fixed_mass
        self.ks_min = 0.0
        self.node_sets = topo.node_sets
        self.coef, self.coefb, self.omega, self.omegab = {}, {}, None, None

    def initialize(self, store=False):
        self.topo.initialize(store)
        self.omega = self.topo.get_frequencies()
        self.coef = {name: self.topo.get_point_coefficients(name)[1] for name in self.node_sets}
        self.ks_min, self.omegab, self.coefb = self._eval_min_frequency(self.omega, self.coef, self.ks_param, self.fixed_mass)

    def initialize_adjoint(self):
        self.topo.initialize_adjoint()

    def finalize_adjoint(self):
        self.topo.add_frequency_derivatives(self.omegab)
        for name in self.node_sets:
            self.topo.add_point_derivative(name, None, self.coefb[name])
        self.topo.finalize_adjoint()

    def get_min_frequency(self):
        return self.ks_min

    @staticmethod
    def _soft_min(vals, p):
        lo = np.min(vals)
        e = np.exp(-p * (vals - lo))
        return lo - np.log(np.sum(e)) / p, e / np.sum(e)

    def _eval_min_frequency(self, omega, xcoef, ks_param=30.0, fixed_mass=1.0):
        from scipy.linalg import eigh
        N = len(omega)
        names = list(xcoef)
        pencil = {}
        inner = {}
        for name in names:
            c0 = xcoef[name]
            lam0, Q0 = eigh(np.diag(omega ** 2), np.eye(N) + fixed_mass * (c0.T @ c0))
            om0 = np.sqrt(lam0)
            pencil[name] = (om0, Q0)
            inner[name] = self._soft_min(om0, ks_param)                      # (KS of this set, weights over its modes)
        # outer soft minimum over the sets, anchored at min(omega, KS values) as the reference does (:757-768)
        vals = np.array([inner[name][0] for name in names])
        anchor = min(np.min(omega), np.min(vals)) if len(vals) else np.min(omega)
        e = np.exp(-ks_param * (vals - anchor))
        ks = anchor - np.log(np.sum(e)) / ks_param
        wset = e / np.sum(e)
        omegab = np.zeros(N)
        xcoefb = {}
        for name, ws in zip(names, wset):
            c0 = xcoef[name]
            om0, Q0 = pencil[name]
            wb = 0.5 * inner[name][1] * ws / om0                             # d ks / d lam0_i
            omegab += 2.0 * omega * np.einsum("ik,k,ik->i", Q0, wb, Q0)      # through K0 = diag(omega^2)
            # through M0 = I + m c^T c:  d lam0_i = -lam0_i q_i^T dM0 q_i
            sc = 2.0 * wb * fixed_mass * om0 ** 2
            xcoefb[name] = -(c0 @ Q0) * sc[None, :] @ Q0.T
        return ks, omegab, xcoefb


def make_thermal_model(nx=128, ny=128, Lx=1.0, Ly=1.0, rfact=4.0, **kwargs):
    """examples/thermal.py ``make_model`` (:1475-1509)."""
    conn, X = fe.grid_mesh(nx, ny, Lx, Ly)
    fltr = fe.NodeFilter(conn, X, r0=rfact * (Ly / ny))
    return ThermalTopologyAnalysis(fltr, conn, X, **kwargs)


def natural_frequency_design_map(nx, ny, rfact=4.0, Mx=3, My=3, ns=2):
    """Design-variable map, node sets and element sets of examples/natural_frequency.py ``make_model`` (:896-954):
    Mx x My patches of non-design material (dvmap = -1, the filter treats them as x = 1) and a four-fold mirror
    symmetry of the remaining nodes.  Same visiting order as the reference, so the design-variable numbering is
    bit-identical (checked against the reference's own ``fltr.dvmap`` in tests/test_fullsize_golden_gpu.py)."""
    nodes = np.arange((nx + 1) * (ny + 1), dtype=np.int64).reshape(nx + 1, ny + 1)
    dvmap = np.zeros((nx + 1, ny + 1), dtype=np.int64)
    node_sets, element_sets = {}, {}
    ns = max(int(ns * ny // 32), int(rfact // 2))
    sx, sy = nx // (Mx - 1), ny // (My - 1)
    for i in range(Mx):
        for j in range(My):
            if i < Mx // 2:
                imin, imax = max(0, sx * i - ns + 1), min(nx, sx * i + ns + 1)
            else:
                lo, hi = max(0, sx * (Mx - i - 1) - ns + 1), min(nx, sx * (Mx - i - 1) + ns + 1)
                imin, imax = max(0, nx - hi), min(nx, nx - lo)
            if j < My // 2:
                jmin, jmax = max(0, sy * j - ns), min(ny, sy * j + ns)
            else:
                lo, hi = max(0, sy * (My - j - 1) - ns), min(ny, sy * (My - j - 1) + ns)
                jmin, jmax = max(0, ny - hi), min(ny, ny - lo)
            ii, jj = np.meshgrid(np.arange(imin, imax), np.arange(jmin, jmax), indexing="ij")
            name = "node[%d,%d]" % (i, j)
            node_sets[name] = nodes[ii, jj].ravel()
            element_sets[name] = (ii + nx * jj).ravel()
            dvmap[imin:imax, jmin:jmax] = -1
    index = 0
    for i in range(nx // 2 + 1):
        js = np.nonzero(dvmap[i, : ny // 2 + 1] >= 0)[0]         # ascending j, as the reference's inner loop (:945-951)
        ids = index + np.arange(len(js))
        for a in (i, nx - i):
            dvmap[a, js] = ids
            dvmap[a, ny - js] = ids
        index += len(js)
    return dvmap.ravel(), index, node_sets, element_sets


def make_natural_frequency_model(nx=128, ny=64, Lx=1.0, Ly=1.0, rfact=4.0, N=10, Mx=3, My=3, ns=2, symmetry=True,
                                 projection=False, b0=None, **kwargs):
    """examples/natural_frequency.py ``make_model`` (:850-976), including its symmetric design-variable map and the
    non-design patches (``symmetry=False``: one design variable per node)."""
    conn, X = fe.grid_mesh(nx, ny, Lx, Ly)
    if symmetry:
        dvmap, ndv, node_sets, element_sets = natural_frequency_design_map(nx, ny, rfact, Mx, My, ns)
        fltr = fe.NodeFilter(conn, X, r0=rfact * (Ly / ny), dvmap=dvmap, num_design_vars=ndv, projection=bool(projection),
                             beta=b0)
    else:
        node_sets, element_sets = {}, {}
        fltr = fe.NodeFilter(conn, X, r0=rfact * (Ly / ny), projection=bool(projection), beta=b0)
    model = NaturalFrequencyAnalysis(fltr, conn, X, N=N, **kwargs)
    model.node_sets, model.element_sets = node_sets, element_sets
    return model


# ------------------------------------------------------------------------------------------
# linearised buckling -- examples/buckling.py
# ------------------------------------------------------------------------------------------
class BucklingTopologyAnalysis:
    """examples/buckling.py ``TopologyAnalysis`` (:19-118): (K + BLF * G(u)) phi = 0 on the reduced dofs, with
    the fundamental path K_r u_r = f_r, shift-invert about ``sigma`` (mode="buckling": A = G_r, B = K_r,
    factor of K_r + sigma G_r, :548-640), and the gradient of an eigenvector functional including the
    fundamental-path adjoint (:896-982).  Every vector of length n, 2*nnodes or nelems stays in HBM."""

    def __init__(self, fltr, conn, X, bcs, forces={}, E=1.0, nu=0.3, ptype_K="simp", ptype_M="simp", ptype_G="simp",
                 rho0_K=1e-6, rho0_M=1e-9, rho0_G=1e-9, p=3.0, q=5.0, density=1.0, sigma=3.0, N=10, m=None,
                 solver_type="IRAM", tol=0.0, rtol=1e-10, eig_atol=1e-5, adjoint_method="shift-invert",
                 adjoint_options=None, cost=1, deriv_type="tensor", seed=0):
        self.fltr = fltr
        self.seed = seed                  # IRAM start vector (deterministic: the eigensolve is replicated when sharded)
        self.conn, self.X = np.asarray(conn), np.asarray(X)
        self.prob = fe.BucklingQ4Problem(self.conn, self.X, bcs, forces, E=E, nu=nu, density=density, p=float(p),
                                         q=float(q), rho0_K=rho0_K, rho0_G=rho0_G, ptype_K=ptype_K.lower(),
                                         ptype_G=ptype_G.lower())
        self.nelems, self.nnodes, self.nvars = self.prob.nelems, self.prob.nnodes, self.prob.ndof
        self.reduced, self.f = self.prob.reduced, self.prob.f
        self.E, self.nu, self.p, self.q, self.density = E, nu, p, q, density
        self.rho0_K, self.rho0_G = rho0_K, rho0_G
        self.sigma, self.N, self.m = sigma, N, m
        self.solver_type, self.tol, self.rtol, self.eig_atol = solver_type, tol, rtol, eig_atol
        if adjoint_method == "shift-invert":
            adjoint_method = "sibk"
        self.adjoint_method = adjoint_method
        self.adjoint_options = dict(adjoint_options or {})
        self.cost, self.deriv_type = cost, deriv_type
        self.x = 0.5 * np.ones(self.fltr.num_design_vars)                 # buckling.py:81
        self.symbolic = None
        self.sharding = None
        self.Q = self.lam = None
        self.profile = {"nnodes": self.nnodes, "nelems": self.nelems, "solver_type": solver_type,
                        "adjoint_method": adjoint_method, "N": N}

    # ---- forward (:548-640, 811-838) ----------------------------------------------------------------
    def solve_eigenvalue_problem(self, store=False):
        t0 = _now()
        Kr = self.prob.assemble_K()
        if self.symbolic is None:                                         # once per sparsity pattern
            ts = _now()
            coords, dofpn = self.prob.dof_coords()
            sym = D.Symbolic(self.prob.indptr, self.prob.indices, self.prob.nred, coords=coords, dof_per_node=dofpn)
            self.symbolic = (sym, sym.assembly_map_device(Kr.indptr, Kr.indices))
            self.profile["symbolic analysis time"] = _now() - ts
            t0 += self.profile["symbolic analysis time"]
        self.Kfact = SpLuOperator(Kr, symbolic=self.symbolic, max_rhs=1)  # fundamental path K_r u_r = f_r
        ur = self.Kfact.solve_dev(self.prob.fr_d)
        self.u_d = self.prob.set_displacement(ur)
        Gr = self.prob.assemble_G()
        t1 = _now()
        self.profile["matrix assembly time"] = t1 - t0
        vals = D.axpby(1.0, Kr.data, float(self.sigma), Gr.data)          # K_r + sigma G_r on the shared pattern
        self.factor = SpLuOperator(Kr.with_values(vals), symbolic=self.symbolic)
        self.factor.count = 0
        self.Kr, self.Gr = Kr, Gr
        if self.solver_type == "IRAM":
            if self.m is None:
                self.m = max(2 * self.N + 1, 60)
            self.eig_solver = IRAM(N=self.N, m=self.m, eig_atol=self.eig_atol, mode="buckling")
            self.eig_solver.seed = self.seed
            self.eig_solver.reference_pairing = getattr(self, "reference_pairing", False)
            self.eig_solver.block_size = getattr(self, "block_size", None)
        else:
            if self.m is None:
                self.m = max(3 * self.N + 1, 60)
            self.eig_solver = BasicLanczos(N=self.N, m=self.m, eig_atol=self.eig_atol, tol=self.tol, mode="buckling")
        self.eig_solver.sharding = self.sharding
        mu, _ = self.eig_solver.solve(Gr, Kr, self.factor, self.sigma)
        t2 = _now()
        self.profile["solve preconditioner count"] = self.factor.count
        self.profile["eigenvalue solve time"] = t2 - t1
        self.profile["m"] = self.m
        self.profile["eig_solver.m"] = str(self.eig_solver.m)
        self.BLF = np.asarray(mu)[: self.N]
        self.Qr = self.eig_solver._Phi_d                                   # (n_r, N) device
        return self.BLF, self.Qr

    def initialize(self, store=False, x=None):
        if x is not None:
            self.x = x
        self.x_d = to_dev(self.x)
        self.rho = self.fltr.apply(self.x_d)
        self.rhoE = self.prob.set_density(rho=self.rho)
        self.lam, self.Qr = self.solve_eigenvalue_problem(store)
        return

    @property
    def u(self):
        return to_host(self.u_d)

    def compliance(self):
        return float(np.dot(self.f, self.u))

    def initialize_adjoint(self):
        self.xb = D.zeros(self.x_d.shape[0])
        self.rhoEb = D.zeros(self.nelems)
        self.lamb = np.zeros(self.N)
        self.Qrb = D.zeros(self.prob.nred, self.N)

    # ---- objectives (:702-760) ----------------------------------------------------------------------
    def _aggregate_weights(self, rho, mode):
        lam = np.asarray(self.lam)
        if mode == "exp":
            eta = np.exp(-rho * (lam - np.min(lam)))
            a = b = None
        else:
            a = np.tanh(rho * (lam - 0.0))
            b = np.tanh(rho * (lam - 50.0))
            eta = a - b
        return eta / np.sum(eta), a, b

    def get_eigenvector_aggregate(self, rho, node, mode="tanh"):
        """h = sum_i eta_i Q[node, i]^2 (:702-724); as in the reference ``node`` indexes the FULL dof vector."""
        eta, _, _ = self._aggregate_weights(rho, mode)
        qn = self._dof_values(node)
        return float(np.sum(eta * qn * qn))

    def _dof_values(self, dof):
        f2r = np.full(self.nvars, -1, dtype=np.int64)
        f2r[self.reduced] = np.arange(len(self.reduced))
        r = int(f2r[dof])
        self._dof_row = r
        return to_host(self.Qr[r]) if r >= 0 else np.zeros(self.N)

    def add_eigenvector_aggregate_derivative(self, hb, rho, node, mode="tanh"):
        """Seeds of h = sum_i eta_i Q[node, i]^2 (:726-760): Qrb[row(node), i] += 2 hb eta_i Q[node, i] and
        lamb_i -= hb rho eta_i (a_i + b_i) (Q[node, i]^2 - h)."""
        eta, a, b = self._aggregate_weights(rho, mode)
        qn = self._dof_values(node)
        h = float(np.sum(eta * qn * qn))
        if self._dof_row >= 0:
            row = self.Qrb[self._dof_row]
            row += small_to_dev(2.0 * hb * eta * qn)
        if mode == "exp":
            self.lamb -= hb * rho * eta * (qn * qn - h)
        else:
            self.lamb -= hb * rho * eta * (a + b) * (qn * qn - h)
        return h

    def eval_ks_buckling(self, ks_rho=160.0):
        """KS maximum of mu_i = 1 / BLF_i (:641-646)."""
        mu = 1.0 / np.asarray(self.BLF)
        c = np.max(mu)
        return float(c + np.log(np.sum(np.exp(ks_rho * (mu - c)))) / ks_rho)

    def eval_ks_buckling_derivative(self, ks_rho=160.0):
        """Design gradient of ``eval_ks_buckling`` (:648-700, tensor form): eigenvalue sensitivities only (no eigenvector
        adjoint), with the fundamental-path adjoint K_r a_r = -(dG/du)_r.  Everything stays in HBM; returns the
        gradient in the flavour of ``self.x``."""
        t0 = _now()
        pr = self.prob
        mu = 1.0 / np.asarray(self.BLF)
        eta = np.exp(ks_rho * (mu - np.max(mu)))
        eta /= np.sum(eta)
        Q = self.Qr
        etaQ = Q.clone()
        D.col_scale(etaQ, small_to_dev(eta), mode=0)
        etamuQ = Q.clone()
        D.col_scale(etamuQ, small_to_dev(eta * mu), mode=0)
        dKdx = pr.dKdx.device_call(etamuQ, Q)                               # per element
        dGdu = pr.dGdu.device_call(etaQ, Q)                                 # full dof vector
        adj_r = self.Kfact.solve_dev(pr.reduce_vector(dGdu))
        adj = pr.full_vector(adj_r)                                         # = -adjoint of the reference (:689-690)
        dGdx = pr.dGdx.device_call(etaQ, Q)
        dfde = D.axpby(-1.0, dGdx, -1.0, dKdx)                              # -(dGdx + dKdx) ...
        D.axpby(1.0, dfde, 1.0, pr.dK_single(adj, self.u_d), out=dfde)      # ... - dK(adj_ref, u) = + dK(adj, u)
        dfdrho = pr.scatter_to_nodes(dfde)
        g = self.fltr.apply_gradient(dfdrho, self.x_d)
        self.profile["total derivative time"] = self.profile.get("total derivative time", 0.0) + (_now() - t0)
        return like_input(g, self.x)

    # ---- reverse (:866-982) -------------------------------------------------------------------------
    def finalize_adjoint(self):
        res_list = []
        self.factor.count = 0
        t0 = _now()
        psir, corr_data = self.eig_solver.solve_adjoint(self.Qrb, rtol=self.rtol, method=self.adjoint_method,
                                                        callback=res_list.append, **self.adjoint_options)
        t1 = _now()
        self.psir = psir
        self.profile["adjoint preconditioner count"] = self.factor.count
        self.profile["adjoint solution time"] = t1 - t0
        self.profile["adjoint residuals"] = res_list
        self.profile["adjoint iterations"] = len(res_list)
        self.profile["adjoint correction data"] = corr_data
        pr = self.prob
        # d/du of the eigen-terms: dfdu0 = sum_i w_i^T (dG/du) phi_i  (dB/du = 0: K does not depend on u)
        dfdu0 = D.zeros(self.nvars)
        self.eig_solver.add_total_derivative(self.lamb, self.Qrb, psir, pr.dGdu, None, dfdu0, adj_corr_data=corr_data,
                                             deriv_type=self.deriv_type)
        # explicit design dependence of G and K
        self.eig_solver.add_total_derivative(self.lamb, self.Qrb, psir, pr.dGdx, pr.dKdx, self.rhoEb,
                                             adj_corr_data=corr_data, deriv_type=self.deriv_type)
        # adjoint of the fundamental path: K_r psi_u = -dfdu0_r, rhob += psi_u^T (dK/drho) u   (:974-979)
        psiu_r = self.Kfact.solve_dev(pr.reduce_vector(dfdu0))
        psiu = pr.full_vector(psiu_r)
        D.axpby(1.0, self.rhoEb, -1.0, pr.dK_single(psiu, self.u_d), out=self.rhoEb)
        self.rhob = pr.scatter_to_nodes(self.rhoEb)
        g = self.fltr.apply_gradient(self.rhob, self.x_d)
        D.axpby(1.0, self.xb, 1.0, g, out=self.xb)
        t2 = _now()
        self.profile["total derivative time"] = t2 - t1
        return

    def time_to_gradient(self):
        p = self.profile
        return p["eigenvalue solve time"] + p["adjoint solution time"] + p["total derivative time"]


def domain_compressed_column(nx=64, ny=128, Lx=1.0, Ly=2.0, shear_force=False):
    """examples/buckling.py ``domain_compressed_column`` (:1300-1369), vectorised: mesh, left-right symmetry map
    of the design variables, clamped bottom edge, compressive (or shear) load on the top edge."""
    conn, X = fe.grid_mesh(nx, ny, Lx, Ly)
    nodes = np.arange((nx + 1) * (ny + 1), dtype=np.int64).reshape(nx + 1, ny + 1)
    dvmap = np.zeros((nx + 1, ny + 1), dtype=np.int64)
    index = 0
    for i in range(nx // 2 + 1):                      # same visiting order as the reference (:1336-1341)
        dvmap[i, :] = index + np.arange(ny + 1)
        dvmap[nx - i, :] = dvmap[i, :]
        index += ny + 1
    bcs = {int(nodes[i, 0]): [0, 1] for i in range(nx + 1)}
    P = 1e-3
    forces = {}
    if shear_force:
        for i in range(nx + 1):
            forces[int(nodes[i, ny])] = [P / (nx + 1), 0]
    else:
        offset = int(np.ceil(nx / 30))
        for i in range(offset):
            forces[int(nodes[nx // 2 - i - 1, ny])] = [0, -P / (2 * offset + 1)]
            forces[int(nodes[nx // 2 + i + 1, ny])] = [0, -P / (2 * offset + 1)]
        forces[int(nodes[nx // 2, ny])] = [0, -P / (2 * offset + 1)]
    return conn, X, dvmap.flatten(), index, bcs, forces


def make_buckling_model(nx=64, ny=128, Lx=1.0, Ly=2.0, rfact=4.0, N=10, shear_force=False, **kwargs):
    """examples/buckling.py ``make_model`` (:1372-1409)."""
    conn, X, dvmap, ndv, bcs, forces = domain_compressed_column(nx=nx, ny=ny, Lx=Lx, Ly=Ly, shear_force=shear_force)
    fltr = fe.NodeFilter(conn, X, r0=rfact * (Lx / nx), dvmap=dvmap, num_design_vars=ndv)
    return BucklingTopologyAnalysis(fltr, conn, X, bcs=bcs, forces=forces, N=N, **kwargs)
