"""Device-side finite-element operators for the Q4 problems of the reference examples.

Host part (numpy, runs once per mesh): mesh numbering, DOF maps, the CSR pattern of K/M and the
integer assembly maps (element, a, b) -> CSR non-zero, node -> element adjacency, the conic
filter pattern.  Device part (kernels of csrc/fe.cu): material interpolation, gather-form
assembly of K and M, the bilinear sensitivity forms ``w^T (dK/drho_e) v`` / ``w^T (dM/drho_e) v``
summed over modes, element -> node gather, filter products.

Mirrors (paths relative to the reference root):
  mesh / connectivity       examples/thermal.py:1475-1498, examples/natural_frequency.py:850-894
  DOF map and COO lists     examples/thermal.py:79-92, examples/natural_frequency.py:90-104
  K, M assembly             examples/thermal.py:126-148,192-214, examples/natural_frequency.py:134-160,205-236
  dK, dM callbacks          examples/thermal.py:150-190,216-246, examples/natural_frequency.py:162-203,238-284
  node scatter              examples/thermal.py:612-615
  node filter               examples/node_filter.py:61-88 (construction), :164-217 (apply / gradient)
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import device as D
from ._hostdev import is_dev, to_dev, to_host, like_input


def grid_mesh(nx, ny, Lx=1.0, Ly=1.0):
    """Structured Q4 mesh with the examples' numbering: node(i, j) = i*(ny+1) + j,
    element e = i + nx*j, counter-clockwise connectivity (examples/thermal.py:1475-1498)."""
    nodes = np.arange((nx + 1) * (ny + 1), dtype=np.int64).reshape(nx + 1, ny + 1)
    X = np.zeros(((nx + 1) * (ny + 1), 2))
    X[:, 0] = np.repeat(np.linspace(0, Lx, nx + 1), ny + 1)
    X[:, 1] = np.tile(np.linspace(0, Ly, ny + 1), nx + 1)
    conn = np.empty((nx * ny, 4), dtype=np.int64)
    conn[:, 0] = nodes[:-1, :-1].T.ravel()
    conn[:, 1] = nodes[1:, :-1].T.ravel()
    conn[:, 2] = nodes[1:, 1:].T.ravel()
    conn[:, 3] = nodes[:-1, 1:].T.ravel()
    return conn, X


def element_dofs(conn, dof):
    """var[e] = the element's global DOF list (examples/natural_frequency.py:90-92)."""
    conn = np.asarray(conn, dtype=np.int64)
    return (conn[:, :, None] * dof + np.arange(dof)[None, None, :]).reshape(conn.shape[0], -1)


def assembly_structure(var, ndof):
    """CSR pattern of sum_e P_e^T K_e P_e plus, for every non-zero, the list of (e, a, b) sources.

    The (i, j) lists are in ``Ke.flatten()`` order (examples/natural_frequency.py:94-104) and the
    CSR is what ``coo_matrix((vals, (i, j))).tocsr()`` produces (:157-158): duplicates summed,
    column indices sorted.  Returns indptr (int32), indices (int32), src_ptr (int64, nnz+1),
    src (int64, flat COO positions e*ne*ne + a*ne + b grouped by non-zero, ascending)."""
    ne = var.shape[1]
    i = np.repeat(var, ne, axis=1).ravel()
    j = np.tile(var, (1, ne)).ravel()
    key = i * np.int64(ndof) + j
    order = np.argsort(key, kind="stable")
    skey = key[order]
    first = np.ones(len(skey), dtype=bool)
    first[1:] = skey[1:] != skey[:-1]
    ukey = skey[first]
    rows = (ukey // ndof).astype(np.int64)
    indices = (ukey % ndof).astype(np.int32)
    indptr = np.zeros(ndof + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    indptr = np.cumsum(indptr).astype(np.int32)
    src_ptr = np.concatenate([np.nonzero(first)[0], [len(skey)]]).astype(np.int64)
    return indptr, indices, src_ptr, order.astype(np.int64)


def node_adjacency(conn, nnodes):
    """CSR node -> elements (gather form of the ``np.add.at`` scatter, examples/thermal.py:612-615)."""
    conn = np.asarray(conn, dtype=np.int64)
    e = np.repeat(np.arange(conn.shape[0], dtype=np.int64), conn.shape[1])
    v = conn.ravel()
    order = np.lexsort((e, v))
    nptr = np.zeros(nnodes + 1, dtype=np.int64)
    np.add.at(nptr, v + 1, 1)
    return np.cumsum(nptr).astype(np.int32), e[order].astype(np.int32)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class _Half:
    """One of the two reference callbacks ``dAdx(w, v)`` / ``dBdx(w, v)``."""

    def __init__(self, parent, which):
        self.parent, self.which = parent, which
        self.fused_with = None

    def device_call(self, W, V):
        if self.which == "A":
            return self.parent.device_call_fused(W, None, V, 1.0, 0.0)
        return self.parent.device_call_fused(None, W, V, 0.0, 1.0)

    def __call__(self, w, v):
        out = self.device_call(to_dev(w), to_dev(v))
        return like_input(out, w)


class Q4Problem:
    """Thermal (1 DOF/node) or plane-stress (2 DOF/node) Q4 model with its operators in HBM.

    kind "thermal": K_e = kappa(rho_e) sum_q detJ Be^T Be, M_e = c(rho_e) sum_q detJ N N^T
      with kappa = kappa0((1-beta) rho^p + beta), c = cp*density*((1-beta) rho + beta)
      (examples/thermal.py:23-28,132,198).
    kind "plane_stress": K_e = s(rho_e) sum_q detJ Be^T C0 Be, M_e = density*rho_e sum_q detJ He^T He
      with s = rho^p + rho0_K (SIMP) or rho/(1+q(1-rho)) + rho0_K (RAMP)
      (examples/natural_frequency.py:83-86,140-155,219-231).
    """

    def __init__(self, conn, X, kind="thermal", E=1.0, nu=0.3, kappa=1.0, density=1.0, heat_capacity=1.0, p=3.0,
                 beta=1e-6, rho0_K=1e-6, ptype_K="simp", q=5.0):
        if kind not in ("thermal", "plane_stress"):
            raise ValueError("unknown kind %r" % kind)
        dev = D.dev()
        self.kind = kind
        self.kid = 0 if kind == "thermal" else 1
        self.conn = np.asarray(conn, dtype=np.int64)
        self.X = np.asarray(X, dtype=np.float64)
        self.nelems = self.conn.shape[0]
        self.nnodes = int(self.conn.max()) + 1
        self.dof = 1 if kind == "thermal" else 2
        self.ndof = self.dof * self.nnodes
        self.var = element_dofs(self.conn, self.dof)
        self.indptr, self.indices, src_ptr, src = assembly_structure(self.var, self.ndof)
        self.nnz = len(self.indices)
        nptr, nelem = node_adjacency(self.conn, self.nnodes)
        if kind == "thermal":
            self.law, self.par = 0, np.array([p, kappa, heat_capacity * density, beta])
        elif ptype_K == "simp":
            self.law, self.par = 1, np.array([p, 1.0, density, rho0_K])
        elif ptype_K == "ramp":
            self.law, self.par = 2, np.array([q, 1.0, density, rho0_K])
        else:
            raise ValueError("unknown ptype_K %r" % ptype_K)
        C0 = E * np.array([[1.0, nu, 0.0], [nu, 1.0, 0.0], [0.0, 0.0, 0.5 * (1.0 - nu)]]) / (1.0 - nu**2)
        self.C0 = C0
        # ---- device mirrors -------------------------------------------------------------
        self.conn_d = torch.as_tensor(self.conn.astype(np.int32), device=dev)
        self.xy_d = to_dev(self.X)
        self.cmat6_d = to_dev(np.array([C0[0, 0], C0[0, 1], C0[0, 2], C0[1, 1], C0[1, 2], C0[2, 2]]))
        self.indptr_d = torch.as_tensor(self.indptr, device=dev)
        self.indices_d = torch.as_tensor(self.indices, device=dev)
        self.src_ptr_d = torch.as_tensor(src_ptr, device=dev)
        self.src_d = torch.as_tensor(src, device=dev)
        self.nptr_d = torch.as_tensor(nptr, device=dev)
        self.nelem_d = torch.as_tensor(nelem, device=dev)
        self.rhoE_d = D.empty(self.nelems)
        self.ks_d, self.ms_d = D.empty(self.nelems), D.empty(self.nelems)
        self.dk_d, self.dm_d = D.empty(self.nelems), D.empty(self.nelems)
        self._par_c = (ctypes.c_double * 4)(*[float(v) for v in self.par])
        self.sharding = None          # dist.ModeSharding: element-range shards of the df/dx reduction
        self.dAdx, self.dBdx = _Half(self, "A"), _Half(self, "B")
        self.dAdx.fused_with, self.dBdx.fused_with = self.dBdx, self.dAdx

    # ---- material -----------------------------------------------------------------------------
    def set_density(self, rho=None, rhoE=None):
        """Nodal density rho (filtered design) -> element density + material factors, or set the
        element densities directly."""
        lib = _lib.load()
        if rhoE is not None:
            self.rhoE_d.copy_(to_dev(rhoE))
            _lib.check(lib.eigd_q4_material(self.law, self.nelems, None, _ptr(self.rhoE_d), self._par_c, None,
                                            _ptr(self.ks_d), _ptr(self.ms_d), _ptr(self.dk_d), _ptr(self.dm_d)), "q4_material")
        else:
            rho_d = to_dev(rho)
            _lib.check(lib.eigd_q4_material(self.law, self.nelems, _ptr(self.conn_d), _ptr(rho_d), self._par_c,
                                            _ptr(self.rhoE_d), _ptr(self.ks_d), _ptr(self.ms_d), _ptr(self.dk_d),
                                            _ptr(self.dm_d)), "q4_material")
        return self.rhoE_d

    def assemble_tangent(self, rho_t):
        """Directional derivative of (K, M) along the nodal density direction ``rho_t`` (device, (nnodes,)), on the same
        pattern: the assembly kernel with the material factors replaced by (dk/drho_e) rho_e', (dm/drho_e) rho_e'.
        Used by the dual-number (complex-step) mode of the drivers."""
        lib = _lib.load()
        rhoEt, s1, s2, s3, s4 = (D.empty(self.nelems) for _ in range(5))
        # element mean of the nodal direction (the material outputs of this call are scratch)
        _lib.check(lib.eigd_q4_material(self.law, self.nelems, _ptr(self.conn_d), _ptr(to_dev(rho_t)), self._par_c, _ptr(rhoEt),
                                        _ptr(s1), _ptr(s2), _ptr(s3), _ptr(s4)), "q4_material")
        kt, mt = self.dk_d * rhoEt, self.dm_d * rhoEt
        Kv, Mv = D.empty(self.nnz), D.empty(self.nnz)
        D.q4_assemble(self.kid, self.conn_d, self.xy_d, kt, mt, self.cmat6_d, self.src_ptr_d, self.src_d, self.nnz, Kv, Mv)
        shape = (self.ndof, self.ndof)
        return (D.CsrDevice(self.indptr_d, self.indices_d, Kv, shape), D.CsrDevice(self.indptr_d, self.indices_d, Mv, shape))

    # ---- assembly -------------------------------------------------------------------------------
    def assemble(self):
        """K(rho), M(rho) as CsrDevice sharing one pattern (values computed in gather form)."""
        Kv, Mv = D.empty(self.nnz), D.empty(self.nnz)
        D.q4_assemble(self.kid, self.conn_d, self.xy_d, self.ks_d, self.ms_d, self.cmat6_d, self.src_ptr_d, self.src_d,
                      self.nnz, Kv, Mv)
        shape = (self.ndof, self.ndof)
        K = D.CsrDevice(self.indptr_d, self.indices_d, Kv, shape)
        M = D.CsrDevice(self.indptr_d, self.indices_d, Mv, shape)
        return K, M

    def dof_coords(self):
        """Node coordinates for geometric nested dissection (``SpLuOperator(coords=..., dof_per_node=...)``)."""
        return self.X, self.dof

    # ---- sensitivities ----------------------------------------------------------------------------
    def device_call_fused(self, WA, WB, V, cA, cB):
        """cA * sum_k WA_k^T (dK/drho_e) V_k + cB * sum_k WB_k^T (dM/drho_e) V_k per element (device)."""
        V2 = V if V.dim() == 2 else V.unsqueeze(1)
        WA2 = None if WA is None else (WA if WA.dim() == 2 else WA.unsqueeze(1))
        WB2 = None if WB is None else (WB if WB.dim() == 2 else WB.unsqueeze(1))
        V2 = V2 if V2.is_contiguous() else V2.contiguous()
        WA2 = None if WA2 is None else (WA2 if WA2.is_contiguous() else WA2.contiguous())
        WB2 = None if WB2 is None else (WB2 if WB2.is_contiguous() else WB2.contiguous())
        shard = self.sharding
        if shard is None or shard.world == 1:
            out = D.zeros(self.nelems)
            D.q4_quadforms(self.kid, self.conn_d, self.xy_d, self.cmat6_d, WA2, WB2, V2, self.dk_d, self.dm_d,
                           float(cA), -float(cB), out)
            return out
        lo, hi = shard.my_range(self.nelems)                  # element-range shard + one all-gather (dist.py)
        part = D.zeros(hi - lo)
        if hi > lo:
            D.q4_quadforms(self.kid, self.conn_d[lo:hi], self.xy_d, self.cmat6_d, WA2, WB2, V2, self.dk_d[lo:hi],
                           self.dm_d[lo:hi], float(cA), -float(cB), part)
        return shard.allgather_ranges(part, self.nelems)

    def scatter_to_nodes(self, evals, scale=0.25):
        """rho_b[v] = scale * sum_{e ni v} evals[e] (examples/thermal.py:612-615)."""
        ev = to_dev(evals)
        out = D.empty(self.nnodes)
        D.node_gather(self.nptr_d, self.nelem_d, ev, scale, out)
        return like_input(out, evals)


def reduced_structure(indptr, indices, src_ptr, src, reduced, ndof):
    """Pattern and (e, a, b) source lists of ``matrix[reduced, :][:, reduced]`` (examples/buckling.py:505-510)
    taken from those of the full matrix; the result equals scipy's fancy-indexed CSR bit for bit
    (``reduced`` ascending keeps rows and sorted columns in order)."""
    reduced = np.asarray(reduced, dtype=np.int64)
    f2r = np.full(ndof, -1, dtype=np.int64)
    f2r[reduced] = np.arange(len(reduced))
    rows = np.repeat(np.arange(ndof, dtype=np.int64), np.diff(indptr))
    keep = (f2r[rows] >= 0) & (f2r[indices] >= 0)
    cnt = np.zeros(len(reduced) + 1, dtype=np.int64)
    np.add.at(cnt, f2r[rows[keep]] + 1, 1)
    r_indptr = np.cumsum(cnt).astype(np.int32)
    r_indices = f2r[indices[keep]].astype(np.int32)
    lens = np.diff(src_ptr)[keep]
    r_src_ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    start = src_ptr[:-1][keep]
    pos = np.repeat(start - r_src_ptr[:-1], lens) + np.arange(int(r_src_ptr[-1]), dtype=np.int64)
    return r_indptr, r_indices, r_src_ptr, src[pos]


class _BucklingCallback:
    """dAdu / dAdx / dBdx of examples/buckling.py:924-979 as device operators on REDUCED (n_r, N) operands."""

    def __init__(self, parent, which):
        self.parent, self.which = parent, which
        self.fused_with = None

    def device_call(self, Wr, Vr):
        pr = self.parent
        W, V = pr.full_vector(Wr), pr.full_vector(Vr)
        W2 = W if W.dim() == 2 else W.unsqueeze(1)
        V2 = V if V.dim() == 2 else V.unsqueeze(1)
        if self.which == "dGdu":
            due = D.empty(pr.nelems, 8)
            D.q4_gderiv(pr.conn_d, pr.xy_d, pr.cmat6_d, W2, V2, pr.kg_d, pr.dkg_d, pr.u_d, 1.0, None, due)
            return D.q4_dof_gather(pr.nptr_d, pr.nelem_d, pr.nlocal_d, due, D.empty(pr.ndof))
        out = D.zeros(pr.nelems)
        if self.which == "dGdx":
            D.q4_gderiv(pr.conn_d, pr.xy_d, pr.cmat6_d, W2, V2, pr.kg_d, pr.dkg_d, pr.u_d, 1.0, out, None)
        else:                                                                     # dKdx
            D.q4_quadforms(1, pr.conn_d, pr.xy_d, pr.cmat6_d, W2, None, V2, pr.base.dk_d, pr.base.dm_d, 1.0, 0.0, out)
        return out

    def __call__(self, w, v):
        return like_input(self.device_call(to_dev(w), to_dev(v)), w)


class BucklingQ4Problem:
    """Plane-stress Q4 column of examples/buckling.py in HBM: K(x) with Dirichlet rows removed, the fundamental
    path K_r u_r = f_r, the stress stiffness G(u, x) and the three sensitivity callbacks.

    Element-level outputs stay per element (the node scatter ``np.add.at ... *0.25`` of :212-216, :337-341 is
    linear and is applied once at the end by the driver).  Reference: examples/buckling.py:152-343, 499-518.
    """

    def __init__(self, conn, X, bcs, forces, E=1.0, nu=0.3, density=1.0, p=3.0, q=5.0, rho0_K=1e-6, rho0_G=1e-9,
                 ptype_K="simp", ptype_G="simp"):
        dev = D.dev()
        self.base = Q4Problem(conn, X, "plane_stress", E=E, nu=nu, density=density, p=p, rho0_K=rho0_K,
                              ptype_K=ptype_K, q=q)
        b = self.base
        self.conn, self.X = b.conn, b.X
        self.nelems, self.nnodes, self.ndof = b.nelems, b.nnodes, b.ndof
        self.conn_d, self.xy_d, self.cmat6_d = b.conn_d, b.xy_d, b.cmat6_d
        self.nptr_d, self.nelem_d = b.nptr_d, b.nelem_d
        # Dirichlet conditions and loads (:120-150)
        fixed = np.zeros(self.ndof, dtype=bool)
        for node, uv in bcs.items():
            for index in uv:
                fixed[2 * int(node) + int(index)] = True
        self.reduced = np.nonzero(~fixed)[0]
        self.nred = len(self.reduced)
        self.f = np.zeros(self.ndof)
        for node, fv in forces.items():
            self.f[2 * int(node)] += fv[0]
            self.f[2 * int(node) + 1] += fv[1]
        src_ptr, src = to_host_i64(b.src_ptr_d), to_host_i64(b.src_d)
        self.indptr, self.indices, r_src_ptr, r_src = reduced_structure(b.indptr, b.indices, src_ptr, src, self.reduced,
                                                                        self.ndof)
        self.nnz = len(self.indices)
        self.indptr_d = torch.as_tensor(self.indptr, device=dev)
        self.indices_d = torch.as_tensor(self.indices, device=dev)
        self.src_ptr_d = torch.as_tensor(r_src_ptr, device=dev)
        self.src_d = torch.as_tensor(r_src, device=dev)
        self.reduced_d = torch.as_tensor(self.reduced.astype(np.int32), device=dev)
        self.fr_d = to_dev(self.f[self.reduced])
        # local index of every node inside each adjacent element (dof gather of :318-320)
        e = np.repeat(np.arange(self.nelems, dtype=np.int64), 4)
        v = self.conn.ravel()
        loc = np.tile(np.arange(4, dtype=np.int64), self.nelems)
        order = np.lexsort((e, v))
        self.nlocal_d = torch.as_tensor(loc[order].astype(np.int32), device=dev)
        # G material law (:231-234) and the derivative the reference applies to it (:333-336)
        if ptype_G == "simp":
            self.lawG, parG, pardG = 1, [p, 1.0, density, rho0_G], [p, 1.0, density, rho0_G]
        elif ptype_G == "ramp":
            self.lawG, parG, pardG = 2, [q, 1.0, density, rho0_G], [q + 1.0, 1.0, density, rho0_G]
        else:
            raise ValueError("unknown ptype_G %r" % ptype_G)
        self._parG = (ctypes.c_double * 4)(*[float(t) for t in parG])
        self._pardG = (ctypes.c_double * 4)(*[float(t) for t in pardG])
        self.kg_d, self.dkg_d = D.empty(self.nelems), D.empty(self.nelems)
        self._scratch = D.empty(self.nelems)
        self.u_d = None
        self.dGdu = _BucklingCallback(self, "dGdu")
        self.dGdx = _BucklingCallback(self, "dGdx")
        self.dKdx = _BucklingCallback(self, "dKdx")

    # ---- Dirichlet maps (:499-518) ----------------------------------------------------------------
    def full_vector(self, vr):
        return D.expand_rows(self.reduced_d, vr, self.ndof)

    def reduce_vector(self, v):
        return D.reduce_rows(self.reduced_d, v)

    def dof_coords(self):
        """Per-dof coordinates of the reduced system for geometric nested dissection."""
        return self.X[self.reduced // 2], 1

    # ---- material ------------------------------------------------------------------------------------
    def set_density(self, rho=None, rhoE=None):
        rhoE_d = self.base.set_density(rho=rho, rhoE=rhoE)
        lib = _lib.load()
        s = self._scratch
        _lib.check(lib.eigd_q4_material(self.lawG, self.nelems, None, _ptr(rhoE_d), self._parG, None, _ptr(self.kg_d),
                                        _ptr(s), _ptr(self.dkg_d), _ptr(s)), "q4_material")
        if self.lawG == 2:      # the reference differentiates the RAMP law of G with q + 1 (:336)
            tmp = D.empty(self.nelems)
            _lib.check(lib.eigd_q4_material(self.lawG, self.nelems, None, _ptr(rhoE_d), self._pardG, None, _ptr(tmp),
                                            _ptr(s), _ptr(self.dkg_d), _ptr(s)), "q4_material")
        return rhoE_d

    # ---- assembly -------------------------------------------------------------------------------------
    def assemble_K(self):
        """K_r(rho) = reduce_matrix(get_stiffness_matrix(rhoE)) (:152-176, 505-510), values gathered directly
        through the reduced source lists."""
        b = self.base
        Kv, Mv = D.empty(self.nnz), D.empty(self.nnz)
        D.q4_assemble(1, self.conn_d, self.xy_d, b.ks_d, b.ms_d, self.cmat6_d, self.src_ptr_d, self.src_d, self.nnz, Kv, Mv)
        return D.CsrDevice(self.indptr_d, self.indices_d, Kv, (self.nred, self.nred))

    def set_displacement(self, ur_d):
        """u = full_vector(u_r): the fundamental path the stress stiffness is linearised about (:561-562)."""
        self.u_d = self.full_vector(ur_d)
        return self.u_d

    def assemble_G(self):
        """G_r(u, rho) = reduce_matrix(get_stress_stiffness_matrix(rhoE, u)) (:220-255)."""
        sdet = D.empty(self.nelems, 4, 3)
        D.q4_stress(self.conn_d, self.xy_d, self.cmat6_d, self.kg_d, self.u_d, sdet)
        Gv = D.empty(self.nnz)
        D.q4_assemble_geometric(self.conn_d, self.xy_d, sdet, self.src_ptr_d, self.src_d, self.nnz, Gv)
        return D.CsrDevice(self.indptr_d, self.indices_d, Gv, (self.nred, self.nred))

    def dK_single(self, psi_full, u_full):
        """psi^T (dK/drho_e) u per element for one pair of full vectors (:178-218 before the node scatter)."""
        out = D.zeros(self.nelems)
        b = self.base
        D.q4_quadforms(1, self.conn_d, self.xy_d, self.cmat6_d, psi_full.reshape(-1, 1), None, u_full.reshape(-1, 1),
                       b.dk_d, b.dm_d, 1.0, 0.0, out)
        return out

    def scatter_to_nodes(self, evals, scale=0.25):
        return self.base.scatter_to_nodes(evals, scale)


def _contig1(t):
    return t if t.is_contiguous() else t.contiguous()


def to_host_i64(t):
    return t.detach().cpu().numpy().astype(np.int64)


class NodeFilter:
    """Conic node filter F[i, j] ~ max(0, r0 - |X_i - X_j|), rows normalised to 1
    (examples/node_filter.py:61-88), applied on the device as CSR products (:164-217).
    Spatial filter with optional design-variable map (the symmetric column of examples/buckling.py:1331-1344)
    and optional tanh projection (``projection=True``, ``beta``, ``eta``)."""

    ftype = "spatial"

    def __init__(self, conn, X, r0=1.0, ftype="spatial", dvmap=None, num_design_vars=None, beta=10.0, eta=0.5,
                 projection=False):
        from scipy import sparse, spatial
        if ftype not in ("spatial", "helmholtz"):
            raise ValueError("unknown filter type %r" % ftype)
        self.ftype = ftype
        self.projection = bool(projection)
        self.beta = 10.0 if beta is None else float(beta)      # reference default (node_filter.py:30-31)
        self.eta = float(eta)
        self.X = np.asarray(X, dtype=np.float64)
        self.nnodes = self.X.shape[0]
        self.r0 = r0
        self.dvmap = None if dvmap is None else np.asarray(dvmap, dtype=np.int64)
        self.offset_d = None
        if ftype == "helmholtz":
            self._init_helmholtz(conn, r0, num_design_vars)
            return
        tree = spatial.cKDTree(self.X)
        Dm = tree.sparse_distance_matrix(tree, r0, output_type="coo_matrix")
        w = r0 - Dm.data
        F = sparse.coo_matrix((w, (Dm.row, Dm.col)), shape=(self.nnodes, self.nnodes)).tocsr()
        rs = np.asarray(F.sum(axis=1)).ravel()
        F = (sparse.diags(1.0 / rs) @ F).tocsr()
        self.dvmap = None if dvmap is None else np.asarray(dvmap, dtype=np.int64)
        self.offset_d = None
        if self.dvmap is None:
            self.num_design_vars = self.nnodes
        else:
            # design-variable map (examples/node_filter.py:164-168, 211-214): rho = F x[dvmap] with x := 1 where
            # dvmap < 0, gradient scattered back with np.add.at.  Folded into the operator: F <- F P,
            # P[i, dvmap[i]] = 1, plus the constant F 1_{dvmap<0}; the transpose product is then the gather form
            # of the np.add.at scatter.
            self.num_design_vars = int(num_design_vars if num_design_vars is not None else self.dvmap.max() + 1)
            act = self.dvmap >= 0
            P = sparse.coo_matrix((np.ones(int(act.sum())), (np.nonzero(act)[0], self.dvmap[act])),
                                  shape=(self.nnodes, self.num_design_vars)).tocsr()
            if not act.all():
                self.offset_d = to_dev(F @ (~act).astype(np.float64))
            F = (F @ P).tocsr()
        F.sort_indices()
        FT = F.T.tocsr()
        FT.sort_indices()
        self.F, self.FT = F, FT
        self.F_d = D.CsrDevice.from_scipy(F)
        self.FT_d = D.CsrDevice.from_scipy(FT)

    def _init_helmholtz(self, conn, r0, num_design_vars):
        """Helmholtz (PDE) filter of examples/node_filter.py:90-162: rho = A^-1 (B x) with A = r0^2 int grad N . grad N +
        int N N and B = int N N on the Q4 mesh.  A and B are the unit-material thermal stiffness / capacity matrices of
        ``Q4Problem`` (assembled by the same gather kernel), A is factored once by the GPU LDL^T (the reference uses
        scipy's ``factorized``) and applied to one right-hand side per call."""
        from .eigenvector_derivatives import SpLuOperator
        prob = Q4Problem(conn, self.X, "thermal", kappa=1.0, density=1.0, heat_capacity=1.0, p=1.0, beta=0.0)
        prob.set_density(rhoE=np.ones(prob.nelems))
        K, M = prob.assemble()
        self._hB = M
        vals = D.axpby(float(r0) ** 2, K.data, 1.0, M.data)
        self._hA = SpLuOperator(K.with_values(vals), coords=self.X, dof_per_node=1, max_rhs=1)
        self._hprob = prob
        if self.dvmap is None:
            self.num_design_vars = self.nnodes
            self._dv_d = self._act_d = None
        else:
            dev = D.dev()
            self.num_design_vars = int(num_design_vars if num_design_vars is not None else self.dvmap.max() + 1)
            act = self.dvmap >= 0
            self._dv_d = torch.as_tensor(np.where(act, self.dvmap, 0), device=dev)
            self._act_d = torch.as_tensor(act, device=dev)
            self._act_idx = torch.as_tensor(np.nonzero(act)[0], device=dev)
            self._act_dv = torch.as_tensor(self.dvmap[act], device=dev)

    def _expand_design(self, x_d):
        if self.dvmap is None:
            return x_d
        xn = x_d.index_select(0, self._dv_d)                     # x[dvmap], and x := 1 where dvmap < 0 (:166-168)
        return torch.where(self._act_d, xn, torch.ones_like(xn))

    def _filtered(self, x_d):
        if self.ftype == "helmholtz":
            return self._hA.solve_dev(self._hB.spmm(self._expand_design(x_d)))
        rho = self.F_d.spmm(x_d)
        if self.offset_d is not None:
            D.axpby(1.0, rho, 1.0, self.offset_d, out=rho)
        return rho

    def _project(self, rho, g=None):
        out = D.empty(rho.shape[0])
        _lib.check(_lib.load().eigd_filter_project(rho.shape[0], self.beta, self.eta, _ptr(rho), _ptr(g), _ptr(out)),
                   "filter_project")
        return out

    def apply_tangent(self, x, xt):
        """Directional derivative of ``apply`` at x along xt (dual-number mode of the drivers)."""
        x_d, xt_d = to_dev(x), to_dev(xt)
        if self.ftype == "helmholtz":
            xt_n = xt_d if self.dvmap is None else torch.where(self._act_d, xt_d.index_select(0, self._dv_d), torch.zeros_like(self._dv_d, dtype=xt_d.dtype))
            rt = self._hA.solve_dev(self._hB.spmm(xt_n))
        else:
            rt = self.F_d.spmm(xt_d)                       # the constant offset of the dv-map has no derivative
        if self.projection:
            rt = self._project(self._filtered(x_d), _contig1(rt))
        return rt

    def apply(self, x):
        rho = self._filtered(to_dev(x))
        if self.projection:                                    # node_filter.py:174-181
            rho = self._project(rho)
        return like_input(rho, x)

    def apply_gradient(self, g, x=None, rho=None):
        g_d = to_dev(g)
        if self.projection:                                    # node_filter.py:187-203: needs the unprojected field
            if x is None:
                raise ValueError("apply_gradient with projection needs the design x")
            g_d = self._project(self._filtered(to_dev(x)), _contig1(g_d))
        if self.ftype == "helmholtz":                          # g0 = B^T A^-1 grad (:205-209), then the dv-map scatter (:211-214)
            g0 = self._hB.spmm(self._hA.solve_dev(_contig1(g_d)))
            if self.dvmap is not None:
                out = D.zeros(self.num_design_vars)
                out.index_add_(0, self._act_dv, g0.index_select(0, self._act_idx))
                g0 = out
            return like_input(g0, g)
        return like_input(self.FT_d.spmm(g_d), g)
