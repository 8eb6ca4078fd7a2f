"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the finite-element side of the reference
examples (meshes, Q4 element matrices, sensitivities, conic node filter).

Follows (paths relative to the reference root):
  mesh / connectivity      examples/thermal.py:1475-1498, examples/natural_frequency.py:850-894
  shape functions, B, detJ examples/fe_utils.py:4-16, 19-55, 124-156
  thermal K, dK, M, dM     examples/thermal.py:126-148, 150-190, 192-214, 216-246
  plane stress K, dK, M, dM examples/natural_frequency.py:134-160, 162-203, 205-236, 238-284
  node filter              examples/node_filter.py:61-88 (construction), 164-217 (apply / gradient)
  element -> node scatter  examples/thermal.py:612-615

It is the checker for the CUDA element kernels and the source of synthetic K, M for the
benchmarks; only tests/, __graft_entry__.smoke() and bench.py's setup / cpu_baseline use it.
Pinned against the real reference by tests/golden/make_golden.py (fixtures in tests/golden).
"""
import numpy as np
from scipy import sparse, spatial

GP = 1.0 / np.sqrt(3.0)


def grid_mesh(nx, ny, Lx=1.0, Ly=1.0):
    """nodes(i, j) = i*(ny+1) + j, element e = i + nx*j, counter-clockwise connectivity."""
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


def shape_derivs(xi, eta):
    N = 0.25 * np.array([(1 - xi) * (1 - eta), (1 + xi) * (1 - eta), (1 + xi) * (1 + eta), (1 - xi) * (1 + eta)])
    Nxi = 0.25 * np.array([-(1 - eta), (1 - eta), (1 + eta), -(1 + eta)])
    Neta = 0.25 * np.array([-(1 - xi), -(1 + xi), (1 + xi), (1 - xi)])
    return N, Nxi, Neta


def q4_geometry(conn, X):
    """Per element and Gauss point: N (4,4), Nx, Ny (nelems,4gp,4), detJ (nelems,4gp)."""
    xe, ye = X[conn, 0], X[conn, 1]
    ne = conn.shape[0]
    Nq = np.zeros((4, 4))
    Nx = np.zeros((ne, 4, 4))
    Ny = np.zeros((ne, 4, 4))
    detJ = np.zeros((ne, 4))
    for q in range(4):
        xi = GP if (q & 1) else -GP
        eta = GP if (q & 2) else -GP
        N, Nxi, Neta = shape_derivs(xi, eta)
        J00, J10, J01, J11 = xe @ Nxi, ye @ Nxi, xe @ Neta, ye @ Neta
        det = J00 * J11 - J01 * J10
        i00, i01, i10, i11 = J11 / det, -J01 / det, -J10 / det, J00 / det
        Nx[:, q, :] = np.outer(i00, Nxi) + np.outer(i10, Neta)
        Ny[:, q, :] = np.outer(i01, Nxi) + np.outer(i11, Neta)
        Nq[q] = N
        detJ[:, q] = det
    return Nq, Nx, Ny, detJ


def element_dofs(conn, dof):
    return (conn[:, :, None] * dof + np.arange(dof)[None, None, :]).reshape(conn.shape[0], -1)


def coo_index(var):
    """(i, j) lists in (element, local row, local col) order == Ke.flatten() order."""
    ne = var.shape[1]
    i = np.repeat(var, ne, axis=1).ravel()
    j = np.tile(var, (1, ne)).ravel()
    return i, j


class Q4Model:
    """kind = 'thermal' (1 dof/node) or 'plane_stress' (2 dof/node)."""

    def __init__(self, conn, X, kind="thermal", E=1.0, nu=0.3, kappa=1.0, density=1.0, heat_capacity=1.0,
                 p=3.0, beta=1e-6, rho0_K=1e-6):
        self.conn, self.X, self.kind = np.asarray(conn), np.asarray(X, dtype=float), kind
        self.nelems = conn.shape[0]
        self.nnodes = int(conn.max()) + 1
        self.dof = 1 if kind == "thermal" else 2
        self.ndof = self.dof * self.nnodes
        self.var = element_dofs(self.conn, self.dof)
        self.i, self.j = coo_index(self.var)
        self.Nq, self.Nx, self.Ny, self.detJ = q4_geometry(self.conn, self.X)
        self.kappa, self.density, self.heat_capacity = kappa, density, heat_capacity
        self.p, self.beta, self.rho0_K = p, beta, rho0_K
        C0 = E * np.array([[1.0, nu, 0.0], [nu, 1.0, 0.0], [0.0, 0.0, 0.5 * (1.0 - nu)]]) / (1.0 - nu**2)
        self.C0 = C0
        self.cmat6 = np.array([C0[0, 0], C0[0, 1], C0[0, 2], C0[1, 1], C0[1, 2], C0[2, 2]])

    # ---- material laws and their derivatives w.r.t. element density ---------------------
    def k_scale(self, rhoE):
        if self.kind == "thermal":
            return self.kappa * ((1 - self.beta) * rhoE**self.p + self.beta)
        return rhoE**self.p + self.rho0_K

    def k_scale_deriv(self, rhoE):
        if self.kind == "thermal":
            return (1 - self.beta) * self.kappa * self.p * rhoE ** (self.p - 1.0)
        return self.p * rhoE ** (self.p - 1.0)

    def m_scale(self, rhoE):
        if self.kind == "thermal":
            return self.heat_capacity * self.density * ((1 - self.beta) * rhoE + self.beta)
        return self.density * rhoE

    def m_scale_deriv(self, rhoE):
        if self.kind == "thermal":
            return (1 - self.beta) * self.heat_capacity * self.density * np.ones_like(rhoE)
        return self.density * np.ones_like(rhoE)

    def element_density(self, rho):
        return 0.25 * rho[self.conn].sum(axis=1)

    # ---- unit element matrices -------------------------------------------------------------
    def _B(self):
        """Strain-displacement rows per Gauss point: (nelems, 4gp, nstrain, ne)."""
        if self.kind == "thermal":
            return np.stack([self.Nx, self.Ny], axis=2)
        ne = self.nelems
        B = np.zeros((ne, 4, 3, 8))
        B[:, :, 0, 0::2] = self.Nx
        B[:, :, 1, 1::2] = self.Ny
        B[:, :, 2, 0::2] = self.Ny
        B[:, :, 2, 1::2] = self.Nx
        return B

    def _H(self):
        if self.kind == "thermal":
            return np.broadcast_to(self.Nq[None, :, None, :], (self.nelems, 4, 1, 4))
        H = np.zeros((self.nelems, 4, 2, 8))
        H[:, :, 0, 0::2] = self.Nq[None]
        H[:, :, 1, 1::2] = self.Nq[None]
        return H

    def unit_Ke(self):
        B = self._B()
        C = np.eye(2) if self.kind == "thermal" else self.C0
        return np.einsum("nq,nqia,ij,nqjb->nab", self.detJ, B, C, B)

    def unit_Me(self):
        H = self._H()
        return np.einsum("nq,nqia,nqib->nab", self.detJ, H, H)

    def assemble(self, rhoE):
        Ke = self.k_scale(rhoE)[:, None, None] * self.unit_Ke()
        Me = self.m_scale(rhoE)[:, None, None] * self.unit_Me()
        K = sparse.coo_matrix((Ke.ravel(), (self.i, self.j)), shape=(self.ndof, self.ndof)).tocsr()
        M = sparse.coo_matrix((Me.ravel(), (self.i, self.j)), shape=(self.ndof, self.ndof)).tocsr()
        return K, M

    # ---- sensitivities: sum over modes of w_e^T (d Ke / d rhoE) v_e ------------------------
    def dK(self, rhoE, W, V):
        B = self._B()
        C = np.eye(2) if self.kind == "thermal" else self.C0
        We, Ve = W[self.var, ...], V[self.var, ...]
        if W.ndim == 1:
            We, Ve = We[..., None], Ve[..., None]
        sw = np.einsum("nqia,nak->nqik", B, We)
        sv = np.einsum("nqia,nak->nqik", B, Ve)
        val = np.einsum("nq,ij,nqik,nqjk->n", self.detJ, C, sw, sv)
        return self.k_scale_deriv(rhoE) * val

    def dM(self, rhoE, W, V):
        H = self._H()
        We, Ve = W[self.var, ...], V[self.var, ...]
        if W.ndim == 1:
            We, Ve = We[..., None], Ve[..., None]
        hw = np.einsum("nqia,nak->nqik", H, We)
        hv = np.einsum("nqia,nak->nqik", H, Ve)
        val = np.einsum("nq,nqik,nqik->n", self.detJ, hw, hv)
        return self.m_scale_deriv(rhoE) * val

    def scatter_to_nodes(self, e_vals):
        out = np.zeros(self.nnodes)
        for a in range(4):
            np.add.at(out, self.conn[:, a], e_vals)
        return 0.25 * out

    def node_adjacency(self):
        """CSR node -> elements (for the gather-form scatter kernel)."""
        e = np.repeat(np.arange(self.nelems), 4)
        v = self.conn.ravel()
        order = np.lexsort((e, v))
        nptr = np.zeros(self.nnodes + 1, dtype=np.int32)
        np.add.at(nptr, v + 1, 1)
        return np.cumsum(nptr).astype(np.int32), e[order].astype(np.int32)

    def assembly_sources(self, K):
        """For every CSR non-zero of K: the (element, a, b) slots that sum into it."""
        Kc = sparse.csr_matrix(K)
        ne = self.var.shape[1]
        key = self.i.astype(np.int64) * self.ndof + self.j
        # position of each (i, j) in the sorted CSR
        rows = np.repeat(np.arange(Kc.shape[0]), np.diff(Kc.indptr))
        csr_key = rows.astype(np.int64) * self.ndof + Kc.indices
        pos = np.searchsorted(csr_key, key)
        assert (csr_key[pos] == key).all()
        order = np.argsort(pos, kind="stable")
        src = order.astype(np.int64)  # flat index e*ne*ne + a*ne + b
        src_ptr = np.zeros(Kc.nnz + 1, dtype=np.int64)
        np.add.at(src_ptr, pos + 1, 1)
        return np.cumsum(src_ptr), src, pos


class ConicFilter:
    """F[i, j] ~ max(0, r0 - |X_i - X_j|), rows normalised (examples/node_filter.py:61-88)."""

    def __init__(self, X, r0):
        tree = spatial.cKDTree(X)
        D = tree.sparse_distance_matrix(tree, r0, output_type="coo_matrix")
        n = X.shape[0]
        w = r0 - D.data
        # the COO output carries the zero-distance diagonal explicitly
        F = sparse.coo_matrix((w, (D.row, D.col)), shape=(n, n)).tocsr()
        rs = np.asarray(F.sum(axis=1)).ravel()
        self.F = sparse.diags(1.0 / rs) @ F
        self.F = self.F.tocsr()
        self.FT = self.F.T.tocsr()
        self.r0 = r0
        self.num_design_vars = n

    def apply(self, x):
        return self.F @ x

    def apply_gradient(self, g, x=None):
        return self.FT @ g
