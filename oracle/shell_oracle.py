"""TEST INFRASTRUCTURE ONLY -- host (scipy) counterpart of eigd_b200/shell.py.

The reference's CRM example takes K, M and the design sensitivities from TACS (absent), so there
is no reference implementation of the shell element to restate.  What this module pins instead is everything the
device does with the element matrices: it assembles K(t) = sum_e P_e^T (t_e E1_e + t_e^3 E3_e) P_e and M(t) with
scipy's COO -> CSR (the way the reference's 2-D examples assemble, examples/natural_frequency.py:157-158), applies
the clamped-node reduction of examples/crm.py:143-181, and evaluates the per-component sensitivities
w^T (dK/dx_c) v, w^T (dM/dx_c) v with numpy einsum -- the quantities crm.py:331-355 obtains from
``addMatDVSensInnerProduct``.  tests/golden/make_golden.py feeds these host matrices to the UNMODIFIED reference
solvers (IRAM + sibk + add_eig_total_derivative, per-mode "vector" form) to freeze the C4-shaped fixture.
The unit element matrices themselves (geometry only) come from eigd_b200.shell.shell_unit_matrices and are checked
here by their invariants: symmetry, six rigid-body modes of the free structure, positive definite mass."""
import numpy as np
import scipy.sparse as sp


class ShellOracle:
    def __init__(self, conn, X, comp, fixed_nodes, E1, E3, F1, F3, scale=100.0):
        self.conn = np.asarray(conn, dtype=np.int64)
        self.nelems, self.nnodes = self.conn.shape[0], np.asarray(X).shape[0]
        self.comp = np.asarray(comp, dtype=np.int64)
        self.ncomp = int(self.comp.max()) + 1
        self.scale = scale
        self.E1, self.E3, self.F1, self.F3 = E1, E3, F1, F3
        self.var = (self.conn[:, :, None] * 6 + np.arange(6)[None, None, :]).reshape(self.nelems, 24)
        fixed = np.zeros(self.nnodes, dtype=bool)
        fixed[np.asarray(fixed_nodes, dtype=np.int64)] = True
        self.reduced = (np.nonzero(~fixed)[0][:, None] * 6 + np.arange(6)[None, :]).ravel()
        self.i = np.repeat(self.var, 24, axis=1).ravel()
        self.j = np.tile(self.var, (1, 24)).ravel()
        self.ndof_full = 6 * self.nnodes

    def thickness(self, x):
        return (np.asarray(x, dtype=float) / self.scale)[self.comp]

    def assemble(self, x):
        t = self.thickness(x)
        Ke = t[:, None, None] * self.E1 + (t ** 3)[:, None, None] * self.E3
        Me = t[:, None, None] * self.F1 + (t ** 3)[:, None, None] * self.F3
        n = self.ndof_full
        K = sp.coo_matrix((Ke.ravel(), (self.i, self.j)), shape=(n, n)).tocsr()
        M = sp.coo_matrix((Me.ravel(), (self.i, self.j)), shape=(n, n)).tocsr()
        r = self.reduced
        Kr, Mr = K[r, :][:, r].tocsr(), M[r, :][:, r].tocsr()
        Kr.sort_indices()
        Mr.sort_indices()
        return Kr, Mr

    def _sens(self, U1, U3, x, wr, vr):
        t = self.thickness(x)
        w = np.zeros(self.ndof_full)
        v = np.zeros(self.ndof_full)
        w[self.reduced], v[self.reduced] = wr, vr
        we, ve = w[self.var], v[self.var]
        q = np.einsum("ea,eab,eb->e", we, U1, ve) + 3.0 * t ** 2 * np.einsum("ea,eab,eb->e", we, U3, ve)
        out = np.zeros(self.ncomp)
        np.add.at(out, self.comp, q)
        return out / self.scale

    def dK(self, x, wr, vr):
        return self._sens(self.E1, self.E3, x, wr, vr)

    def dM(self, x, wr, vr):
        return self._sens(self.F1, self.F3, x, wr, vr)
