"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the supernodal multifrontal LDL^T.

There is no reference source for this stage: the reference delegates to SuperLU through
``scipy.sparse.linalg.splu`` (reference eigd/eigenvector_derivatives.py:11-23).  What is
pinned here is (i) the symbolic invariants of ``eigd_b200/csrc/symbolic.cpp`` (elimination
tree, column counts, front structures, relative indices, assembly map), restated with plain
numpy/python, and (ii) the front-by-front arithmetic the CUDA kernels in
``eigd_b200/csrc/factor.cu`` perform, so that small cases can be compared front by front.
The solve is checked against ``splu`` of the same matrix in the tests.

Only tests/ and __graft_entry__.smoke() import this module.
"""
import numpy as np
import scipy.sparse as sp


# ---------------------------------------------------------------------------------------
# symbolic restatement (independent of the C++ code: operates on the permuted pattern)
# ---------------------------------------------------------------------------------------
def etree_and_counts(A, perm):
    """Elimination tree and strict-lower column counts of P A P^T (perm[new] = old)."""
    n = A.shape[0]
    iperm = np.empty(n, dtype=np.int64)
    iperm[perm] = np.arange(n)
    C = sp.coo_matrix(A)
    r, c = iperm[C.row], iperm[C.col]
    keep = r > c
    lo = sp.csr_matrix((np.ones(keep.sum()), (r[keep], c[keep])), shape=(n, n))
    lo.sum_duplicates()
    parent = -np.ones(n, dtype=np.int64)
    anc = -np.ones(n, dtype=np.int64)
    for i in range(n):
        for j in lo.indices[lo.indptr[i]:lo.indptr[i + 1]]:
            while j != -1 and j < i:
                nx = anc[j]
                anc[j] = i
                if nx == -1:
                    parent[j] = i
                j = nx
    count = np.zeros(n, dtype=np.int64)
    mark = -np.ones(n, dtype=np.int64)
    for i in range(n):
        mark[i] = i
        for k in lo.indices[lo.indptr[i]:lo.indptr[i + 1]]:
            while k != -1 and mark[k] != i:
                count[k] += 1
                mark[k] = i
                k = parent[k]
    return parent, count


def front_structures(A, perm, sn_first, parent):
    """Below-diagonal row lists of each front for a given supernode partition."""
    n = A.shape[0]
    iperm = np.empty(n, dtype=np.int64)
    iperm[perm] = np.arange(n)
    C = sp.coo_matrix(A)
    r, c = iperm[C.row], iperm[C.col]
    hi, lo = np.maximum(r, c), np.minimum(r, c)
    keep = hi > lo
    bycol = sp.csc_matrix((np.ones(keep.sum()), (hi[keep], lo[keep])), shape=(n, n))
    bycol.sum_duplicates()
    ns = len(sn_first) - 1
    col2sn = np.repeat(np.arange(ns), np.diff(sn_first))
    sn_parent = np.array([-1 if parent[sn_first[k + 1] - 1] < 0 else col2sn[parent[sn_first[k + 1] - 1]]
                          for k in range(ns)])
    rows = [None] * ns
    children = [[] for _ in range(ns)]
    for k in range(ns):
        if sn_parent[k] >= 0:
            children[sn_parent[k]].append(k)
    for k in range(ns):
        last = sn_first[k + 1] - 1
        s = set()
        for cc in range(sn_first[k], last + 1):
            s.update(int(x) for x in bycol.indices[bycol.indptr[cc]:bycol.indptr[cc + 1]] if x > last)
        for ch in children[k]:
            s.update(int(x) for x in rows[ch] if x > last)
        rows[k] = np.array(sorted(s), dtype=np.int64)
    return rows, sn_parent


# ---------------------------------------------------------------------------------------
# numeric restatement, driven by the arrays the C++ symbolic phase exports
# ---------------------------------------------------------------------------------------
class MultifrontalOracle:
    """Front-by-front LDL^T (no pivoting, static perturbation) + solves, in numpy.

    sym: dict with perm, sn_first, sn_rowptr, sn_rows, sn_parent, front_off, rel, level_ptr,
    level_sn (int64 arrays as returned by eigd_symbolic_get) and amap (assembly map).
    """

    def __init__(self, sym):
        self.sym = sym
        self.ns = len(sym["sn_first"]) - 1

    def _front_view(self, F, k):
        s = self.sym
        nc = s["sn_first"][k + 1] - s["sn_first"][k]
        nb = s["sn_rowptr"][k + 1] - s["sn_rowptr"][k]
        f = nc + nb
        o = s["front_off"][k]
        return F[o:o + f * f].reshape(f, f).T, nc, nb  # column-major storage -> view[i, j]

    def factor(self, vals, piv_tol=1e-11):
        s = self.sym
        F = np.zeros(s["front_off"][-1])
        amap = s["amap"]
        m = amap >= 0
        F[amap[m]] = vals[m]
        amax = np.abs(vals).max()
        thr = piv_tol * amax
        nneg = nper = 0
        dinv = np.zeros(s["sn_first"][-1])
        for lvl in range(len(s["level_ptr"]) - 1):
            for k in s["level_sn"][s["level_ptr"][lvl]:s["level_ptr"][lvl + 1]]:
                Fk, nc, nb = self._front_view(F, k)
                # extend-add of the children's contribution blocks (children are at lower levels)
                for ch in np.nonzero(s["sn_parent"] == k)[0]:
                    Fc, ncc, nbc = self._front_view(F, ch)
                    rel = s["rel"][s["sn_rowptr"][ch]:s["sn_rowptr"][ch + 1]]
                    C = np.tril(Fc[ncc:, ncc:])
                    Fk[np.ix_(rel, rel)] += C
                # dense partial LDL^T on the lower triangle
                for j in range(nc):
                    d = Fk[j, j]
                    if abs(d) < thr:
                        d = thr if d >= 0 else -thr
                        nper += 1
                        Fk[j, j] = d
                    if d < 0:
                        nneg += 1
                    dinv[s["sn_first"][k] + j] = 1.0 / d
                    col = Fk[j + 1:, j].copy()
                    Fk[j + 1:, j] = col / d
                    Fk[j + 1:, j + 1:] -= np.tril(np.outer(col, col / d))
        self.F, self.dinv, self.info = F, dinv, (nneg, nper)
        return self

    def solve(self, B):
        s = self.sym
        B = np.asarray(B, dtype=float)
        squeeze = B.ndim == 1
        if squeeze:
            B = B[:, None]
        perm = s["perm"]
        y = B[perm].copy()
        # forward, fan-out form (the CUDA code passes contributions up the tree instead;
        # the sums are the same)
        for k in range(self.ns):
            Fk, nc, nb = self._front_view(self.F, k)
            c0 = s["sn_first"][k]
            L11 = np.tril(Fk[:nc, :nc], -1) + np.eye(nc)
            y[c0:c0 + nc] = np.linalg.solve(L11, y[c0:c0 + nc])
            rows = s["sn_rows"][s["sn_rowptr"][k]:s["sn_rowptr"][k + 1]]
            y[rows] -= Fk[nc:, :nc] @ y[c0:c0 + nc]
        y *= self.dinv[:, None]
        for k in range(self.ns - 1, -1, -1):
            Fk, nc, nb = self._front_view(self.F, k)
            c0 = s["sn_first"][k]
            rows = s["sn_rows"][s["sn_rowptr"][k]:s["sn_rowptr"][k + 1]]
            t = y[c0:c0 + nc] - Fk[nc:, :nc].T @ y[rows]
            L11 = np.tril(Fk[:nc, :nc], -1) + np.eye(nc)
            y[c0:c0 + nc] = np.linalg.solve(L11.T, t)
        X = np.empty_like(y)
        X[perm] = y
        return X[:, 0] if squeeze else X
