"""Integer structures of the symbolic analysis (csrc/symbolic.cpp), BIT-EXACT against the independent numpy restatement
oracle/multifrontal_oracle.py (north_star: "bit-exact for symbolic and indexing structures").  No GPU needed: the symbolic
phase is host C++ behind the C-ABI (eigd_symbolic_create / eigd_symbolic_get).

Checked for geometric and graph nested dissection, 1 and 2 dofs per node:
  * the permutation is a permutation; supernodes tile the columns; levels cover every supernode once, parents above children;
  * the elimination tree of the permuted pattern and the strict-lower column counts (etree_and_counts);
  * the below-diagonal row list of every front and the supernodal parent (front_structures);
  * the relative indices (child row -> position in the parent front) against their definition;
  * nnz(L) and the flop count derived from the counts;
  * the CSR -> front-slot assembly map against its definition."""
import numpy as np
import pytest
import scipy.sparse as sp

from eigd_b200 import fe
from eigd_b200.device import Symbolic
import multifrontal_oracle as mo


def build(nx, ny, dof, use_coords):
    conn, X = fe.grid_mesh(nx, ny, 1.0, 0.7)
    var = fe.element_dofs(conn, dof)
    ndof = dof * X.shape[0]
    indptr, indices, _, _ = fe.assembly_structure(var, ndof)
    A = sp.csr_matrix((np.ones(len(indices)), indices, indptr), shape=(ndof, ndof))
    sym = Symbolic(indptr, indices, ndof, coords=X if use_coords else None, dof_per_node=dof)
    return A, sym


@pytest.mark.parametrize("nx,ny,dof,use_coords", [(21, 16, 1, True), (13, 10, 2, True), (17, 12, 1, False), (9, 11, 2, False)])
def test_symbolic_arrays_bit_exact_vs_numpy_restatement(nx, ny, dof, use_coords):
    A, sym = build(nx, ny, dof, use_coords)
    n = A.shape[0]
    arr = sym.arrays()
    perm = arr["perm"]
    assert np.array_equal(np.sort(perm), np.arange(n))
    snf = arr["sn_first"]
    assert snf[0] == 0 and snf[-1] == n and np.all(np.diff(snf) > 0)
    ns = len(snf) - 1
    # elimination tree and column counts of P A P^T
    parent_o, count_o = mo.etree_and_counts(A, perm)
    assert np.array_equal(arr["parent"], parent_o)
    assert np.array_equal(arr["colcount"], count_o)
    # supernodal structures
    rows_o, snpar_o = mo.front_structures(A, perm, snf, parent_o)
    assert np.array_equal(arr["sn_parent"], snpar_o)
    rp = arr["sn_rowptr"]
    for k in range(ns):
        assert np.array_equal(arr["sn_rows"][rp[k]:rp[k + 1]], rows_o[k]), k
    # relaxed supernodes (collapsed leaf subtrees, amalgamation): the dense front contains every column's true structure
    for k in range(ns):
        cols = np.arange(snf[k], snf[k + 1])
        assert np.all(count_o[cols] <= (len(rows_o[k]) + (snf[k + 1] - 1 - cols)))
    # relative indices: row i of child c sits at position rel in the parent's front (pivot columns first, then rows)
    rel = arr["rel"]
    for k in range(ns):
        p = arr["sn_parent"][k]
        if p < 0:
            continue
        prow = np.concatenate([np.arange(snf[p], snf[p + 1]), rows_o[p]])
        assert np.array_equal(prow[rel[rp[k]:rp[k + 1]]], rows_o[k]), k
    # levels: every supernode once, parents strictly above their children
    assert np.array_equal(np.sort(arr["level_sn"]), np.arange(ns))
    kids = arr["sn_parent"] >= 0
    assert np.all(arr["sn_level"][arr["sn_parent"][kids]] > arr["sn_level"][kids])
    # totals
    assert sym.query("exact_nnzL") == int(count_o.sum())
    f = np.array([snf[k + 1] - snf[k] + len(rows_o[k]) for k in range(ns)])
    nc = np.diff(snf)
    assert sym.query("nnzL") == int(sum(nc[k] * f[k] - nc[k] * (nc[k] + 1) // 2 for k in range(ns)))
    assert sym.query("sum_front") == int(f.sum()) and sym.query("max_front") == int(f.max())
    # assembly map: CSR non-zero (r, c) with P r >= P c -> front_off[k] + local row + local col * f
    iperm = np.empty(n, dtype=np.int64)
    iperm[perm] = np.arange(n)
    amap = sym.assembly_map_host()
    col2sn = np.repeat(np.arange(ns), nc)
    front_off = arr["front_off"]
    Ac = sp.csr_matrix(A)
    for r in range(0, n, max(1, n // 97)):
        for pidx in range(Ac.indptr[r], Ac.indptr[r + 1]):
            pr, pc = iperm[r], iperm[Ac.indices[pidx]]
            if pr < pc:
                assert amap[pidx] == -1
                continue
            k = col2sn[pc]
            prow = np.concatenate([np.arange(snf[k], snf[k + 1]), rows_o[k]])
            lr = int(np.nonzero(prow == pr)[0][0])
            assert amap[pidx] == front_off[k] + lr + (pc - snf[k]) * f[k]
