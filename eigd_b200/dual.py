"""Complex-step operands on the device as (value, tangent) pairs.

The reference's examples verify gradients with the complex-step method (examples/thermal.py:652-661,
examples/buckling.py test_eigenvector_aggregate_derivatives): the design is perturbed by ``i h p`` with
h = 1e-30, K and M become complex, ``splu`` factors a complex matrix and ``BasicLanczos`` runs in complex
arithmetic WITHOUT conjugation, i.e. the imaginary parts are propagated as forward derivatives
(eigd/eigenvector_derivatives.py:1387-1414 says so explicitly for the reduced eigenproblem).  Here the same
first-order propagation is carried through the real fp64 kernels as dual numbers: a complex operand
``a + i b`` is the pair (a, b), products keep the first-order terms only (exact, whereas complex arithmetic
with h = 1e-30 drops ``b1 b2`` by underflow), the shifted matrix is factorised once in real arithmetic and a
dual solve is two real solves:

    (A + i dA)(x + i dx) = b + i db   ->   A x = b,   A dx = db - dA x.

Results are returned as complex numpy arrays (value + i tangent), which is what the reference returns.
"""
import numpy as np
import scipy.sparse as sp

from . import device as D
from ._hostdev import small_to_dev, to_dev, to_host


def is_complex_matrix(A):
    return isinstance(getattr(A, "data", None), np.ndarray) and hasattr(A, "indptr") and np.iscomplexobj(A.data)


def split_csr(A, symmetric=False):
    """scipy complex CSR/CSC -> (CsrDevice of the real parts, CsrDevice of the imaginary parts), one pattern."""
    if not (sp.issparse(A) and (A.format == "csr" or (symmetric and A.format == "csc"))):
        A = sp.csr_matrix(A)
    if not A.has_sorted_indices:
        A = A.sorted_indices()
    mk = sp.csc_matrix if A.format == "csc" else sp.csr_matrix
    Ar = mk((np.ascontiguousarray(A.data.real), A.indices, A.indptr), shape=A.shape)
    At = mk((np.ascontiguousarray(A.data.imag), A.indices, A.indptr), shape=A.shape)
    return D.CsrDevice.from_scipy(Ar, symmetric=symmetric), D.CsrDevice.from_scipy(At, symmetric=symmetric)


class Dual:
    """Host dual scalar."""
    __slots__ = ("r", "t")

    def __init__(self, r, t=0.0):
        self.r, self.t = float(r), float(t)

    def __mul__(self, o):
        o = o if isinstance(o, Dual) else Dual(o)
        return Dual(self.r * o.r, self.r * o.t + self.t * o.r)

    def inv(self):
        return Dual(1.0 / self.r, -self.t / (self.r * self.r))

    def sqrt(self):
        s = np.sqrt(self.r)
        return Dual(s, self.t / (2.0 * s))


class DualVec:
    """Device dual vector (value, tangent)."""
    __slots__ = ("r", "t")

    def __init__(self, r, t):
        self.r, self.t = r, t

    def copy_from(self, o):
        self.r.copy_(o.r)
        self.t.copy_(o.t)

    def axpy(self, h, v):
        """self -= h * v   (dual scalar h, dual vector v)"""
        D.axpby(1.0, self.r, -h.r, v.r, out=self.r)
        D.axpby(1.0, self.t, -h.r, v.t, out=self.t)
        if h.t != 0.0:
            D.axpby(1.0, self.t, -h.t, v.r, out=self.t)

    def scale(self, s):
        """self *= s"""
        D.axpby(s.r, self.t, s.t, self.r, out=self.t)
        D.axpby(s.r, self.r, 0.0, self.r, out=self.r)


def _dot(x, y):
    return float(to_host(D.col_dot(x, y))[0])


def dot(w, bv):
    """bilinear (no conjugation) product w . bv of two dual vectors"""
    return Dual(_dot(w.r, bv.r), _dot(w.t, bv.r) + _dot(w.r, bv.t))


def spmv(Br, Bt, x, out):
    Br.spmm(x.r, out=out.r)
    Br.spmm(x.t, out=out.t)
    Bt.spmm(x.r, out=out.t, alpha=1.0, beta=1.0)
    return out


def solve(factor, b, out):
    """out = (mat + i mat_t)^{-1} b to first order; ``factor`` is a SpLuOperator built from a complex matrix."""
    factor.solve_dev(b.r, out=out.r)
    rhs = factor.tangent.spmm(out.r)
    D.axpby(1.0, b.t, -1.0, rhs, out=rhs)
    factor.count -= 1                           # the reference counts one (complex) solve
    factor.solve_dev(rhs, out=out.t)
    return out


def eigh_dual(Tr, Tt):
    """Reference ``_eigh`` (:1387-1414): eigen-decomposition of T + i dT with dT as a forward derivative."""
    lam, Q = np.linalg.eigh(Tr)
    Dm = Q.T @ Tt @ Q
    lam_t = np.diag(Dm).copy()
    diff = lam[None, :] - lam[:, None]                      # lam[j] - lam[i]
    with np.errstate(divide="ignore", invalid="ignore"):
        C = np.where(diff != 0.0, Dm / diff, 0.0)
    np.fill_diagonal(C, 0.0)
    return lam, lam_t, Q, Q @ C


def basic_lanczos(Ar, At, Br, Bt, factor, sigma, N, m_max, tol, Ntarget, mode):
    """Dual-number form of ``BasicLanczos.solve`` with full orthogonalisation (reference :1496-1551, 1607-1650)."""
    n = Br.shape[0]
    dev_empty = lambda: DualVec(D.empty(n), D.zeros(n))
    V = [dev_empty() for _ in range(m_max + 1)]
    BV = [dev_empty() for _ in range(m_max + 1)]
    v0 = np.random.default_rng(12345).uniform(size=n, low=-1.0, high=1.0)              # :1514-1515
    V[0].r.copy_(to_dev(v0))
    V[0].t.zero_()
    spmv(Br, Bt, V[0], BV[0])
    s = dot(V[0], BV[0]).sqrt().inv()
    V[0].scale(s)
    BV[0].scale(s)
    alpha = [Dual(0.0) for _ in range(m_max)]
    beta = [Dual(0.0) for _ in range(m_max)]
    w, bw = dev_empty(), dev_empty()
    Nc = N if Ntarget is None else Ntarget
    m = m_max

    def reduced(mm):
        Tr, Tt = np.zeros((mm, mm)), np.zeros((mm, mm))
        for i in range(mm):
            Tr[i, i], Tt[i, i] = alpha[i].r, alpha[i].t
        for i in range(mm - 1):
            Tr[i, i + 1] = Tr[i + 1, i] = beta[i].r
            Tt[i, i + 1] = Tt[i + 1, i] = beta[i].t
        th, th_t, Y, Y_t = eigh_dual(Tr, Tt)
        if mode == "normal":                                                            # :1432-1437
            lam, lam_t = 1.0 / th + sigma, -th_t / th**2
            idx = np.argsort(lam)
        else:
            lam = sigma * th / (th - 1.0)
            lam_t = -sigma * th_t / (th - 1.0) ** 2
            idx = np.argsort(-1.0 / lam)
        return th, th_t, Y, Y_t, Tr, Tt, lam, lam_t, idx

    for i in range(1, m_max + 1):
        solve(factor, BV[i - 1], w)                                                     # :1522
        if i > 1:
            w.axpy(beta[i - 2], V[i - 2])
        for j in range(i - 1, -1, -1):                                                  # :1526-1532
            h = dot(w, BV[j])
            w.axpy(h, V[j])
            if j == i - 1:
                alpha[i - 1] = h
        spmv(Br, Bt, w, bw)
        beta[i - 1] = dot(w, bw).sqrt()
        s = beta[i - 1].inv()
        V[i].copy_from(w)
        BV[i].copy_from(bw)
        V[i].scale(s)
        BV[i].scale(s)
        if i >= 2:
            th, th_t, Y, Y_t, Tr, Tt, lam, lam_t, idx = reduced(i)
            err = np.abs(beta[i - 1].r * Y[i - 1, idx])
            bad = np.nonzero(err >= tol)[0]
            if (len(err) if len(bad) == 0 else bad[0]) >= Nc:
                m = i
                break
    th, th_t, Y, Y_t, Tr, Tt, lam, lam_t, idx = reduced(m)
    # Phi = V[:, :m] Y0 with dual V and dual Y0
    Vr = D.empty(m, n)
    Vt = D.empty(m, n)
    for j in range(m):
        Vr[j].copy_(V[j].r)
        Vt[j].copy_(V[j].t)
    return dict(m=m, alpha=np.array([a.r + 1j * a.t for a in alpha]), beta=np.array([b.r + 1j * b.t for b in beta]),
                theta=th + 1j * th_t, Y=Y + 1j * Y_t, T=Tr + 1j * Tt, lam=lam + 1j * lam_t, indices=idx, Vr=Vr, Vt=Vt,
                V=V, Y_r=Y, Y_t=Y_t)


def ritz_vectors(res, cols):
    """(n, len(cols)) complex Ritz vectors V Y[:, cols] of a dual Lanczos run."""
    Vr, Vt = res["Vr"], res["Vt"]
    Yr = np.ascontiguousarray(res["Y_r"][:, cols])
    Yt = np.ascontiguousarray(res["Y_t"][:, cols])
    n = Vr.shape[1]
    Pr, Pt = D.empty(n, len(cols)), D.empty(n, len(cols))
    D.gemm_nn(Vr.T, small_to_dev(Yr), Pr, alpha=1.0, beta=0.0)
    D.gemm_nn(Vt.T, small_to_dev(Yr), Pt, alpha=1.0, beta=0.0)
    D.gemm_nn(Vr.T, small_to_dev(Yt), Pt, alpha=1.0, beta=1.0)
    return to_host(Pr) + 1j * to_host(Pt)
