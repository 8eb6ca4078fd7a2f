"""Device-resident restarted shift-and-invert Lanczos behind the reference's ``eigsh_mod``.

Mirrors the contract of reference ``eigd/arpack.py`` (``eigsh_mod`` at :104-118 returning
``(d, z, Tm, v)`` as extracted at :58-101): ``d`` the k wanted eigenvalues in ascending order,
``z`` their B-orthonormal eigenvectors (n, k), ``v`` the (n, ncv) B-orthonormal Krylov basis of
the final factorisation and ``Tm = v^T B OP v`` its (ncv, ncv) projected operator.

The reference drives ARPACK's implicitly restarted Lanczos (dsaupd/dseupd) on the host, one
SuperLU solve and one SciPy SpMV per reverse-communication step.  Here the whole recurrence
lives in HBM: the basis V and B*V are stored one vector per row, OP = (A - sigma B)^{-1} B is
the multifrontal LDL^T solve plus a CSR SpMV, orthogonalisation is classical Gram-Schmidt with
an unconditional second pass (DGKS) done as two tall-skinny GEMV pairs, and the only host work
is the ncv x ncv symmetric eigenproblem once per restart cycle (one small D2H per cycle).
Restarting is thick restart (Wu & Simon), which is mathematically equivalent to ARPACK's
implicit restart with exact shifts; consequently ``Tm`` is diagonal-plus-arrow-plus-tridiagonal
instead of tridiagonal, which the callers (``IRAM.solve``: ``eigh(T)``; ``laa``) do not care
about.  Only the call pattern eigd uses is implemented (sigma given, OPinv given,
which="LM", mode "normal" or "buckling"); the remaining scipy modes raise NotImplementedError
because there is no CPU fallback to delegate to.
"""
import numpy as np
import torch

from . import device as D
from ._hostdev import as_csr_device, to_dev, to_host, is_dev, small_to_dev


class ArpackError(RuntimeError):
    """Same role as scipy.sparse.linalg.ArpackError."""


class ArpackNoConvergence(ArpackError):
    def __init__(self, msg, eigenvalues, eigenvectors):
        ArpackError.__init__(self, msg)
        self.eigenvalues = eigenvalues
        self.eigenvectors = eigenvectors


class LanczosState:
    """Device-side result of the restarted Lanczos process (kept by IRAM as its cache)."""

    def __init__(self):
        self.Vt = None       # (ncv, n) device, rows are B-orthonormal Krylov vectors
        self.T = None        # (ncv, ncv) numpy
        self.theta = None    # Ritz values of OP for the wanted set, in output order
        self.d = None        # eigenvalues of the pencil (numpy, k)
        self.Z = None        # (n, k) device eigenvectors
        self.nops = 0        # operator applications (solves)
        self.ncycles = 0
        self.resid = None


def _theta_to_lambda(theta, sigma, mode):
    if mode == "normal":
        return 1.0 / theta + sigma                      # eigd/eigenvector_derivatives.py:1960
    return sigma * theta / (theta - 1.0)                # :1963


_SEEDED_V0 = {}


def lanczos_thick_restart(Bip, factor, k, ncv, sigma, mode="normal", tol=0.0, maxiter=None, v0=None, seed=None):
    """Thick-restart Lanczos on OP = factor o Bip in the Bip inner product; returns LanczosState."""
    n = Bip.shape[0]
    if ncv > n:
        raise ValueError("ncv must be k<ncv<=n")
    if maxiter is None:
        maxiter = 10 * n
    eps = np.finfo(np.float64).eps
    if tol <= 0.0:
        tol = eps
    eps23 = eps ** (2.0 / 3.0)

    Vt = D.empty(ncv + 1, n)
    BVt = D.empty(ncv + 1, n)
    tmp = D.empty(ncv, n)
    hbuf = D.zeros(ncv + 1)
    h2 = D.zeros(ncv + 1)
    ab = D.zeros(2, ncv + 1)          # row 0: alpha_j, row 1: beta_j^2
    st = LanczosState()

    # start vector: forced into the range of OP as ARPACK's dgetv0 does
    if v0 is not None:
        w = to_dev(v0, copy=True).reshape(n)
    elif seed is None:
        w = to_dev(np.random.default_rng().uniform(-1.0, 1.0, n))
    else:
        # a seeded start vector is the same array every time: keep its device copy (drawing n host randoms and
        # uploading them costs about a millisecond at C2 during which the device idles)
        key = (int(seed), int(n), str(D.dev()))
        w = _SEEDED_V0.get(key)
        if w is None:
            if len(_SEEDED_V0) >= 4:
                _SEEDED_V0.clear()
            w = _SEEDED_V0[key] = to_dev(np.random.default_rng(seed).uniform(-1.0, 1.0, n))
    bw = Bip.spmm(w)
    v = factor.solve_dev(bw)
    st.nops += 1
    Bip.spmm(v, out=BVt[0])
    nrm2 = D.col_dot(v, BVt[0])
    Vt[0].copy_(v)
    D.col_scale(Vt[0], nrm2, mode=2)
    D.col_scale(BVt[0], nrm2, mode=2)

    T = np.zeros((ncv, ncv))
    j0 = 0
    w = D.empty(n)
    last_beta = 0.0
    best = np.inf
    stagnant = 0
    while True:
        # steps j0 .. ncv-1 of the recurrence in one native call (csrc/krylov.cu): no host round trip per step
        D.lanczos_extend(factor, Bip, Vt, BVt, j0, ncv, w, hbuf, h2, ab)
        st.nops += ncv - j0
        st.ncycles += 1
        abh = to_host(ab)                                    # the one D2H of the cycle
        if not np.all(np.isfinite(abh[:, j0:ncv])) or np.any(abh[1, j0:ncv] <= 0.0):
            raise ArpackError("Lanczos breakdown: non-finite or non-positive B-norm (is B positive definite "
                              "and the shifted matrix non-singular?)")
        alpha = abh[0]
        beta = np.sqrt(abh[1])
        for j in range(j0, ncv):
            T[j, j] = alpha[j]
            if j + 1 < ncv:
                T[j, j + 1] = T[j + 1, j] = beta[j]
        last_beta = beta[ncv - 1]
        theta, Y = np.linalg.eigh(T)
        order = np.argsort(-np.abs(theta))                   # which = "LM"
        bounds = np.abs(last_beta * Y[ncv - 1, :])
        wanted = order[:k]
        conv = bounds[wanted] <= tol * np.maximum(eps23, np.abs(theta[wanted]))
        nconv = int(conv.sum())
        worst = float(np.max(bounds[wanted] / np.maximum(eps23, np.abs(theta[wanted]))))
        # the estimates bottom out near eps*|theta| (absolute accuracy of the small eigenvectors):
        # accept when they stop improving at that level
        if worst < 0.5 * best:
            best, stagnant = worst, 0
        else:
            stagnant += 1
        done = nconv == k or (worst <= 64.0 * eps and stagnant >= 1)
        if done or st.nops >= maxiter or ncv >= n:
            break
        # ---- thick restart: keep the wanted Ritz vectors plus a share of the converged count --
        kk = min(k + min(nconv, (ncv - k) // 2), ncv - 1)
        if kk == 1 and ncv > 3:
            kk = 2
        keep = order[:kk]
        Yk = small_to_dev(Y[:, keep])
        D.gemm_nn(Vt[:ncv].T, Yk, tmp[:kk].T, alpha=1.0, beta=0.0)
        Vt[:kk].copy_(tmp[:kk])
        D.gemm_nn(BVt[:ncv].T, Yk, tmp[:kk].T, alpha=1.0, beta=0.0)
        BVt[:kk].copy_(tmp[:kk])
        Vt[kk].copy_(Vt[ncv])
        BVt[kk].copy_(BVt[ncv])
        T[:, :] = 0.0
        T[np.arange(kk), np.arange(kk)] = theta[keep]
        T[kk, :kk] = last_beta * Y[ncv - 1, keep]
        T[:kk, kk] = T[kk, :kk]
        j0 = kk

    if nconv < k and not done:
        lam = _theta_to_lambda(theta[wanted][conv], sigma, mode)
        raise ArpackNoConvergence("ARPACK error -1: No convergence (%d iterations, %d/%d eigenvectors converged)"
                                  % (st.nops, nconv, k), lam, None)
    lam = _theta_to_lambda(theta[wanted], sigma, mode)
    srt = np.argsort(lam)                                    # ascending algebraic order (eigd/arpack.py:237-239)
    sel = wanted[srt]
    st.d = lam[srt]
    st.theta = theta[sel]
    st.resid = bounds[sel]
    st.T = T.copy()
    st.eigh_T = (theta.copy(), Y.copy())     # eigh of the returned T and the Ritz columns behind Z, for the caller
    st.sel = sel.copy()
    st.Vt = Vt[:ncv]
    Z = D.empty(n, k)
    D.gemm_nn(st.Vt.T, small_to_dev(Y[:, sel]), Z, alpha=1.0, beta=0.0)
    st.Z = Z
    return st


BLOCK_SIZE = int(__import__("os").environ.get("EIGD_LANCZOS_BLOCK", "2"))


def lanczos_block_thick_restart(Bip, factor, k, ncv, sigma, P, mode="normal", tol=0.0, maxiter=None, v0=None, seed=None):
    """Block thick-restart Lanczos (block size P) on OP = factor o Bip in the Bip inner product; returns LanczosState.

    Same contract as ``lanczos_thick_restart`` (Vt: ncv B-orthonormal rows, T = V^T B OP V, Z = V Y[:, sel]); T is block
    tridiagonal with P x P blocks plus the restart arrow.  One native call per restart cycle (csrc/block_krylov.cu),
    one device -> host copy of the cycle's P x P blocks; the ncv x ncv ``eigh`` runs on the host as in the reference
    (eigd/eigenvector_derivatives.py:1958).  P right-hand sides per triangular solve cost little more than one on
    this hardware, and the number of operator applications stays about the same, so the eigensolve needs about 1/P of
    the sequential solves of the single-vector recurrence."""
    n = Bip.shape[0]
    if maxiter is None:
        maxiter = 10 * n
    eps = np.finfo(np.float64).eps
    if tol <= 0.0:
        tol = eps
    eps23 = eps ** (2.0 / 3.0)
    Vt = D.empty(ncv + P, n)
    BVt = D.empty(ncv + P, n)
    tmp = D.empty(ncv, n)
    nstep_max = (ncv + P - 1) // P + 1
    blocks = D.zeros(2, nstep_max, P, P)          # [0]: diagonal blocks A_j, [1]: sub-diagonal blocks R_j of the cycle
    H1, H2 = D.zeros((ncv + P) * P), D.zeros((ncv + P) * P)
    scratch = D.zeros(3 * P * P)
    st = LanczosState()
    # start block: P vectors forced into the range of OP (as ARPACK's dgetv0 does for its one vector)
    if v0 is not None:
        w0 = np.asarray(to_host(v0) if is_dev(v0) else v0, dtype=float).reshape(n)
        rest = np.random.default_rng(0 if seed is None else seed).uniform(-1.0, 1.0, (P - 1, n))
        W0 = to_dev(np.vstack([w0[None, :], rest]))
    elif seed is None:
        W0 = to_dev(np.random.default_rng().uniform(-1.0, 1.0, (P, n)))
    else:
        key = (int(seed), int(n), int(P), str(D.dev()))
        W0 = _SEEDED_V0.get(key)
        if W0 is None:
            if len(_SEEDED_V0) >= 4:
                _SEEDED_V0.clear()
            W0 = _SEEDED_V0[key] = to_dev(np.random.default_rng(seed).uniform(-1.0, 1.0, (P, n)))
    BW0 = Bip.spmm(W0.T)                                    # (n, P) row-major
    factor.solve_dev(BW0, out=Vt[:P].T)
    st.nops += P
    D.block_lanczos_start(Bip, Vt, BVt, P, scratch)
    T = np.zeros((ncv + P, ncv + P))
    if ncv % P:
        raise ValueError("block Lanczos needs ncv to be a multiple of the block size")
    j0, m0 = 0, P
    ncv_c = ncv
    best, stagnant = np.inf, 0
    while True:
        nsteps = (ncv_c - j0) // P
        D.block_lanczos_extend(factor, Bip, Vt, BVt, P, j0, m0, ncv_c, blocks[0], blocks[1], H1, H2, scratch)
        st.nops += ncv_c - j0
        st.ncycles += 1
        blk = to_host(blocks[:, :nsteps])                   # the one D2H of the cycle
        if not np.all(np.isfinite(blk)):
            raise ArpackError("Lanczos breakdown: rank-deficient Krylov block or non-positive B-norm (is B positive "
                              "definite and the shifted matrix non-singular?)")
        for s_ in range(nsteps):
            j, m = j0 + s_ * P, m0 + s_ * P
            A_j, R_j = blk[0, s_], blk[1, s_]
            T[j:j + P, j:j + P] = 0.5 * (A_j + A_j.T)
            T[m:m + P, j:j + P] = R_j
            T[j:j + P, m:m + P] = R_j.T
        Tm = T[:ncv_c, :ncv_c]
        R_last = T[ncv_c:ncv_c + P, ncv_c - P:ncv_c]
        theta, Y = np.linalg.eigh(Tm)
        order = np.argsort(-np.abs(theta))                  # which = "LM"
        bounds = np.linalg.norm(R_last @ Y[ncv_c - P:ncv_c, :], axis=0)
        wanted = order[:k]
        relb = bounds[wanted] / np.maximum(eps23, np.abs(theta[wanted]))
        conv = relb <= tol
        nconv = int(conv.sum())
        worst = float(relb.max())
        if worst < 0.5 * best:
            best, stagnant = worst, 0
        else:
            stagnant += 1
        # the estimates of a block recurrence bottom out at a few tens of eps * |theta| (rounding of the P x P block
        # factorisations): accept there instead of iterating on noise
        done = nconv == k or worst <= 64.0 * eps
        if done or st.nops >= maxiter or ncv_c >= n:
            break
        # ---- thick restart: keep the wanted Ritz vectors plus a share of the converged count; the kept count leaves
        # a whole number of blocks for the next cycle
        kk = max(min(k + min(nconv, (ncv - k) // 2), ncv - 2 * P), 1)
        kk += (ncv - kk) % P                                # ... so that the next cycle ends exactly at ncv
        if kk > ncv - P:
            kk -= P
        keep = order[:kk]
        Yk = small_to_dev(Y[:, keep])
        D.gemm_nn(Vt[:ncv_c].T, Yk, tmp[:kk].T, alpha=1.0, beta=0.0)
        res_V = Vt[ncv_c:ncv_c + P].clone()
        Vt[:kk].copy_(tmp[:kk])
        D.gemm_nn(BVt[:ncv_c].T, Yk, tmp[:kk].T, alpha=1.0, beta=0.0)
        res_BV = BVt[ncv_c:ncv_c + P].clone()
        BVt[:kk].copy_(tmp[:kk])
        Vt[kk:kk + P].copy_(res_V)
        BVt[kk:kk + P].copy_(res_BV)
        S = R_last @ Y[ncv_c - P:ncv_c, keep]               # P x kk coupling of the residual block to the kept Ritz vectors
        T[:, :] = 0.0
        T[np.arange(kk), np.arange(kk)] = theta[keep]
        T[kk:kk + P, :kk] = S
        T[:kk, kk:kk + P] = S.T
        j0, m0 = kk, kk + P

    if nconv < k and not done:
        lam = _theta_to_lambda(theta[wanted][conv], sigma, mode)
        raise ArpackNoConvergence("ARPACK error -1: No convergence (%d iterations, %d/%d eigenvectors converged)"
                                  % (st.nops, nconv, k), lam, None)
    lam = _theta_to_lambda(theta[wanted], sigma, mode)
    srt = np.argsort(lam)
    sel = wanted[srt]
    st.d = lam[srt]
    st.theta = theta[sel]
    st.resid = bounds[sel]
    st.T = Tm.copy()
    st.eigh_T = (theta.copy(), Y.copy())
    st.sel = sel.copy()
    st.Vt = Vt[:ncv_c]
    Z = D.empty(n, k)
    D.gemm_nn(st.Vt.T, small_to_dev(Y[:, sel]), Z, alpha=1.0, beta=0.0)
    st.Z = Z
    return st


def eigsh_mod(A, k=6, M=None, sigma=None, which="LM", v0=None, ncv=None, maxiter=None, tol=0,
              return_eigenvectors=True, Minv=None, OPinv=None, mode="normal", return_state=False, seed=None, block=None):
    """Signature of reference eigd/arpack.py:104-118 (scipy's eigsh plus the 4-tuple return).

    A : for mode="normal" unused beyond its shape; for mode="buckling" the matrix ARPACK
        multiplies by (eigd passes the stiffness matrix, eigd/eigenvector_derivatives.py:1941-1942).
    M : inner-product matrix of mode "normal".   OPinv : ``SpLuOperator`` of the shifted matrix.
    Returns (d, z, Tm, v) as numpy arrays; with return_state=True the device-side LanczosState
    is appended so callers can keep the basis in HBM.
    """
    n = A.shape[0]
    if A.shape[0] != A.shape[1]:
        raise ValueError(f"expected square matrix (shape={A.shape})")
    if k <= 0:
        raise ValueError("k must be greater than 0.")
    if k >= n:
        raise NotImplementedError("k >= n needs a dense eigensolver; not part of the device path")
    if sigma is None or OPinv is None:
        raise NotImplementedError("eigd_b200.eigsh_mod implements the call pattern eigd uses: sigma and OPinv given "
                                  "(shift-invert); other scipy modes have no device implementation")
    if which != "LM":
        raise NotImplementedError("only which='LM' is implemented (the value eigd passes)")
    if Minv is not None:
        raise NotImplementedError("Minv is not used in shift-invert mode")
    if mode == "normal":
        if M is None:
            raise NotImplementedError("standard (M=None) problems are outside eigd's call pattern")
        Bip = as_csr_device(M)
    elif mode == "buckling":
        Bip = as_csr_device(A)
    else:
        raise ValueError("unrecognized mode '%s'" % mode)
    if not hasattr(OPinv, "solve_dev"):
        raise TypeError("OPinv must be an eigd_b200.SpLuOperator (device factorisation); got %r" % type(OPinv))
    if ncv is None:
        ncv = min(n, max(2 * k + 1, 20))
    if ncv > n or ncv <= k:
        raise ValueError("ncv must be k<ncv<=n")
    P = BLOCK_SIZE if block is None else int(block)
    while P > 1 and (P > 4 or ncv % P or ncv - k < 3 * P or ncv + P > 128 or n < ncv + P):
        P -= 1                        # the basis must hold a whole number of blocks (and a few of them): else smaller blocks
    if P > 1:
        st = lanczos_block_thick_restart(Bip, OPinv, k, ncv, float(sigma), P, mode=mode, tol=float(tol), maxiter=maxiter,
                                         v0=v0, seed=seed)
    else:
        st = lanczos_thick_restart(Bip, OPinv, k, ncv, float(sigma), mode=mode, tol=float(tol), maxiter=maxiter, v0=v0, seed=seed)
    if not return_eigenvectors:
        return st.d
    if return_state:                  # the caller keeps the basis and the eigenvectors in HBM (IRAM.solve)
        return (st.d, None, st.T, None, st)
    return (st.d, to_host(st.Z), st.T, to_host(st.Vt).T.copy())
