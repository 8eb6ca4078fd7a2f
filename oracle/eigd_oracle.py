"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy / scipy) of the reference gradient path.

Restates, routine by routine, ``eigd/eigenvector_derivatives.py`` and the (T, V) contract of
``eigd/arpack.py`` of smdogroup/eigd (paths below are relative to the reference root).  The
heavy arithmetic lives in the same third-party native code the reference calls: SuperLU via
``scipy.sparse.linalg.splu`` and ARPACK ``dsaupd/dseupd`` via scipy's private
``_SymmetricArpackParams`` (scipy 1.18.1 / numpy 2.3.5 in this image; the reference pins
nothing tighter than scipy>=1.7, setup.py:28-31).

Pinning: the reference ships no tests or golden vectors, so this port is
pinned against outputs of the unmodified reference run in the build container through
``oracle/ref_loader.py``; the frozen vectors and the generating script are in tests/golden/.

Only tests/, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of bench.py may import this module.  The product (eigd_b200) never does.
"""
import warnings

import numpy as np
from scipy.sparse.linalg import splu


# ----------------------------------------------------------------------------------------
# factor wrapper: eigd/eigenvector_derivatives.py:11-23
# ----------------------------------------------------------------------------------------
class SpLu:
    def __init__(self, mat):
        self.lu = splu(mat.tocsc())
        self.shape = mat.shape
        self.dtype = mat.dtype
        self.count = 0

    def __call__(self, x):
        x = np.asarray(x)
        if x.ndim == 2:
            self.count += x.shape[1]
            # the reference reaches SuperLU one column at a time (scipy LinearOperator._matmat)
            return np.column_stack([self.lu.solve(np.ascontiguousarray(x[:, j])) for j in range(x.shape[1])])
        self.count += 1
        return self.lu.solve(x.astype(self.dtype))


def project(U, V, X):
    """X <- X - U (V^T X): eigd/eigenvector_derivatives.py:26-30"""
    X -= U @ (V.T @ X)
    return X


def lam_from_theta(theta, sigma, mode):
    """Inverse spectral transform + sort key: :1432-1437, :1960-1965"""
    if mode == "normal":
        lam = 1.0 / theta + sigma
        return lam, np.argsort(lam)
    lam = sigma * theta / (theta - 1.0)
    return lam, np.argsort(-1.0 / lam)


def apply_L(A, B, lam, X, mode):
    """L_i x = (A - lam_i B) x (normal) or (B + lam_i A) x (buckling): :262-265"""
    if mode == "normal":
        return A @ X - (B @ X) * lam
    return B @ X + (A @ X) * lam


# ----------------------------------------------------------------------------------------
# ARPACK with basis extraction: eigd/arpack.py:58-101 (contract), :422-442 (driver)
# ----------------------------------------------------------------------------------------
def eigsh_with_basis(A, B, factor, k, sigma, ncv, mode="normal", tol=0.0, v0=None, rng=None, maxiter=None):
    from scipy.sparse.linalg._eigen.arpack.arpack import _SymmetricArpackParams
    from scipy.sparse.linalg import aslinearoperator, LinearOperator

    n = A.shape[0]
    op = LinearOperator((n, n), matvec=lambda x: factor(x), dtype=float)
    if mode == "normal":
        amode, matvec, M_matvec = 3, None, aslinearoperator(B).matvec
    else:
        # eigd passes A0 = B (stiffness) as the ARPACK "A" in buckling mode: :1941-1942
        amode, matvec, M_matvec = 4, aslinearoperator(B).matvec, None
    p = _SymmetricArpackParams(n, k, "d", matvec, amode, M_matvec, op.matvec, sigma, ncv, v0, maxiter, "LM", tol, rng)
    while not p.converged:
        p.iterate()
    ncv = p.ncv
    h = p.workl[0:2 * ncv].copy()
    V = np.array(p.v, copy=True).reshape(-1).reshape((ncv, n)).T.copy()
    d, z = p.extract(True)
    T = np.diag(h[ncv:2 * ncv]) + np.diag(h[1:ncv], 1) + np.diag(h[1:ncv], -1)
    return d, z, T, V


class IRAMOracle:
    """eigd/eigenvector_derivatives.py:1873-1986"""

    def __init__(self, N=10, m=None, eig_atol=1e-5, tol=0.0, mode="normal"):
        self.N = N
        self.m = max(20, 2 * N + 1) if m is None else max(20, 2 * N + 1, m)
        self.tol, self.eig_atol, self.mode = tol, eig_atol, mode

    def solve(self, A, B, factor, sigma, rng=None):
        self.A, self.B, self.factor, self.sigma = A, B, factor, sigma
        self.lam, self.Phi, self.T, self.V = eigsh_with_basis(A, B, factor, self.N, sigma, self.m, self.mode, self.tol, rng=rng)
        self.theta, self.Y = np.linalg.eigh(self.T)
        eigs, self.indices = lam_from_theta(self.theta, sigma, self.mode)
        if abs(eigs[self.indices[self.N - 1]] - eigs[self.indices[self.N]]) < self.eig_atol:
            warnings.warn("IRAM: Ritz values %d and %d are numerically repeated." % (self.N, self.N + 1))
        for i in range(self.N):  # modal-assurance sign alignment, :1978-1984
            q = self.V @ self.Y[:, self.indices[i]]
            if self.Phi[:, i] @ q < 0.0:
                self.Y[:, self.indices[i]] *= -1.0
        return self.lam, self.Phi

    def solve_adjoint(self, Phib, method="sibk", rtol=1e-10, atol=1e-30, lanczos_guess=True, **kw):
        return solve_adjoint(self, self.lam, self.V, Phib, method, rtol, atol, lanczos_guess, **kw)


class BasicLanczosOracle:
    """eigd/eigenvector_derivatives.py:1331-1650 (ortho_type='full')"""

    def __init__(self, N=10, m=60, tol=1e-14, Ntarget=None, eig_atol=1e-5, mode="normal"):
        self.N, self.m_max, self.tol, self.Ntarget, self.eig_atol, self.mode = N, m, tol, Ntarget, eig_atol, mode

    def _reduced(self, alpha, beta, sigma, m):
        T = np.diag(alpha[:m]) + np.diag(beta[:m - 1], 1) + np.diag(beta[:m - 1], -1)
        theta, Y = np.linalg.eigh(T)
        lam, idx = lam_from_theta(theta, sigma, self.mode)
        return theta, Y, T, lam, idx

    def solve(self, A, B, factor, sigma):
        n = A.shape[0]
        self.A, self.B, self.factor, self.sigma = A, B, factor, sigma
        alpha, beta = np.zeros(self.m_max), np.zeros(self.m_max)
        V = np.zeros((n, self.m_max + 1))
        V[:, 0] = np.random.default_rng(12345).uniform(size=n, low=-1.0, high=1.0)
        V[:, 0] /= np.sqrt(V[:, 0] @ (B @ V[:, 0]))
        self.m = self.m_max
        for i in range(1, self.m_max + 1):
            w = factor(B @ V[:, i - 1])
            if i > 1:
                w -= beta[i - 2] * V[:, i - 2]
            for j in range(i - 1, -1, -1):  # modified Gram-Schmidt, B inner product, descending
                h = w @ (B @ V[:, j])
                w -= h * V[:, j]
                if j == i - 1:
                    alpha[i - 1] = h
            beta[i - 1] = np.sqrt(w @ (B @ w))
            V[:, i] = w / beta[i - 1]
            if i >= 2:
                theta, Y, T, lam, idx = self._reduced(alpha, beta, sigma, i)
                Nc = self.N if self.Ntarget is None else self.Ntarget
                err = np.abs(beta[i - 1] * Y[i - 1, idx])
                bad = np.nonzero(err >= self.tol)[0]
                if (len(err) if len(bad) == 0 else bad[0]) >= Nc:
                    self.m = i
                    break
        self.alpha, self.beta, self.Vfull = alpha, beta, V
        self.theta, self.Y, self.T, self.lam, self.indices = self._reduced(alpha, beta, sigma, self.m)
        if self.Ntarget is not None:
            self.N = self.Ntarget
            while self.N < self.m and abs(self.lam[self.indices[self.N - 1]] - self.lam[self.indices[self.N]]) < self.eig_atol:
                self.N += 1
        self.lam0 = self.lam[self.indices[:self.N]]
        self.Y0 = self.Y[:, self.indices[:self.N]]
        self.eig_res = np.abs(beta[-1] * self.Y0[-1, :])
        self.V = V[:, :self.m]
        self.Phi = self.V @ self.Y0
        return self.lam0, self.Phi

    def solve_adjoint(self, Phib, method="sibk", rtol=1e-10, atol=1e-30, lanczos_guess=True, **kw):
        return solve_adjoint(self, self.lam0, self.V, Phib, method, rtol, atol, lanczos_guess, **kw)


def solve_adjoint(s, lam, V, Phib, method, rtol, atol, lanczos_guess, **kw):
    """Dispatch of IRAM/BasicLanczos.solve_adjoint: :1652-1797, :1988-2134"""
    n = Phib.shape[0]
    if method == "dl":
        lanczos_guess = False
    if lanczos_guess or method == "laa":
        psi = laa(Phib, s.B, s.factor, s.sigma, lam, V, s.Y, s.theta, s.indices, mode=s.mode)
    else:
        psi = np.zeros((n, len(lam)))
    if method == "laa":
        return psi, adjoint_correction(lam, s.Phi, psi, Phib=Phib, eig_atol=s.eig_atol, mode=s.mode)
    if method == "dl":
        return dl(Phib, s.B, s.factor, s.sigma, lam, s.Phi, s.indices, V, s.T, s.Y, s.theta, s.eig_atol, s.mode)
    fn = {"sibk": sibk, "pcpg": pcpg, "pgmres": pgmres}[method]
    psi, data, _ = fn(Phib, s.A, s.B, lam, s.Phi, mode=s.mode, psi=psi, sigma=s.sigma, factor=s.factor, rtol=rtol,
                      atol=atol, eig_atol=s.eig_atol, **kw)
    return psi, data


# ----------------------------------------------------------------------------------------
# adjoint correction and total derivative: :303-391, :33-182, :185-275
# ----------------------------------------------------------------------------------------
def adjoint_correction(lam, Phi, psi, G=None, Phib=None, eig_atol=1e-5, mode="normal"):
    N = len(lam)
    if G is None:
        G = -Phi.T @ Phib
    G0 = G if mode == "normal" else np.diag(lam) @ G
    data = {}
    for i in range(N):
        for j in range(i):
            dl_ = lam[j] - lam[i]
            if abs(lam[i] - lam[j]) < eig_atol:
                xi = 0.5 * (G0[j, i] - G0[i, j]) / dl_
                eta = 0.5 * (lam[i] * G0[j, i] - lam[j] * G0[i, j]) / dl_
                data.setdefault(i, []).append((j, xi, eta))
                data.setdefault(j, []).append((i, xi, eta))
            else:
                psi[:, i] += (G0[j, i] / dl_) * Phi[:, j]
                psi[:, j] += (G0[i, j] / (-dl_)) * Phi[:, i]
    return data


def total_derivative_W(lam, Phi, lamb, Phib, psi, data, mode):
    """The two (n, N) weight matrices of the 'tensor' form: :135-180.
    Returns (WA, WB, signB): dfdx += dAdx(WA, Phi); dfdx += signB * dBdx(WB, Phi)."""
    N = Phi.shape[1]
    beta = 0.5 * np.einsum("ij,ij->j", Phi, Phib)
    if mode == "normal":
        WA = Phi * lamb + psi
        WB = Phi * (beta + lam * lamb) + psi * lam
        cA, cB, signB = 1, 2, -1.0  # xi feeds A, eta feeds B
    else:
        WA = (Phi * lamb + psi) * lam
        WB = Phi * (lamb - beta) + psi
        cA, cB, signB = 2, 1, 1.0   # eta feeds A, xi feeds B
    for i in range(N):
        for item in data.get(i, []):
            WA[:, i] += item[cA] * Phi[:, item[0]]
            WB[:, i] += item[cB] * Phi[:, item[0]]
    return WA, WB, signB


def add_total_derivative(lam, Phi, lamb, Phib, psi, dAdx, dBdx, dfdx, data=None, mode="normal"):
    WA, WB, signB = total_derivative_W(lam, Phi, lamb, Phib, psi, data or {}, mode)
    if dAdx is not None:
        dfdx += dAdx(WA, Phi)
    if dBdx is not None:
        dfdx += signB * dBdx(WB, Phi)
    return dfdx


def adjoint_residual_norm(A, B, lam, Phi, Phib, psi, mode="normal", b_ortho=False):
    N = Phi.shape[1]
    BPhi = B @ Phi
    res, ortho = np.zeros(N), np.zeros(N)
    for i in range(N):
        b = -(Phib[:, i] - BPhi[:, i] * (Phi[:, i] @ Phib[:, i]))
        r = apply_L(A, B, lam[i], psi[:, i], mode) - b
        if b_ortho:
            r = project(BPhi, Phi, r)
            ortho[i] = np.abs(BPhi.T @ psi[:, i]).max()
        else:
            ortho[i] = abs(BPhi[:, i] @ psi[:, i])
        res[i] = np.linalg.norm(r)
    return res, ortho


# ----------------------------------------------------------------------------------------
# Lanczos adjoint approximation: :394-523 (b_ortho=True branch, the only one the classes use)
# ----------------------------------------------------------------------------------------
def laa(Phib, B, factor, sigma, lam, V, Y, theta, indices, mode="normal"):
    m, N = len(theta), Phib.shape[1]
    Yb = V.T @ Phib
    D = np.zeros((m, N))
    rest, first = indices[N:], indices[:N]
    # D[indices[i], j] = Y[:, indices[i]] . Yb[:, j] / (theta[indices[j]] - theta[indices[i]]), i >= N
    D[rest, :] = (Y[:, rest].T @ Yb) / (theta[first][None, :] - theta[rest][:, None])
    S = D / (lam - sigma)
    if mode == "buckling":
        S = sigma * S
    return -factor(B @ (V @ (Y @ S)))


# ----------------------------------------------------------------------------------------
# shift-invert block Krylov, bs_target = 1, update_guess = False: :1052-1328
# ----------------------------------------------------------------------------------------
def _lstsq(alpha, H, r):
    H0 = np.eye(H.shape[0], H.shape[1]) - alpha * H
    y = np.linalg.lstsq(H0, r, rcond=None)[0]
    return y, np.linalg.norm(H0 @ y - r)


def sibk(Phib, A, B, lam, Phi, mode="normal", psi=None, sigma=None, factor=None, rtol=1e-10, atol=1e-30,
         eig_atol=1e-5, maxiter=50, callback=None, nrestart=2, **unused):
    n, N = Phib.shape
    rnorm0 = np.sqrt(np.max(np.sum(Phib**2, axis=0)))
    BPhi = B @ Phi
    G = -Phi.T @ Phib
    psi = np.zeros((n, N)) if psi is None else psi
    R = project(BPhi, Phi, -Phib - apply_L(A, B, lam, psi, mode))
    info = []
    for i in range(N):
        W = np.zeros((n, maxiter + 1))
        Z = np.zeros((n, maxiter))
        H = np.zeros((maxiter + 1, maxiter))
        beta0 = np.linalg.norm(R[:, i])
        if callback is not None:
            callback(beta0)
        if beta0 < rtol * rnorm0 or beta0 < atol:
            info.append(0)
            continue
        w = project(BPhi, Phi, R[:, i].copy())
        r0 = np.linalg.norm(w)
        W[:, 0] = w / r0
        alpha = (lam[i] - sigma) if mode == "normal" else -(lam[i] - sigma)
        for j in range(1, maxiter + 1):
            Z[:, j - 1] = factor(W[:, j - 1])
            w = project(BPhi, Phi, (B if mode == "normal" else A) @ Z[:, j - 1])
            for k in range(j - 1, -1, -1):
                H[k, j - 1] = w @ W[:, k]
                w -= H[k, j - 1] * W[:, k]
            w = project(BPhi, Phi, w)
            H[j, j - 1] = np.linalg.norm(w)
            W[:, j] = w / H[j, j - 1]
            rhs = np.zeros(j + 1)
            rhs[0] = r0
            y, res = _lstsq(alpha, H[:j + 1, :j], rhs)
            if callback is not None:
                callback(res)
            if res < rtol * rnorm0 or res < atol or j == maxiter:
                psi[:, i] += Z[:, :j] @ y
                info.append(j if res < rtol * rnorm0 or res < atol else -1)
                break
    data = adjoint_correction(lam, Phi, psi, G=G, eig_atol=eig_atol, mode=mode)
    return psi, data, info


# ----------------------------------------------------------------------------------------
# projected preconditioned CG: :699-869
# ----------------------------------------------------------------------------------------
def pcpg(Phib, A, B, lam, Phi, mode="normal", psi=None, sigma=None, factor=None, rtol=1e-10, atol=1e-30,
         eig_atol=1e-5, maxiter=100, reset=25, callback=None):
    n, N = Phib.shape
    psi = np.zeros((n, N)) if psi is None else psi
    rnorm0 = np.sqrt(np.max(np.sum(Phib**2, axis=0)))
    BPhi = B @ Phi
    G = np.zeros((N, N))
    info = []
    for i in range(N):
        R = -Phib[:, i] - apply_L(A, B, lam[i], psi[:, i], mode)
        G[:, i] = Phi.T @ R
        R -= BPhi @ G[:, i]
        P0, prev, ok = np.zeros(n), 1.0, False
        for k in range(maxiter):
            res = np.linalg.norm(R)
            if callback is not None:
                callback(res)
            if res < rtol * rnorm0 or res < atol:
                ok = True
                break
            Z = project(Phi, BPhi, factor(project(BPhi, Phi, R.copy())))
            zr = Z @ R
            P = Z.copy() if k % reset == 0 else Z + (zr / prev) * P0
            prev = zr
            LP = apply_L(A, B, lam[i], P, mode)
            a = zr / (LP @ P)
            psi[:, i] += a * P
            R = R - a * LP
            P0 = P
        info.append(ok)
    data = adjoint_correction(lam, Phi, psi, G=G, eig_atol=eig_atol, mode=mode)
    return psi, data, info


# ----------------------------------------------------------------------------------------
# projected right-preconditioned GMRES: :872-1040
# ----------------------------------------------------------------------------------------
def pgmres(Phib, A, B, lam, Phi, mode="normal", psi=None, sigma=None, factor=None, rtol=1e-10, atol=1e-30,
           eig_atol=1e-5, maxiter=50, callback=None):
    n, N = Phib.shape
    psi = np.zeros((n, N)) if psi is None else psi
    rnorm0 = np.sqrt(np.max(np.sum(Phib**2, axis=0)))
    BPhi = B @ Phi
    G = np.zeros((N, N))
    info = []
    for i in range(N):
        R = -Phib[:, i] - apply_L(A, B, lam[i], psi[:, i], mode)
        G[:, i] = Phi.T @ R
        R -= BPhi @ G[:, i]
        beta = np.linalg.norm(R)
        if callback is not None:
            callback(beta)
        if beta < rtol * rnorm0 or beta < atol:
            info.append(0)
            continue
        W = np.zeros((n, maxiter + 1))
        Z = np.zeros((n, maxiter))
        H = np.zeros((maxiter + 1, maxiter))
        W[:, 0] = R / beta
        for j in range(maxiter):
            Z[:, j] = factor(project(BPhi, Phi, W[:, j].copy()))
            w = project(BPhi, Phi, apply_L(A, B, lam[i], Z[:, j], mode))
            for k in range(j + 1):
                H[k, j] = w @ W[:, k]
                w -= H[k, j] * W[:, k]
            H[j + 1, j] = np.linalg.norm(w)
            W[:, j + 1] = w / H[j + 1, j]
            rhs = np.zeros(j + 2)
            rhs[0] = beta
            y = np.linalg.lstsq(H[:j + 2, :j + 1], rhs, rcond=None)[0]
            res = np.linalg.norm(H[:j + 2, :j + 1] @ y - rhs)
            if callback is not None:
                callback(res)
            if res < rtol * rnorm0 or res < atol or j == maxiter - 1:
                psi[:, i] += Z[:, :j + 1] @ y
                info.append(j if res < rtol * rnorm0 or res < atol else -1)
                break
    data = adjoint_correction(lam, Phi, psi, G=G, eig_atol=eig_atol, mode=mode)
    return psi, data, info


# ----------------------------------------------------------------------------------------
# differentiated Lanczos (reverse mode through the three-term recurrence): :526-696
# ----------------------------------------------------------------------------------------
def dl(Phib, B, factor, sigma, lam, Phi, indices, V, T, Y, theta, eig_atol=1e-5, mode="normal"):
    m, N = len(theta), Phib.shape[1]
    repeated = bool(np.any(np.abs(np.diff(lam)) < eig_atol))
    first = indices[:N]
    G = BPhi = None
    R = Phib
    if repeated:
        BPhi = B @ Phi
        G = -Phi.T @ Phib
        R = Phib + BPhi @ G
    Vb = R @ Y[:, first].T
    Yb = V.T @ R
    D = np.zeros((m, m))
    for i in range(m):
        for j in range(N):
            ii, jj = indices[i], indices[j]
            if ii == jj or (i < N and abs(lam[i] - lam[j]) < eig_atol):
                continue
            D[ii, jj] = (Y[:, ii] @ Yb[:, j]) / (theta[jj] - theta[ii])
    Tb = Y @ (D @ Y.T)
    t = B @ factor(B @ V[:, m - 1])
    Vb += np.outer(t, Tb[:m, m - 1])
    u = factor(B @ (V @ Tb[:, m - 1]))
    Vb[:, m - 1] += B @ u
    for i in range(m - 2, -1, -1):
        lo = max(i - 1, 0)
        t = B @ (V[:, lo:i + 2] @ T[lo:i + 2, i])
        c0 = V[:, i + 1] @ Vb[:, i + 1] - T[i + 1, i] * Tb[i + 1, i]
        sb = (Vb[:, i + 1] - c0 * (B @ V[:, i + 1])) / T[i + 1, i]
        if i > 0:
            Vb[:, i - 1] -= T[i - 1, i] * sb
        Vb[:, i] -= T[i, i] * sb
        hb = V[:, :i + 1].T @ sb - Tb[:i + 1, i]
        Vb[:, :i + 1] -= np.outer(t, hb)
        sb = sb - B @ (V[:, :i + 1] @ hb)
        Vb[:, i + 1] = u
        u = factor(sb)
        Vb[:, i] += B @ u
    Vb[:, 0] = u
    S = Y[:, first] / (lam - sigma)
    if mode == "buckling":
        S = sigma * S
    psi = -Vb @ S
    data = {}
    if repeated:
        psi = project(Phi, BPhi, psi)
        data = adjoint_correction(lam, Phi, psi, G=G, eig_atol=eig_atol, mode=mode)
    return psi, data
