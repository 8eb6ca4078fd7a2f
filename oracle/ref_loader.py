"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference (smdogroup/eigd).

The reference (pure Python, /root/reference) does not import under scipy >= 1.15
(SURVEY.md section 8c): its ``eigd/arpack.py`` binds scipy-private names that no longer
exist, and ``SpLuOperator`` never runs ``LinearOperator.__init__``.  This module loads the
reference's ``eigd/eigenvector_derivatives.py`` verbatim from where it lies, with

  * a replacement ``eigd.arpack`` module exposing ``eigsh_mod`` (same 4-tuple contract
    as reference eigd/arpack.py:58-101: ``(d, z, Tm, v)``), built on scipy's current
    ``_SymmetricArpackParams``;
  * ``SpLuOperator.__init__`` wrapped so ``LinearOperator.__init__`` runs first;
  * a stub ``matplotlib`` so the examples import.

It exists so that ``tests/golden/make_golden*.py`` can run the real reference in the build container
and freeze its outputs as fixtures, so that ``bench.py --impl reference`` can time the reference's own
CPU path, and so that ``tests/test_dropin_examples_gpu.py`` can run the reference's unmodified example
drivers against the ``eigd`` alias of this repository.  The tree it loads is ``baseline/_ref`` (offline
install made by ``baseline/install_reference.py``; git-ignored, shipped to the GPU box) and, only in the
build container and as a last resort, ``/root/reference``.  Nothing in the product imports this file.
"""
import importlib
import importlib.util
import os
import sys
import types
from unittest import mock

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _default_root():
    """EIGD_REFERENCE_ROOT, else the offline install baseline/_ref (baseline/install_reference.py: the copy that
    travels to the GPU box), else the build container's read-only /root/reference."""
    env = os.environ.get("EIGD_REFERENCE_ROOT")
    if env:
        return env
    for cand in (os.path.join(_REPO, "baseline", "_ref"), "/root/reference"):
        if os.path.isfile(os.path.join(cand, "eigd", "eigenvector_derivatives.py")):
            return cand
    return "/root/reference"


REF_ROOT = _default_root()


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "eigd", "eigenvector_derivatives.py"))


def _make_arpack_module():
    from scipy.sparse.linalg._eigen.arpack.arpack import _SymmetricArpackParams
    from scipy.sparse.linalg._interface import aslinearoperator

    class _Params(_SymmetricArpackParams):
        def extract_all(self):
            ncv = self.ncv
            # tridiagonal T lives in workl[0:2*ncv]: off-diagonal h[1:ncv], diagonal h[ncv:2ncv]
            h = self.workl[0 : 2 * ncv].copy()
            # the C ARPACK writes V column-major into scipy's C-ordered (n, ncv) buffer
            v = np.array(self.v, copy=True).reshape(-1).reshape((ncv, self.n)).T.copy()
            d, z = self.extract(True)
            Tm = np.zeros((ncv, ncv))
            idx = np.arange(ncv - 1)
            Tm[idx, idx + 1] = h[1:ncv]
            Tm[idx + 1, idx] = h[1:ncv]
            Tm[np.arange(ncv), np.arange(ncv)] = h[ncv : 2 * ncv]
            return d, z, Tm, v

    def eigsh_mod(A, k=6, M=None, sigma=None, which="LM", v0=None, ncv=None, maxiter=None,
                  tol=0, return_eigenvectors=True, Minv=None, OPinv=None, mode="normal",
                  rng=None):
        n = A.shape[0]
        if sigma is None or OPinv is None:
            raise NotImplementedError("shim covers the shift-invert call sites of eigd only")
        if mode == "normal":
            amode, matvec = 3, None
            M_matvec = aslinearoperator(M).matvec if M is not None else None
        elif mode == "buckling":
            amode, M_matvec = 4, None
            matvec = aslinearoperator(A).matvec
        else:
            raise ValueError("unrecognized mode '%s'" % mode)
        Minv_matvec = aslinearoperator(OPinv).matvec
        p = _Params(n, k, np.dtype(A.dtype).char, matvec, amode, M_matvec, Minv_matvec,
                    sigma, ncv, v0, maxiter, which, tol, rng)
        while not p.converged:
            p.iterate()
        return p.extract_all()

    mod = types.ModuleType("eigd.arpack")
    mod.eigsh_mod = eigsh_mod
    return mod


def install_matplotlib_stub():
    if "matplotlib" in sys.modules:
        return
    mpl = types.ModuleType("matplotlib")
    for name in ("pylab", "pyplot", "tri"):
        sub = types.ModuleType("matplotlib." + name)
        sub.subplots = lambda *a, **k: (mock.MagicMock(), mock.MagicMock())
        sub.show = lambda *a, **k: None
        sub.close = lambda *a, **k: None
        sub.Triangulation = mock.MagicMock()
        sub.figure = mock.MagicMock()
        sub.savefig = lambda *a, **k: None
        setattr(mpl, name, sub)
        sys.modules["matplotlib." + name] = sub
    sys.modules["matplotlib"] = mpl


_LOADED = {}


def load_reference(force=False):
    """Return the reference ``eigd`` package (module object) loaded from REF_ROOT."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if not force and getattr(sys.modules.get("eigd"), "_is_reference", False):
        return sys.modules["eigd"]
    if not force and _LOADED:
        for k, v in _LOADED.items():
            sys.modules[k] = v
        return _LOADED["eigd"]
    from scipy.sparse.linalg import LinearOperator

    pkg = types.ModuleType("eigd")
    pkg.__path__ = [os.path.join(REF_ROOT, "eigd")]
    pkg._is_reference = True
    sys.modules["eigd"] = pkg
    sys.modules["eigd.arpack"] = _make_arpack_module()
    sys.modules["eigd.arpack"]._is_reference = True
    spec = importlib.util.spec_from_file_location(
        "eigd.eigenvector_derivatives", os.path.join(REF_ROOT, "eigd", "eigenvector_derivatives.py"))
    ed = importlib.util.module_from_spec(spec)
    ed.__package__ = "eigd"
    ed._is_reference = True
    sys.modules["eigd.eigenvector_derivatives"] = ed
    spec.loader.exec_module(ed)
    _orig = ed.SpLuOperator.__init__

    def _init(self, mat):
        LinearOperator.__init__(self, mat.dtype, mat.shape)
        _orig(self, mat)

    ed.SpLuOperator.__init__ = _init
    for k, v in vars(ed).items():
        if not k.startswith("_"):
            setattr(pkg, k, v)
    pkg.arpack = sys.modules["eigd.arpack"]
    pkg.eigenvector_derivatives = ed
    pkg.__version__ = "1.0.0"
    _LOADED.update({"eigd": pkg, "eigd.arpack": pkg.arpack, "eigd.eigenvector_derivatives": ed})
    return pkg


def load_example(name, against="reference"):
    """Import reference examples/<name>.py (not as __main__).

    against="reference": with the shimmed reference ``eigd`` (golden generation, the CPU reference arm).
    against="alias":     with ``eigd`` resolving to this repository's drop-in alias package (``eigd/`` at the
                         repo root -> eigd_b200), i.e. the unmodified example driver on the GPU path.
    The example binds its ``eigd`` names at import time, so the two flavours can coexist in one process."""
    install_matplotlib_stub()
    exdir = os.path.join(REF_ROOT, "examples")
    if not os.path.isfile(os.path.join(exdir, name + ".py")):
        raise RuntimeError("reference example %s.py not present under %s" % (name, exdir))
    if exdir not in sys.path:
        sys.path.insert(0, exdir)
    modname = ("_ref_example_" if against == "reference" else "_alias_example_") + name
    if modname in sys.modules:
        return sys.modules[modname]
    names = ("eigd", "eigd.arpack", "eigd.eigenvector_derivatives")
    saved = {k: sys.modules.get(k) for k in names}
    try:
        if against == "reference":
            load_reference()                     # registers the reference under the three names
        elif against == "alias":
            for k in names:
                if getattr(sys.modules.get(k), "_is_reference", False) or k != "eigd":
                    sys.modules.pop(k, None)
            if getattr(sys.modules.get("eigd"), "_is_reference", False):
                sys.modules.pop("eigd")
            if _REPO not in sys.path:
                sys.path.insert(0, _REPO)
            alias = importlib.import_module("eigd")
            if getattr(alias, "_is_reference", False) or "eigd_b200" not in getattr(alias.IRAM, "__module__", ""):
                raise RuntimeError("eigd did not resolve to the alias package of this repository")
        else:
            raise ValueError(against)
        spec = importlib.util.spec_from_file_location(modname, os.path.join(exdir, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():               # leave sys.modules["eigd"...] as the caller had it
            if v is not None:
                sys.modules[k] = v
            elif against == "alias":
                pass                             # the alias is the importable `eigd` of this repository: keep it
            else:
                sys.modules.pop(k, None)
    return mod
