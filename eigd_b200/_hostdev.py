"""Host <-> device plumbing shared by the public modules.

The public API keeps the reference's numpy-in / numpy-out conventions
(SURVEY.md section 8b); internally everything is a CUDA fp64 torch tensor.  Callers that already
hold device data (torch CUDA tensors, ``device.CsrDevice``) skip the copies and get device
tensors back.
"""
import numpy as np
import torch

from . import device as D


# bytes moved across PCIe by the public API since the last reset (bench.py "e2e" accounting)
XFER = {"h2d": 0, "d2h": 0}


def is_dev(x):
    return isinstance(x, torch.Tensor)


def as_csr_device(A):
    """scipy sparse / CsrDevice -> CsrDevice (uploads once per call site)."""
    if isinstance(A, D.CsrDevice):
        return A
    inner = getattr(A, "_eigd_csr_device", None)
    if isinstance(inner, D.CsrDevice):
        return inner
    # scipy LinearOperator wrappers around a sparse matrix (aslinearoperator) keep it in .A
    base = getattr(A, "A", None)
    if base is not None and hasattr(base, "tocsr"):
        A = base
    if hasattr(A, "tocsr"):
        if np.iscomplexobj(getattr(A, "data", 0.0)):
            raise NotImplementedError("eigd_b200: complex (complex-step) matrices are supported by SpLuOperator and "
                                      "BasicLanczos.solve only (eigd_b200/dual.py), as in the reference's own checks")
        out = D.CsrDevice.from_scipy(A)
        XFER["h2d"] += out.uploaded_bytes
        return out
    raise TypeError("eigd_b200 needs a scipy sparse matrix or a device.CsrDevice, got %r (there is no "
                    "dense / LinearOperator CPU fallback)" % type(A))


def to_dev(x, copy=False):
    """numpy / torch -> CUDA fp64 tensor, preserving the (n, N) row-major layout."""
    if is_dev(x):
        t = x.to(device=D.dev(), dtype=D.F64)
        return t.clone() if (copy and t.data_ptr() == x.data_ptr()) else t
    a = np.asarray(x)
    if np.iscomplexobj(a):
        raise NotImplementedError("eigd_b200: complex (complex-step) operands are not supported on the device path")
    XFER["h2d"] += a.size * 8
    return D.h2d(np.ascontiguousarray(a, dtype=np.float64))


def to_host(t):
    XFER["d2h"] += t.numel() * t.element_size()
    return D.d2h(t)


def like_input(t, ref):
    """Return the device tensor ``t`` in the flavour of the caller's operand ``ref``."""
    return t if is_dev(ref) else to_host(t)


def small_to_dev(a):
    XFER["h2d"] += np.size(a) * 8
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=D.dev())
