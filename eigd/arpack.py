"""``eigd.arpack`` of the reference (eigd/arpack.py:104-118: ``eigsh_mod`` returning ``(d, z, Tm, v)``),
served by the device-resident thick-restart Lanczos of ``eigd_b200.arpack``."""
from eigd_b200.arpack import ArpackError, ArpackNoConvergence, eigsh_mod  # noqa: F401

__all__ = ["eigsh_mod", "ArpackError", "ArpackNoConvergence"]
