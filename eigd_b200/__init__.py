"""eigd_b200 -- B200-native gradient path of smdogroup/eigd (eigensolve + adjoint + df/dx).

Public surface mirrors the reference's ``eigd`` package (eigd/__init__.py:1-3): everything in
``eigenvector_derivatives`` is re-exported.  Compute runs only through libeigd_b200.so.
"""
__version__ = "1.0.0"

from .eigenvector_derivatives import *  # noqa: F401,F403
from .device import pinned_empty  # noqa: E402,F401  (page-locked numpy arrays for the fastest uploads)
