"""eigd_b200 -- B200-native gradient path of smdogroup/eigd (eigensolve + adjoint + df/dx).

Public surface mirrors the reference's ``eigd`` package (eigd/__init__.py:1-3): everything in
``eigenvector_derivatives`` is re-exported.  Compute runs only through libeigd_b200.so.
"""
__version__ = "1.0.0"

import os as _os

# The calling thread drives the GPU (launches, stream synchronisations, staged copies).  OpenMP worker threads of the
# caller's CPU libraries spin after every parallel region by default and starve it (measured: 2 x on the end-to-end
# gradient, profiles/r2_e2e_host_threads.txt); passive waiting only takes effect when it is chosen before the OpenMP
# runtime is loaded, i.e. when this package is imported before torch / numpy did any threaded work.  A value set by
# the user wins.
_os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")

from .eigenvector_derivatives import *  # noqa: F401,F403
from .device import pinned_empty  # noqa: E402,F401  (page-locked numpy arrays for the fastest uploads)
