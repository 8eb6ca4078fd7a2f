"""Drop-in alias: ``import eigd`` resolves to the B200-native implementation ``eigd_b200``.

The reference's boundary is its public Python surface: the examples do
``from eigd import IRAM, BasicLanczos, SpLuOperator, eval_adjoint_residual_norm`` (reference
examples/natural_frequency.py:11, thermal.py:11, buckling.py:12) and ``from eigd import
add_eig_total_derivative`` (crm.py:5-10); reference eigd/__init__.py:1-3 re-exports everything in
``eigenvector_derivatives``.  With this package first on ``sys.path`` those examples run unchanged on the
GPU (tests/test_dropin_examples_gpu.py executes the reference's unmodified example classes against it).
"""
from eigd_b200 import __version__  # noqa: F401
from eigd_b200.eigenvector_derivatives import *  # noqa: F401,F403
from eigd_b200.eigenvector_derivatives import __all__  # noqa: F401
from . import arpack, eigenvector_derivatives  # noqa: F401,E402
