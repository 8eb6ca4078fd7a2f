"""``eigd.eigenvector_derivatives`` of the reference, served by ``eigd_b200.eigenvector_derivatives``."""
from eigd_b200.eigenvector_derivatives import *  # noqa: F401,F403
from eigd_b200.eigenvector_derivatives import __all__  # noqa: F401
