"""sleqp_b200 -- B200-native sparse KKT factorization backend for SLEQP.

Host-side mirror of the reference's plugin interface (src/main/fact/fact.h:23-70) over the C-ABI of
libsleqp_b200.so. See DESIGN.md.
"""
from .fact import Fact, Mat, ProjectedCG, Symbolic  # noqa: F401
from ._lib import B200Error  # noqa: F401

__version__ = "0.1.0"
