"""opty_b200: a Blackwell-native (sm_100a) collocation-constraint engine behind
the ``opty.direct_collocation.Problem`` / ``ConstraintCollocator`` API."""

from .direct_collocation import Problem, ConstraintCollocator  # noqa: F401
from .utils import parse_free  # noqa: F401

__version__ = '0.1.0'
