"""nellie_b200 — B200-native (sm_100a) implementation of nellie's Filter + Label hot path.

Public surface mirrors ``nellie.segmentation``: :class:`Filter`, :class:`Label`.
"""
from .filtering import Filter  # noqa: F401

try:  # Label arrives with label.cu
    from .labelling import Label  # noqa: F401
except ImportError:  # pragma: no cover
    pass

__all__ = ["Filter", "Label"]
