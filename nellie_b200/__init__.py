"""nellie_b200 — B200-native (sm_100a) implementation of nellie's Filter + Label hot path.

Public surface mirrors ``nellie.segmentation``: :class:`Filter`, :class:`Label`, :class:`Markers`; ``imio`` is the minimal OME-TIFF
layer under them (``StackInfo`` = the ``ImInfo`` attributes and methods the two stages use) and
``pipeline.FramePipeline`` the double-buffered H2D / compute / D2H frame stream behind ``Filter.run``.
"""
from .filtering import Filter  # noqa: F401
from .labelling import Label  # noqa: F401
from .mocap_marking import Markers  # noqa: F401

__all__ = ["Filter", "Label", "Markers"]
