"""Import the UNMODIFIED reference (``/root/reference``) in the build container.

TEST INFRASTRUCTURE ONLY; used by ``oracle/make_golden.py`` and by the optional
``tests/test_oracle_vs_reference.py`` (skipped when ``/root/reference`` is absent, as on
the GPU box).  The reference's ``Filter``/``Label`` import ``nellie.im_info.verifier``
(needs tifffile / ome_types / nd2) and ``nellie.segmentation.networking`` (needs
skimage); none of those are touched by the hot path, so empty stub modules are seeded
into ``sys.modules`` first.
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("NELLIE_REFERENCE_ROOT", "/root/reference")
_STUBS = ["nd2", "ome_types", "tifffile", "skimage", "skimage.filters", "skimage.morphology",
          "skimage.measure"]


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "nellie", "segmentation", "filtering.py"))


def load():
    """Return (Filter, Label) classes of the reference."""
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    for name in _STUBS:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["tifffile"].tifffile = sys.modules["tifffile"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import logging
    from nellie.segmentation.filtering import Filter
    from nellie.segmentation.labelling import Label
    logging.getLogger().setLevel(logging.WARNING)  # the reference sets the root logger to INFO
    return Filter, Label


def im_info_for(frame_shape, dim_res, no_z):
    """Duck-typed ImInfo, as the reference's own tests build it (tests/test_labelling.py:7-22)."""
    axes = "TYX" if no_z else "TZYX"
    return SimpleNamespace(no_t=True, no_z=no_z, shape=(1,) + tuple(frame_shape), axes=axes,
                           dim_res=dict(dim_res))


def read_sample_frame(t=0):
    """Frame ``t`` of sample_data/yeast_3d_mitochondria.ome.tif via Pillow: (17,192,279) uint16."""
    import numpy as np
    from PIL import Image
    path = os.path.join(REFERENCE_ROOT, "sample_data", "yeast_3d_mitochondria.ome.tif")
    im = Image.open(path)
    planes = []
    for z in range(17):
        im.seek(t * 17 + z)
        planes.append(np.array(im))
    return np.stack(planes)


SAMPLE_DIM_RES = {"X": 0.0655, "Y": 0.0655, "Z": 0.25, "T": 4.536}
