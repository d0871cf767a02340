"""OME-TIFF layer under the stages (nellie_b200/imio.py; reference: nellie/im_info/verifier.py:967-1070)."""
import os

import numpy as np
import pytest

from nellie_b200 import imio

SAMPLE = "/root/reference/sample_data/yeast_3d_mitochondria.ome.tif"


@pytest.mark.parametrize("dtype", ["uint16", "float32", "int32"])
@pytest.mark.parametrize("shape,axes", [((3, 5, 17, 23), "TZYX"), ((4, 33, 20), "TYX"), ((1, 6, 9, 11), "TZYX")])
def test_bigtiff_round_trip_and_memmap(tmp_path, dtype, shape, axes):
    rng = np.random.default_rng(1)
    data = (rng.random(shape) * 1000).astype(dtype)
    path = str(tmp_path / "a.ome.tif")
    imio.write_ome_bigtiff(path, shape, dtype, axes, {"X": 0.1, "Y": 0.1, "Z": 0.3, "T": 2.0}, "desc <&>", data)
    tf = imio.TiffFile(path)
    assert tf.big and len(tf.pages) == int(np.prod(shape[:-2]))
    assert tf.ome_shape_axes() == (shape, axes)
    assert tf.contiguous_offset() % imio.DATA_ALIGN == 0
    assert np.array_equal(imio.read_tiff(path), data)
    mm = imio.memmap_ome_tiff(path, "r+")
    assert mm.shape == shape and mm.dtype == np.dtype(dtype) and np.array_equal(mm, data)
    mm[0, ...] = 7                                   # the stages' write pattern: memmap[t, ...] = frame; flush()
    mm.flush()
    del mm
    again = imio.read_tiff(path)
    assert (again[0] == 7).all() and np.array_equal(again[1:], data[1:])


def test_allocate_memory_is_zero_filled_and_described(tmp_path):
    info = imio.StackInfo.from_array(np.zeros((2, 4, 8, 8), np.uint16), "TZYX", {"X": 0.2, "Y": 0.2, "Z": 0.5, "T": 1.0},
                                     str(tmp_path), "cells")
    assert not info.no_z and not info.no_t
    assert os.path.basename(info.im_path) == "cells-TZYX-T1p0_Z0p5_Y0p2_X0p2-ch0-t0_to_1.ome.tif"
    assert info.pipeline_paths["im_preprocessed"].endswith("-ch0-t0_to_1-im_preprocessed.ome.tif")
    mm = info.allocate_memory(info.pipeline_paths["im_instance_label"], dtype="int32", description="instance labels",
                              return_memmap=True)
    assert mm.shape == (2, 4, 8, 8) and mm.dtype == np.int32 and not mm.any()
    desc = imio.TiffFile(info.pipeline_paths["im_instance_label"]).pages[0].description
    assert 'Type="int32"' in desc and "instance labels" in desc and 'PhysicalSizeZ="0.5"' in desc
    raw = info.get_memmap(info.im_path)
    assert raw.shape == (2, 4, 8, 8)


def test_single_frame_keeps_its_t_axis(tmp_path):
    info = imio.StackInfo.from_array(np.ones((5, 6, 7), np.float32), "ZYX", {"X": 1.0, "Y": 1.0, "Z": 1.0, "T": None},
                                     str(tmp_path), "one")
    assert info.axes == "TZYX" and info.shape == (1, 5, 6, 7) and info.no_t and not info.no_z
    assert info.get_memmap(info.im_path).shape == (1, 5, 6, 7)


@pytest.mark.skipif(not os.path.exists(SAMPLE), reason="reference sample only exists in the build container")
def test_reads_the_reference_sample(tmp_path):
    tf = imio.TiffFile(SAMPLE)
    assert not tf.big and tf.bo == ">" and len(tf.pages) == 510
    assert tf.contiguous_offset() is None           # strips interleaved with IFDs: must be converted, not mapped
    shape, axes = tf.ome_shape_axes()
    assert axes == "TZYX" and shape == (30, 17, 192, 279)
    data = tf.read()
    from PIL import Image
    im = Image.open(SAMPLE)
    for page in (0, 17 * 3 + 5, 509):
        im.seek(page)
        assert np.array_equal(np.asarray(im), data.reshape(510, 192, 279)[page])
    info = imio.StackInfo.from_tiff(SAMPLE, {"X": 0.0655, "Y": 0.0655, "Z": 0.25, "T": 3.0}, str(tmp_path))
    assert info.shape == (30, 17, 192, 279) and info.dtype == np.uint16
    assert np.array_equal(info.get_memmap(info.im_path, "r"), data)
