"""The oracle restatement must reproduce the executed reference bit-for-bit (CPU only)."""
import numpy as np
import pytest

from conftest import load_golden, spec_from_meta
from oracle import pipeline as P

FILTER_CASES = ["sample_crop", "phantom3d_iso", "phantom3d_aniso", "phantom2d", "phantom3d_strided", "phantom3d_pow2",
                "phantom3d_cfg3", "phantom3d_nomask"]


@pytest.mark.parametrize("name", FILTER_CASES)
def test_filter_matches_reference(name):
    g = load_golden(name)
    spec = spec_from_meta(g["meta"])
    trace = []
    pre = P.frangi_frame(g["raw"], spec, trace=trace)
    assert np.allclose(P.sigma_schedule(spec), g["sigmas"], rtol=0, atol=0)
    assert [t["gamma"] for t in trace] == g["gamma"].tolist()
    if spec.run_mask:                       # the reference never derives the threshold when mask=False
        assert [t["frob_thr"] for t in trace] == g["frob_thr"].tolist()
    assert np.array_equal(pre, g["frangi_pre"])
    fin = P.finalize_mask(pre, spec)
    assert np.array_equal(fin, g["frangi"])


@pytest.mark.parametrize("name", FILTER_CASES)
def test_label_matches_reference(name):
    g = load_golden(name)
    spec = spec_from_meta(g["meta"])
    it, ft = P.label_thresholds(g["raw"], g["frangi"], spec)
    assert it is None and np.isnan(g["intensity_thresh"])
    assert float(ft) == float(g["frangi_thresh"])
    assert P.label_min_area(spec) == int(g["min_area"])
    labels = P.label_frame(g["frangi"], spec, ft, raw=g["raw"], intensity_thresh=it)
    assert labels.dtype == np.int32
    assert np.array_equal(labels, g["labels"])


@pytest.mark.parametrize("name", ["phantom3d_u16_otsu", "phantom3d_f32_otsu"])
def test_label_intensity_otsu_matches_reference(name):
    """otsu_thresh_intensity=True (labelling.py:457-465, :511-556): the gate threshold is a float64 bin centre for the
    uint16 frame and a float32 one for the float32 frame; the Frangi threshold is taken over the gated sample."""
    g = load_golden(name)
    spec = spec_from_meta(g["meta"])
    assert spec.otsu_thresh_intensity
    assert np.array_equal(P.finalize_mask(P.frangi_frame(g["raw"], spec), spec), g["frangi"])
    it, ft = P.label_thresholds(g["raw"], g["frangi"], spec)
    assert type(it) is (np.float64 if g["raw"].dtype.kind in "iu" else np.float32)
    assert float(it) == float(g["intensity_thresh"]) and float(ft) == float(g["frangi_thresh"])
    labels = P.label_frame(g["frangi"], spec, ft, raw=g["raw"], intensity_thresh=it)
    assert np.array_equal(labels, g["labels"])


@pytest.mark.parametrize("name", ["label3d", "label2d"])
def test_label_only_cases(name):
    g = load_golden(name)
    spec = P.FrameSpec(dim_res=g["meta"]["dim_res"], no_z=g["meta"]["no_z"])
    assert P.label_min_area(spec) == int(g["min_area"])
    labels = P.label_frame(g["frangi"], spec, g["meta"]["frangi_thresh"])
    assert np.array_equal(labels, g["labels"])


def _tiny_spec(no_z=True):
    # the reference's own fixture: tests/test_labelling.py:7-22
    return P.FrameSpec(dim_res={"X": 1.0, "Y": 1.0, "Z": None if no_z else 1.0, "T": 1.0}, no_z=no_z)


def test_reference_unit_case_label_ids_reset_per_frame():
    # tests/test_labelling.py:25-53
    spec = _tiny_spec()
    fr = np.zeros((5, 5), np.float32)
    fr[1:4, 1:4] = 1.0
    for _ in range(2):
        labels = P.label_frame(fr, spec, 0.5)
        assert labels.max() == 1 and set(np.unique(labels)) <= {0, 1}


def test_reference_unit_case_no_input_mutation():
    # tests/test_labelling.py:56-77
    spec = _tiny_spec()
    raw = np.zeros((5, 5), np.float32)
    raw[1:4, 1:4] = 1.0
    fr = raw.copy()
    raw0, fr0 = raw.copy(), fr.copy()
    labels = P.label_frame(fr, spec, 0.5, raw=raw, intensity_thresh=0.5)
    assert labels is not None
    assert np.array_equal(raw, raw0) and np.array_equal(fr, fr0)


def test_filter_does_not_mutate_float32_input():
    g = load_golden("phantom3d_aniso")
    spec = spec_from_meta(g["meta"])
    raw = g["raw"].astype(np.float32)
    keep = raw.copy()
    P.frangi_frame(raw, spec)
    assert np.array_equal(raw, keep)


@pytest.mark.parametrize("name", ["network3d", "network2d"])
def test_network_kernels_match_reference(name):
    """Array kernels of the Network stage (networking.py:261-296, :669-680, :758-797) against the executed reference."""
    import json
    z = np.load(f"{__import__('conftest').GOLDEN_DIR}/{name}.npz")
    no_z = json.loads(str(z["meta"]))["no_z"]
    assert np.array_equal(P.network_remove_connected(z["skel"], no_z), z["cleaned"])
    pc = P.network_pixel_class(z["skel"], no_z)
    assert np.array_equal(pc, z["pixel_class"])
    assert np.array_equal(P.network_branch_labels(pc, no_z), z["branch"])
