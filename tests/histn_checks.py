"""Checks of the general-bin-count Label thresholds (csrc/histn.cu) shared by the CPU run (host-emulated) and the GPU run."""
import ctypes as C
import json
import os

import numpy as np
import torch

from conftest import GOLDEN_DIR

NBINS_CASES = ["label_nbins64_iso", "label_nbins1000_iso", "label_nbins100_sample", "label_nbins33_2d",
               "label_nbins64_u16_otsu", "label_nbins500_f32_otsu"]
HIST_WORDS = 3 + 256
TF_NONE, TF_LOG10 = 0, 2


def load_nbins_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    p = np.load(os.path.join(GOLDEN_DIR, f"{str(z['parent'])}.npz"))
    pm = json.loads(str(p["meta"]))
    return dict(raw=p["raw"], frangi=p["frangi"], dim_res=pm["dim_res"], no_z=pm["no_z"],
                kw=json.loads(str(z["meta"]))["label_kwargs"], intensity_thresh=float(z["intensity_thresh"]),
                frangi_thresh=float(z["frangi_thresh"]), labels=z["labels"])


def thresholds_of(lib, device, vals, nbins, log_domain, f64_edges, otsu_only, stream=None):
    """reset -> minmax -> nb200_histn_threshold on a float32 sample vector; returns the 7 output doubles."""
    dev = torch.device(device)
    v = torch.from_numpy(np.ascontiguousarray(vals, dtype=np.float32)).to(dev)
    state = torch.zeros(HIST_WORDS, dtype=torch.int64, device=dev)
    out = torch.zeros(7, dtype=torch.float64, device=dev)
    ws = torch.empty(int(lib.nb200_histn_workspace_bytes(nbins)), dtype=torch.uint8, device=dev)
    assert lib.nb200_hist_reset(state.data_ptr(), stream) == 0
    assert lib.nb200_hist_minmax(v.data_ptr(), v.numel(), TF_LOG10 if log_domain else TF_NONE, None, state.data_ptr(), stream) == 0
    rc = lib.nb200_histn_threshold(v.data_ptr(), v.numel(), int(log_domain), int(f64_edges), int(otsu_only), int(nbins),
                                   state.data_ptr(), ws.data_ptr(), out.data_ptr(), stream)
    assert rc == 0, rc
    return out.cpu().numpy()


def check_against_numpy(lib, device, stream=None):
    """Triangle and Otsu of the oracle (numpy histogram + float64 cumulative sums) for many bin counts and sample shapes:
    log domain (the Frangi threshold), linear float32 (intensity Otsu of a float frame), float64 edges (integer frame);
    constant samples (min == max), two values, heavy ties on bin edges (integers with nbins dividing the range)."""
    from oracle import pipeline as P
    rng = np.random.default_rng(3)
    samples = {
        "lognormal": np.exp(rng.normal(-3.0, 1.2, 20000)).astype(np.float32),
        "bimodal": np.concatenate([rng.normal(0.02, 0.004, 6000), rng.normal(0.3, 0.05, 3000)]).clip(1e-6).astype(np.float32),
        "two": np.array([0.5] * 40 + [2.0] * 7, np.float32),
        "few": rng.random(37).astype(np.float32) + np.float32(0.01),
    }
    for name, vals in samples.items():
        for nbins in (2, 3, 17, 64, 255, 257, 1000, 4096):
            pos = vals[vals > 0]
            out = thresholds_of(lib, device, vals, nbins, 1, 0, 0, stream)
            lv = np.log10(pos)
            try:
                tri, otsu = P.triangle(lv, nbins=nbins), P.otsu(lv, nbins=nbins)
            except ValueError:                                   # argmax of an empty sequence: status flag instead
                assert out[4] == 1.0, (name, nbins)
                continue
            assert out[3] == 0.0 and out[5] == float(tri) and out[6] == float(otsu), (name, nbins, out, tri, otsu)
            lin = thresholds_of(lib, device, vals, nbins, 0, 0, 1, stream)
            assert lin[0] == float(P.otsu(pos, nbins=nbins)), (name, nbins)
    ints = {"u16": rng.integers(0, 4096, 30000).astype(np.uint16), "u8": rng.integers(0, 256, 5000).astype(np.uint8),
            "const": np.full(100, 7, np.uint16)}
    for name, vals in ints.items():
        for nbins in (2, 16, 64, 100, 256, 1000):
            pos = vals[vals > 0]
            got = thresholds_of(lib, device, vals.astype(np.float32), nbins, 0, 1, 1, stream)
            ref = P.otsu(pos, nbins=nbins)
            if np.isnan(ref):
                assert got[4] == 1.0
                continue
            assert isinstance(ref, np.float64) and got[0] == float(ref), (name, nbins, got[0], ref)
    empty = thresholds_of(lib, device, np.zeros(50, np.float32), 64, 1, 0, 0, stream)
    assert empty[3] == 1.0                                       # no positive sample: "None"
    ws = torch.empty(64, dtype=torch.uint8)
    assert lib.nb200_histn_threshold(ws.data_ptr(), 1, 0, 0, 0, 1, ws.data_ptr(), ws.data_ptr(), ws.data_ptr(), stream) != 0


def check_label_class_on_fixture(name):
    """nellie_b200.Label(histogram_nbins=...) against the executed reference: thresholds bit for bit, labels equal."""
    from types import SimpleNamespace
    from nellie_b200 import Label
    g = load_nbins_case(name)
    shape = g["raw"].shape
    info = SimpleNamespace(no_t=True, no_z=g["no_z"], shape=(1,) + shape, axes="TYX" if g["no_z"] else "TZYX", dim_res=g["dim_res"])
    lab = Label(info, device="b200", **g["kw"])
    lab.num_t = 1
    it, ft = lab._compute_frame_thresholds(g["raw"], g["frangi"])
    if np.isnan(g["intensity_thresh"]):
        assert it is None
    else:
        assert float(it) == g["intensity_thresh"], (it, g["intensity_thresh"])
    assert ft == g["frangi_thresh"], (ft, g["frangi_thresh"])
    labels = lab._run_frame_full_volume(0, g["raw"], g["frangi"], it, ft)
    assert np.array_equal(labels, g["labels"])
