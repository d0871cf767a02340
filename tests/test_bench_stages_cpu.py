"""Dry run of scripts/bench_stages.py (the `hierarchy_stages` section of the N=1 bench line) without a GPU: the script's own
code — stage construction, engine calls, JSON assembly — runs against host-emulated kernels, a CPU device and timer stand-ins
for CUDA events; Filter / Label (GPU-only kernels) are replaced by a threshold and scipy's labelling.  This guards the
script's use of the stage APIs; the numbers it prints here mean nothing."""
import ctypes as C
import importlib
import json
import os
import sys
import time

import numpy as np
import pytest

from conftest import ROOT


def _emu(unit, prefixes):
    from nellie_b200 import _cabi
    so = os.path.join(ROOT, "oracle", "_build", f"{unit}.so")
    if not os.path.exists(so):
        pytest.skip(f"{so} not built (python -c 'import __graft_entry__ as g; g.build()')")
    lib = C.CDLL(so)
    for name, (argtypes, restype) in _cabi._SIGS.items():
        if name.startswith(prefixes):
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = argtypes, restype
    return lib


def test_bench_stages_script_runs_on_emulated_kernels(monkeypatch, capsys):
    import scipy.ndimage as ndi
    import torch

    import nellie_b200
    from nellie_b200 import hu_tracking as H, mocap_marking as M, networking as N, phantoms as PH
    from oracle import pipeline as P

    mlib = _emu("markers_host", ("nb200_markers_", "nb200_gauss_"))
    hlib = _emu("hu_host", ("nb200_hu_",))
    nlib = _emu("network_host", ("nb200_network_",))
    cpu = torch.device("cpu")

    class Event:
        def __init__(self, enable_timing=False):
            self.t = None

        def record(self):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return (other.t - self.t) * 1e3

    class FakeFilter:
        def __init__(self, info, **kw):
            pass

        def _get_t(self):
            pass

        def _set_default_sigmas(self):
            pass

        def filter_frame_device(self, raw):
            return ((raw - 100).clamp(min=0) / 400).to(torch.float32)

    class FakeLabel:
        def __init__(self, info, **kw):
            pass

        def label_frame_device(self, frangi, raw=None):
            lab = ndi.label(frangi.numpy() > 0.3, structure=np.ones((3, 3, 3)))[0].astype(np.int32)
            return torch.from_numpy(lab), 0.3

    engine_cls = M.MarkerEngine

    class EmuMarkers(M.Markers):
        def _torch_device(self):
            return cpu

        def _engine_for(self, shape):
            if self._engine is None:
                self._engine = engine_cls(tuple(shape), False, tuple(float(s) for s in self.sigmas), self.z_ratio,
                                          self.max_radius_px, self.peak_min_distance, "cpu", lib=mlib)
            return self._engine

    class EmuNetwork(N.Network):
        @property
        def device(self):
            return cpu

        def _engine(self):
            if self._net is None:
                self._net = N.NetworkEngine(False, self.scaling, "cpu", lib=nlib)
            return self._net

        def _get_pixel_class(self, s):
            return torch.from_numpy(P.network_pixel_class(s.numpy(), False).astype(np.uint8))

        def _get_branch_skel_labels(self, pc):
            return torch.from_numpy(P.network_branch_labels(pc.numpy(), False).astype(np.int32))

        def _remove_connected_label_pixels(self, s):
            return torch.from_numpy(P.network_remove_connected(s.numpy(), False).astype(np.int32))

    hu_cls = H.HuFeatureEngine
    phantom = PH.tubular_phantom
    real_device = torch.device
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", Event)
    monkeypatch.setattr(torch, "device", lambda *a, **k: real_device("cpu"))
    monkeypatch.setattr(nellie_b200, "Filter", FakeFilter)
    monkeypatch.setattr(nellie_b200, "Label", FakeLabel)
    monkeypatch.setattr(M, "Markers", EmuMarkers)
    monkeypatch.setattr(N, "Network", EmuNetwork)
    monkeypatch.setattr(H, "HuFeatureEngine", lambda shape, no_z, dev: hu_cls(shape, no_z, "cpu", lib=hlib))
    monkeypatch.setattr(PH, "tubular_phantom", lambda shape, seed, device=None, **kw: phantom(shape, seed, device="cpu", **kw))
    monkeypatch.setattr(sys, "argv", ["bench_stages.py", "--size", "48", "--reps", "1"])
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    try:
        mod = importlib.import_module("bench_stages")
        importlib.reload(mod)
        mod.main()
    finally:
        sys.path.remove(os.path.join(ROOT, "scripts"))
    rows = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    line = json.loads(rows[-1])
    assert line["frame"] == [48, 48, 48] and line["marker_scales"] == 5
    for key in ("markers.frame", "markers.peaks_all_scales", "markers.peaks_all_scales_two_step", "hu.features_streaming",
                "network.relabel_objects", "network.add_missing"):
        assert key in line["ms"]
    assert line["objects"] >= 1 and line["relabelled_voxels"] > 0
    assert len(line["property_checks"]) >= 8 and all(line["property_checks"].values()), line["property_checks"]
