"""Checks of the Markers kernels shared by the CPU run (kernels host-emulated, tests/test_markers_cpu.py) and the GPU
run (tests/test_zmarkers_gpu.py): the same code drives either library through ``MarkerEngine`` / raw C-ABI calls."""
import ctypes as C
import json
import os

import numpy as np
import scipy.ndimage as ndi
import torch

from conftest import GOLDEN_DIR

MARKER_CASES = ["markers_sample_crop", "markers_phantom3d_iso", "markers_phantom3d_aniso_frangi", "markers_phantom2d",
                "markers_blobs3d", "markers_blobs2d"]
EDT_CASES = [((9, 31, 40), 6.0), ((1, 37, 53), 20.5), ((24, 30, 28), 30.53), ((1, 1, 90), 9.0), ((33, 1, 17), 4.0),
             ((40, 45), 12.25), ((5, 64, 64), 1.0)]
PEAK_CASES = [((12, 26, 30), 0.3), ((10, 21, 33), 0.1), ((48, 50), None)]


def load_marker_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    meta = json.loads(str(z["meta"]))
    if "parent" in z.files:
        p = np.load(os.path.join(GOLDEN_DIR, f"{str(z['parent'])}.npz"))
        raw, labels, frangi = p["raw"], p["labels"], p["frangi"]
    else:
        raw, labels, frangi = z["raw"], z["labels"], None
    return dict(raw=raw, labels=labels, frangi=frangi, meta=meta, marker=z["marker"], distance=z["distance"],
                border=z["border"], sigmas=z["sigmas"])


def marker_spec(meta):
    from oracle.pipeline import MarkerSpec
    return MarkerSpec(dim_res=meta["dim_res"], no_z=meta["no_z"], **meta["kwargs"])


def reference_test_inputs():
    """The inputs of the reference's tests/test_mocap_marking.py:38-47."""
    intensity = np.zeros((9, 9), dtype=np.float32)
    intensity[4, 4] = 10.0
    labels = np.zeros((9, 9), dtype=np.uint8)
    labels[2:7, 2:7] = 1
    dim_res = {"X": 0.2, "Y": 0.2, "Z": None, "T": 1.0}
    return intensity, labels, dim_res


class Backend:
    """A C-ABI library + the torch device its pointers live on (CUDA library / 'cuda', emulated library / 'cpu')."""

    def __init__(self, lib, device):
        self.lib, self.device = lib, torch.device(device)

    def t(self, a, dtype):
        return torch.from_numpy(np.ascontiguousarray(a).astype(dtype)).to(self.device)

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def stream(self):
        if self.device.type != "cuda":
            return None
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def engine(self, shape, no_z, spec, sigmas=None):
        from oracle import pipeline as P
        from nellie_b200.mocap_marking import MarkerEngine
        sig = P.marker_sigmas(spec) if sigmas is None else sigmas
        return MarkerEngine(shape, no_z, sig, spec.z_ratio(), spec.radii_px()[1], spec.peak_min_distance, self.device,
                            lib=self.lib)


def _np(t):
    return t.cpu().numpy()


def check_fixture(be: Backend, name):
    g = load_marker_case(name)
    spec = marker_spec(g["meta"])
    eng = be.engine(g["raw"].shape, g["meta"]["no_z"], spec)
    frangi = be.t(g["frangi"], np.float32) if spec.use_im == "frangi" else None
    marker, distance, border = eng.run_frame(be.t(g["labels"], np.int32), be.t(g["raw"], np.float32), frangi)
    assert np.array_equal(_np(distance), g["distance"])
    assert np.array_equal(_np(border), g["border"])
    assert np.array_equal(_np(marker), g["marker"])


def blob_mask(shape, rng, n_blobs=6, r_max=None):
    grids = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    mask = np.zeros(shape, bool)
    for _ in range(n_blobs):
        c = [rng.uniform(0, s) for s in shape]
        r = rng.uniform(1.0, 0.45 * max(shape) if r_max is None else r_max)
        mask |= sum((g - ci) ** 2 for g, ci in zip(grids, c)) <= r * r
    mask &= rng.random(shape) > 0.002                       # pin holes
    if mask.all():
        mask.flat[0] = False
    return mask


def check_edt_and_border(be: Backend, shape, clamp):
    """Exact EDT incl. the clamp inside thick objects, objects cut by the frame border (the outside is not background),
    one-plane / one-row frames, and the windowed scan when the nearest background voxel is outside the window."""
    lib = be.lib
    rng = np.random.default_rng(int(sum(shape) * 10 + clamp))
    dims = (1,) + tuple(shape) if len(shape) == 2 else tuple(shape)
    mask = blob_mask(shape, rng)
    labels = be.t(mask.astype(np.int32) * 7, np.int32)
    m = be.empty(shape, torch.uint8)
    b = be.empty(shape, torch.uint8)
    assert lib.nb200_markers_mask_border(labels.data_ptr(), *dims, m.data_ptr(), b.data_ptr(), be.stream()) == 0
    assert np.array_equal(_np(m).astype(bool), mask)
    assert np.array_equal(_np(b).astype(bool), ndi.binary_dilation(mask, iterations=1) ^ mask)
    c32 = np.float32(clamp)
    window = max(1, int(np.ceil(c32)))
    scratch = be.empty(2 * mask.size, torch.int16)
    dist = be.empty(shape, torch.float32)
    assert lib.nb200_markers_edt(m.data_ptr(), *dims, window, C.c_float(float(c32)), scratch.data_ptr(), dist.data_ptr(),
                                 be.stream()) == 0
    ref = ndi.distance_transform_edt(mask).astype(np.float32)
    np.minimum(ref, clamp, out=ref)
    assert np.array_equal(_np(dist), ref)
    # a window that does not cover the clamp is refused, not silently wrong
    assert lib.nb200_markers_edt(m.data_ptr(), *dims, max(1, window - 1), C.c_float(float(window) + 0.5),
                                 scratch.data_ptr(), dist.data_ptr(), be.stream()) != 0


def check_peaks_and_nms(be: Backend, shape, z_res):
    """Noise images (many peaks are certain) through the per-scale response, the multi-scale peak search and the
    suppression; an integer intensity image with equal neighbouring scores through the suppression."""
    from oracle import pipeline as P
    lib = be.lib
    rng = np.random.default_rng(len(shape) * 100 + shape[-1])
    no_z = len(shape) == 2
    spec = P.MarkerSpec(dim_res={"X": 0.1, "Y": 0.1, "Z": z_res, "T": 1.0}, no_z=no_z, max_radius_um=0.6, num_sigma=3)
    mask = ndi.binary_dilation(rng.random(shape) < 0.02, iterations=3)
    distance, _ = P.marker_distance(mask, spec)
    base = (distance + rng.random(shape).astype(np.float32) * np.float32(0.25)).astype(np.float32)
    eng = be.engine(shape, no_z, spec)
    eng.mask.copy_(be.t(mask, np.uint8))
    eng.distance.copy_(be.t(distance, np.float32))
    base_t = be.t(base, np.float32)
    peak = _np(eng.peaks(base_t)).astype(bool)
    ref_peak, ref_best = P.marker_peaks(base, mask, distance, spec)
    assert ref_peak.sum() > 10
    assert np.array_equal(peak, ref_peak)
    assert np.array_equal(_np(eng.best), ref_best)
    two_step = _np(eng.peaks(base_t, fused=False)).astype(bool)             # response volume + separate maximum test
    assert np.array_equal(two_step, ref_peak) and np.array_equal(_np(eng.best), ref_best)
    for s, taps in zip(eng.sigmas, eng.taps):
        d0, d1, d2 = eng.laplace_terms(base_t, taps)
        r = be.empty(shape, torch.float32)
        assert lib.nb200_markers_log_response(d0.data_ptr(), d1.data_ptr(), None if d2 is None else d2.data_ptr(), base.size,
                                              C.c_float(float(np.float32(s ** 2))), r.data_ptr(), be.stream()) == 0
        assert np.array_equal(_np(r), P.marker_log_response(base, spec, s))
    for intensity in (rng.random(shape).astype(np.float32) - np.float32(0.1), rng.integers(0, 4, shape).astype(np.uint16)):
        dense = rng.random(shape) < 0.2                       # far more peaks than real data: windows overlap a lot
        kept = _np(eng.suppress(be.t(dense, np.uint8), be.t(intensity, np.float32))).astype(bool)
        assert np.array_equal(kept, P.marker_nms(dense, intensity, spec))
    assert not _np(eng.suppress(be.t(np.zeros(shape), np.uint8), base_t)).any()


def check_mirror_class_replays_reference_tests(make_markers):
    """tests/test_mocap_marking.py of the reference, written against nellie_b200.Markers.  ``make_markers(info, **kw)``
    constructs the stage object."""
    from types import SimpleNamespace
    intensity, labels, dim_res = reference_test_inputs()
    info = SimpleNamespace(no_t=True, no_z=True, shape=(1, 9, 9), axes="TYX", dim_res=dim_res)

    def setup(**kw):
        m = make_markers(info, num_t=1, **kw)
        m.im_memmap, m.label_memmap = intensity[None], labels[None]
        m.shape = m.label_memmap.shape
        m._set_default_sigmas()
        return m

    full = setup(num_sigma=3, low_memory=False)._run_frame_impl(0, low_memory=False)
    low = setup(num_sigma=3, low_memory=True, max_chunk_voxels=20)._run_frame_impl(0, low_memory=True, chunk_voxels=20)
    for a, b in zip(full, low):
        assert np.array_equal(a, b)
    assert full[0].dtype == np.uint8 and full[1].dtype == np.float32 and full[2].dtype == np.uint8
    assert full[0].sum() == 1 and full[0][4, 4] == 1
    m = make_markers(info, num_t=1)
    mask = np.zeros((7, 7), dtype=bool)
    mask[2:5, 2:5] = True
    _, border = m._distance_im(mask)
    assert border.shape == mask.shape and border.dtype == bool and not np.any(border & mask)


def check_mirror_class_helpers_and_run_on_files(make_markers, tmp_path):
    """_local_max_peak / _remove_close_peaks return the reference's coordinate lists; run() writes the three outputs of
    every frame of a T stack through the im_info memmaps, T-sharded over two shard objects as well."""
    from types import SimpleNamespace
    from nellie_b200.imio import StackInfo
    from oracle import pipeline as P
    g = load_marker_case("markers_phantom3d_iso")
    g2 = load_marker_case("markers_blobs3d")
    spec = marker_spec(g["meta"])
    info1 = SimpleNamespace(no_t=True, no_z=False, shape=(1,) + g["raw"].shape, axes="TZYX", dim_res=g["meta"]["dim_res"])
    m = make_markers(info1, num_t=1)
    m._set_default_sigmas()
    assert [float(s) for s in m.sigmas] == g["sigmas"].tolist()
    assert m._get_sigma_vec(2.0) == P.marker_sigma_vec(spec, 2.0)
    mask = g["labels"] > 0
    distance, border = m._distance_im(mask)
    assert np.array_equal(distance, g["distance"]) and np.array_equal(border, g["border"].astype(bool))
    coords = m._local_max_peak(distance, mask, distance)
    ref_peak, _ = P.marker_peaks(distance, mask, distance, spec)
    assert np.array_equal(coords, np.argwhere(ref_peak))
    kept = m._remove_close_peaks(coords, g["raw"])
    assert np.array_equal(kept, np.argwhere(g["marker"]))
    assert m._remove_close_peaks(coords[:0], g["raw"]).size == 0
    # uint16 raw frame and int32 labels straight from the fixture through _run_frame_impl (native-dtype upload)
    mu = make_markers(SimpleNamespace(no_t=True, no_z=False, shape=(1,) + g2["raw"].shape, axes="TZYX",
                                      dim_res=g2["meta"]["dim_res"]), num_t=1)
    mu.im_memmap, mu.label_memmap = g2["raw"][None], g2["labels"][None]
    mu.shape = mu.label_memmap.shape
    got = mu._run_frame_impl(0)
    assert g2["raw"].dtype == np.uint16
    assert all(np.array_equal(a, g2[k]) for a, k in zip(got, ("marker", "distance", "border")))
    # run() on files: a two-frame stack (same shape: crop both inputs)
    shp = tuple(min(a, b) for a, b in zip(g["raw"].shape, g2["raw"].shape))
    sl = tuple(slice(0, s) for s in shp)
    raws = np.stack([g["raw"][sl].astype(np.float32), g2["raw"][sl].astype(np.float32)])
    labs = np.stack([g["labels"][sl], g2["labels"][sl]]).astype(np.int32)
    dim_res = g["meta"]["dim_res"]
    expect = [P.marker_frame(raws[t], labs[t], P.MarkerSpec(dim_res=dim_res)) for t in range(2)]
    for shards in ([None], [(0, 2), (1, 2)]):
        out_dir = tmp_path / ("whole" if shards == [None] else "sharded")
        info = StackInfo.from_array(raws, "TZYX", dim_res, str(out_dir))
        for key in ("im_instance_label", "im_marker", "im_distance", "im_border"):
            info.create_output_path(key)
        info.allocate_memory(info.pipeline_paths["im_instance_label"], dtype="int32", data=labs)
        for shard in shards:
            make_markers(info, t_shard=shard).run()
        for t in range(2):
            assert np.array_equal(info.get_memmap(info.pipeline_paths["im_marker"])[t], expect[t][0])
            assert np.array_equal(info.get_memmap(info.pipeline_paths["im_distance"])[t], expect[t][1])
            assert np.array_equal(info.get_memmap(info.pipeline_paths["im_border"])[t], expect[t][2])


def check_z_sharded_equals_whole_frame(make_markers, tmp_path, world=3):
    """Markers(z_shard=(r, world)) for every r on shared files: slabs + halo through the ordinary single-GPU sequence, own
    planes written — identical to the oracle on the whole frame (tall frame, several objects crossing the slab seams,
    kernels wide enough that every term of the halo matters)."""
    from nellie_b200.imio import StackInfo
    from oracle import pipeline as P
    rng = np.random.default_rng(77)
    shape = (96, 36, 40)
    labels = np.zeros(shape, np.int32)
    labels[blob_mask(shape, rng, n_blobs=14, r_max=9.0)] = 1
    labels[blob_mask(shape, rng, n_blobs=3, r_max=16.0)] = 2
    zz, yy, xx = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    labels[(zz - 33) ** 2 + (yy - 18) ** 2 + (xx - 20) ** 2 <= 17 ** 2] = 3      # thick object across the first seam
    raw = np.round(rng.random(shape) * 50 + 300 * (labels > 0) * rng.random(shape)).astype(np.uint16)
    dim_res = {"X": 0.2, "Y": 0.2, "Z": 0.25, "T": 1.0}
    info = StackInfo.from_array(raw[None], "TZYX", dim_res, str(tmp_path))
    info.allocate_memory(info.pipeline_paths["im_instance_label"], dtype="int32", data=labels[None])
    halos = set()
    for r in range(world):
        m = make_markers(info, z_shard=(r, world))
        m.run()
        halos.add(m.z_halo())
        e0, e1, z0, z1 = m._z_extent(shape[0])
        assert e0 <= z0 < z1 <= e1 and (e1 - e0) < shape[0]            # a real slab, not the whole frame
    assert len(halos) == 1
    ref = P.marker_frame(raw, labels, P.MarkerSpec(dim_res=dim_res))
    assert ref[0].sum() > 20 and ref[1].max() == 10.0                   # markers exist, the distance clamp is active
    assert np.array_equal(info.get_memmap(info.pipeline_paths["im_marker"])[0], ref[0])
    assert np.array_equal(info.get_memmap(info.pipeline_paths["im_distance"])[0], ref[1])
    assert np.array_equal(info.get_memmap(info.pipeline_paths["im_border"])[0], ref[2])


OPTION_CASES = [
    dict(no_z=False, kw=dict(peak_min_distance=0)),
    dict(no_z=False, kw=dict(num_sigma=1, peak_min_distance=1)),
    dict(no_z=False, kw=dict(min_radius_um=0.6, max_radius_um=0.5)),            # non-positive sigma range: one scale
    dict(no_z=False, kw=dict(use_im="frangi", max_radius_um=0.7), z_res=0.1),
    dict(no_z=True, kw=dict(use_im="frangi", num_sigma=2, peak_min_distance=4)),
    dict(no_z=True, kw=dict(max_radius_um=2.0, num_sigma=8)),                   # radii up to 26: per-axis blur path
    dict(no_z=False, kw=dict(), z_res=0.45),                                    # Z sigma 0.22 .. 0.64: radius-1 .. 3 Z kernels
    dict(no_z=False, kw=dict(num_sigma=2), z_res=0.9),                          # Z sigma 0.11: a radius-0 Z kernel (one tap)
]


def check_option_matrix(be: Backend, case):
    """Constructor options of the reference class that change the kernel sequence (scale list, image used for the peaks,
    suppression radius, anisotropy), each against the oracle on one small frame with uint8 intensities (many ties)."""
    from oracle import pipeline as P
    no_z = case["no_z"]
    shape = (44, 52) if no_z else (14, 30, 34)
    rng = np.random.default_rng(len(str(case)))
    labels = np.zeros(shape, np.int32)
    labels[blob_mask(shape, rng, n_blobs=7, r_max=8.0)] = 1
    labels[blob_mask(shape, rng, n_blobs=2, r_max=12.0)] = 5
    raw = rng.integers(0, 6, shape).astype(np.uint8)
    frangi = (rng.random(shape).astype(np.float32) * (labels > 0)).astype(np.float32)
    dim_res = {"X": 0.1, "Y": 0.1, "Z": None if no_z else case.get("z_res", 0.13), "T": 1.0}
    spec = P.MarkerSpec(dim_res=dim_res, no_z=no_z, **case["kw"])
    ref = P.marker_frame(raw, labels, spec, frangi=frangi)
    eng = be.engine(shape, no_z, spec)
    got = eng.run_frame(be.t(labels, np.int32), be.t(raw, np.float32),
                        be.t(frangi, np.float32) if spec.use_im == "frangi" else None)
    for g, r, name in zip(got, ref, ("marker", "distance", "border")):
        assert np.array_equal(_np(g), r), (name, case)
    assert ref[0].sum() > 0
