"""Checks of the Hu feature kernels shared by the CPU run (kernels host-emulated, tests/test_hu_cpu.py) and the GPU run
(tests/test_zz_hu_gpu.py)."""
import ctypes as C
import json
import os

import numpy as np
import torch

from conftest import GOLDEN_DIR

HU_CASES = ["hu_sample_crop", "hu_phantom3d_iso", "hu_phantom2d", "hu_blobs3d"]
# log-Hu entries: the reference's float64 value is compared with a tolerance; entries whose float64 value is itself rounding
# noise (it moves by more than NOISE when the oracle is evaluated in extended precision) are not a parity target
HU_RTOL, HU_ATOL, NOISE = 1e-9, 1e-9, 1e-7


def load_hu_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    src = np.load(os.path.join(GOLDEN_DIR, f"{str(z['source'])}.npz"))
    meta = json.loads(str(z["meta"]))
    if "parent" in src.files:
        p = np.load(os.path.join(GOLDEN_DIR, f"{str(src['parent'])}.npz"))
        raw, frangi = p["raw"], p["frangi"]
    else:
        raw, frangi = src["raw"], z["frangi"]
    d = {k: z[k] for k in z.files}
    d.update(raw=raw, frangi=frangi, distance=src["distance"], marker=src["marker"], meta=meta)
    return d


def scaling_of(meta):
    r = meta["dim_res"]
    return (r["Y"], r["X"]) if meta["no_z"] else (r["Z"], r["Y"], r["X"])


class Backend:
    def __init__(self, lib, device):
        self.lib, self.device = lib, torch.device(device)

    def f32(self, a):
        a = np.asarray(a)
        if a.dtype == np.uint16:
            a = a.astype(np.int32)
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.device).to(torch.float32).contiguous()

    def engine(self, shape, no_z):
        from nellie_b200.hu_tracking import HuFeatureEngine
        return HuFeatureEngine(shape, no_z, self.device, lib=self.lib)


def well_conditioned(intensity, frangi, distance, marker, scaling, no_z, dense):
    """Mask of the log-Hu entries whose float64 reference value is stable under extended-precision evaluation."""
    from oracle import pipeline as P
    f64 = P.hu_frame_features(intensity, frangi, distance, marker, scaling, no_z, dense=True)[3]
    ext = P.hu_frame_features(intensity, frangi, distance, marker, scaling, no_z, dense=True, real=np.longdouble)[3]
    return np.abs(f64 - ext.astype(np.float64)) <= NOISE


def run_engine(be, g, dense):
    from nellie_b200.hu_tracking import integer_bits
    eng = be.engine(g["raw"].shape, g["meta"]["no_z"])
    coords, stats, hu, used_dense = eng.frame_features(
        be.f32(g["raw"]), integer_bits(g["raw"].dtype), be.f32(g["frangi"]), be.f32(g["distance"]),
        torch.from_numpy(g["marker"]).to(be.device), 5e7 if dense else 0)
    assert used_dense == dense
    return coords.cpu().numpy(), stats.cpu().numpy(), hu.cpu().numpy()


def check_fixture(be: Backend, name):
    g = load_hu_case(name)
    ok = well_conditioned(g["raw"], g["frangi"], g["distance"], g["marker"], scaling_of(g["meta"]), g["meta"]["no_z"], True)
    assert ok.mean() > 0.9
    for tag, dense in (("dense", True), ("stream", False)):
        coords, stats, hu = run_engine(be, g, dense)
        assert np.array_equal(coords, g[f"coords_{tag}"])
        assert stats.dtype == np.float32 and np.array_equal(stats, g[f"stats_{tag}"]), tag      # bit for bit
        ref = g[f"hu_{tag}"]
        assert hu.dtype == ref.dtype and hu.shape == ref.shape
        tol = (HU_ATOL + HU_RTOL * np.abs(ref)) if dense else (2e-6 + 1e-6 * np.abs(ref))       # float32 rows when streaming
        assert (np.abs(hu - ref)[ok] <= tol[ok]).all(), (tag, float(np.abs(hu - ref)[ok].max()))


def check_frame_transforms(be: Backend, shape):
    """The frangi transform (numpy's float32 log10, the shift by the minimum of the negatives) and the doubled 3^d maximum
    filter against the oracle, bit for bit, incl. frames without any negative / positive value."""
    from oracle import pipeline as P
    rng = np.random.default_rng(sum(shape))
    no_z = len(shape) == 2
    eng = be.engine(shape, no_z)
    frames = [np.where(rng.random(shape) < 0.3, 10.0 ** rng.uniform(-6, 1.5, shape), 0.0).astype(np.float32),
              np.where(rng.random(shape) < 0.5, rng.uniform(1.0, 90.0, shape), 0.0).astype(np.float32),      # logs >= 0
              np.zeros(shape, np.float32),
              (rng.standard_normal(shape) * 3).astype(np.float32)]                                           # raw negatives
    for f in frames:
        got = eng.transform_frangi(be.f32(f)).cpu().numpy()
        assert np.array_equal(got, P.hu_transform_frangi(f))
    d = np.where(rng.random(shape) < 0.2, rng.uniform(0, 12, shape), 0).astype(np.float32)
    assert np.array_equal(eng.max_distance(be.f32(d)).cpu().numpy(), P.hu_distance_max(d))


def check_stats_and_bounds(be: Backend, shape, dtype):
    """Boxes (clipped at the frame border, radius 0 included) and the mean / variance of the non-zero voxels for dense and
    streaming reductions: float32 pairwise sums and the integer rules (wrapped squares) bit for bit."""
    from oracle import pipeline as P
    rng = np.random.default_rng(sum(shape) + np.dtype(dtype).itemsize)
    no_z = len(shape) == 2
    ndim = len(shape)
    if np.dtype(dtype).kind == "u":
        frame = (rng.integers(0, np.iinfo(dtype).max, shape) * (rng.random(shape) < 0.7)).astype(dtype)
    else:
        frame = (rng.standard_normal(shape) * 100 * (rng.random(shape) < 0.7)).astype(dtype)
    dmax = np.where(rng.random(shape) < 0.5, rng.uniform(0, 7.5, shape), 0).astype(np.float32)
    marker = rng.random(shape) < 0.01
    marker.flat[0] = marker.flat[-1] = True
    coords = np.argwhere(marker)
    eng = be.engine(shape, no_z)
    eng.distance_max.copy_(be.f32(dmax))
    b, max_half = eng.bounds(torch.from_numpy(coords).to(be.device))
    ref_b = P.hu_bounds(coords, dmax, shape)
    got_b = b.cpu().numpy()
    for a in range(ndim):
        assert np.array_equal(got_b[:, 2 * (3 - ndim + a)], ref_b[2 * a]) and np.array_equal(got_b[:, 2 * (3 - ndim + a) + 1], ref_b[2 * a + 1])
    side = int(np.ceil(dmax[marker].max())) * 2 + 1
    assert 2 * max_half + 1 == side
    from nellie_b200.hu_tracking import integer_bits
    bits = integer_bits(frame.dtype)
    for cube in (side, 0):
        got = eng.roi_stats(be.f32(frame), b, cube, bits).cpu().numpy()
        rois = []
        for i in range(len(coords)):
            sl = tuple(slice(int(ref_b[2 * a][i]), int(ref_b[2 * a + 1][i])) for a in range(ndim))
            roi = frame[sl]
            if cube:
                pad = np.zeros((cube,) * ndim, frame.dtype)
                pad[tuple(slice(0, s) for s in roi.shape)] = roi
                roi = pad
            rois.append(P.hu_mean_and_variance(roi[None])[0])
        assert np.array_equal(got, np.stack(rois)), (cube, dtype)


def check_markers_then_hu_on_files(Markers, HuMomentFeatures, tmp_path):
    """Markers.run() writes im_marker / im_distance; HuMomentFeatures reads them back through the im_info memmaps and
    returns the reference's features for every frame (the oracle evaluates the same files)."""
    from nellie_b200.imio import StackInfo
    from oracle import pipeline as P
    g = load_hu_case("hu_phantom3d_iso")
    src = np.load(f"{GOLDEN_DIR}/phantom3d_iso.npz")
    raws = np.stack([g["raw"], g["raw"][::-1].copy()])
    labs = np.stack([src["labels"], src["labels"][::-1].copy()]).astype(np.int32)
    frs = np.stack([g["frangi"], g["frangi"][::-1].copy()])
    dim_res = g["meta"]["dim_res"]
    info = StackInfo.from_array(raws, "TZYX", dim_res, str(tmp_path))
    for key in ("im_instance_label", "im_preprocessed", "im_marker", "im_distance", "im_border"):
        info.create_output_path(key)
    info.allocate_memory(info.pipeline_paths["im_instance_label"], dtype="int32", data=labs)
    info.allocate_memory(info.pipeline_paths["im_preprocessed"], dtype="float32", data=frs)
    Markers(info).run()
    hu = HuMomentFeatures(info, dense_limit=int(5e7))
    hu._get_t()
    hu._allocate_memory()
    for t in range(2):
        ff = hu._get_frame_features(t)
        marker = info.get_memmap(info.pipeline_paths["im_marker"])[t]
        distance = info.get_memmap(info.pipeline_paths["im_distance"])[t]
        c, p, s, h = P.hu_frame_features(raws[t], frs[t], distance, marker, scaling_of(g["meta"]), False, dense=True)
        assert np.array_equal(ff.coords_voxel, c) and np.array_equal(ff.coords_phys, p)
        assert np.array_equal(ff.stats, s)
        assert np.allclose(ff.hu, h, rtol=1e-7, atol=1e-7)
