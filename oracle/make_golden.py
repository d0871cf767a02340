"""Generate ``tests/golden/*.npz`` by EXECUTING the unmodified reference (build container only).

    python -m oracle.make_golden            # rewrites tests/golden/

Each fixture stores the input frame, the physical spec, and what the reference's
``Filter`` / ``Label`` produce for it through their own methods
(filtering.py:910 ``_run_frame`` + :952 ``_mask_volume`` guarded as in :1014-1018;
labelling.py:511 ``_compute_frame_thresholds`` + :538 ``_run_frame_full_volume``),
plus per-sigma scalars recorded by wrapping (not modifying) the reference methods.
The GPU box has no ``/root/reference``: tests there read only these files.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def _recording_filter(Filter):
    class Rec(Filter):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.rec_gamma, self.rec_frob_thr, self.rec_max_abs, self.rec_mask_frac = [], [], [], []

        def _calculate_gamma(self, g):
            v = super()._calculate_gamma(g)
            self.rec_gamma.append(float(v))
            return v

        def _get_frob_mask(self, frob):
            # recover the threshold the reference derives, by its own helpers
            if self.frob_thresh is None:
                from nellie.utils.gpu_functions import otsu_threshold, triangle_threshold
                pos = self._subsample_for_thresholds(frob)
                thr = 0.0 if pos.size == 0 else float(min(triangle_threshold(pos, xp=np),
                                                          otsu_threshold(pos, xp=np)[0]))
            else:
                thr = float(self.frob_thresh)
            self.rec_frob_thr.append(thr)
            m = super()._get_frob_mask(frob)
            self.rec_mask_frac.append(float(m.mean()))
            return m

    return Rec


def run_reference(raw, dim_res, no_z, filter_kwargs=None, label_kwargs=None, sigmas=None, run_mask=True):
    Filter, Label = ref_shim.load()
    Rec = _recording_filter(Filter)
    info = ref_shim.im_info_for(raw.shape, dim_res, no_z)
    f = Rec(info, device="cpu", **(filter_kwargs or {}))
    f._get_t()
    f._set_default_sigmas()
    if sigmas is not None:
        f.sigmas = [float(s) for s in sigmas]
        f.halo = f._compute_halo()
    f.im_memmap = raw[None].copy()  # the reference mutates float32 inputs
    pre = f._run_frame(0) if run_mask else f._run_frame(0, mask=False)
    pre = np.array(pre, copy=True)
    fin = f._mask_volume(pre.copy()) if float(np.sum(pre)) > 0.0 else pre.copy()
    lab = Label(info, device="cpu", **(label_kwargs or {}))
    lab.num_t = 1
    it, ft = lab._compute_frame_thresholds(raw, fin)
    labels = lab._run_frame_full_volume(0, raw, fin, it, ft)
    return dict(
        sigmas=np.asarray(f.sigmas, dtype=np.float64),
        gamma=np.asarray(f.rec_gamma, dtype=np.float64),
        frob_thr=np.asarray(f.rec_frob_thr, dtype=np.float64),
        mask_frac=np.asarray(f.rec_mask_frac, dtype=np.float64),
        frangi_pre=pre.astype(np.float32), frangi=fin.astype(np.float32),
        intensity_thresh=np.float64(np.nan if it is None else it),
        frangi_thresh=np.float64(np.nan if ft is None else ft),
        min_area=np.int64(lab.min_area_pixels), labels=labels.astype(np.int32),
    )


def save_case(name, raw, dim_res, no_z, **kw):
    out = run_reference(raw, dim_res, no_z, **kw)
    meta = dict(dim_res=dim_res, no_z=bool(no_z),
                filter_kwargs=kw.get("filter_kwargs") or {}, label_kwargs=kw.get("label_kwargs") or {},
                explicit_sigmas=None if kw.get("sigmas") is None else [float(s) for s in kw["sigmas"]],
                run_mask=bool(kw.get("run_mask", True)))
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    np.savez_compressed(path, raw=raw, meta=np.asarray(json.dumps(meta)), **out)
    nz = int((out["frangi"] > 0).sum())
    print(f"{name}: shape={raw.shape} sigmas={np.round(out['sigmas'], 3).tolist()} nonzero={nz} "
          f"labels={int(out['labels'].max())} thr={float(out['frangi_thresh']):.6g} "
          f"size={os.path.getsize(path) / 1024:.0f} KiB")


def label_only_cases():
    """Direct ``Label._get_labels`` cases: hand-built responses with holes, specks and seams."""
    _, Label = ref_shim.load()
    rng = np.random.default_rng(11)
    cases = {}
    # 3-D: smooth random field thresholded into blobs, with carved cavities and specks
    import scipy.ndimage as ndi
    field = ndi.gaussian_filter(rng.standard_normal((20, 44, 48)), 2.0).astype(np.float32)
    field = (field - field.min()) / (field.max() - field.min())
    field[rng.random(field.shape) < 0.002] = 1.0          # specks (removed by the area filter)
    field[8:11, 20:24, 20:24] = 0.0                        # cavity candidates
    cases["label3d"] = (field.astype(np.float32), {"X": 0.2, "Y": 0.2, "Z": 0.3, "T": 1.0}, False, 0.55)
    f2 = ndi.gaussian_filter(rng.standard_normal((72, 80)), 2.5).astype(np.float32)
    f2 = (f2 - f2.min()) / (f2.max() - f2.min())
    f2[rng.random(f2.shape) < 0.01] = 1.0
    cases["label2d"] = (f2.astype(np.float32), {"X": 0.1, "Y": 0.1, "Z": None, "T": 1.0}, True, 0.6)
    for name, (fr, dim_res, no_z, thr) in cases.items():
        info = ref_shim.im_info_for(fr.shape, dim_res, no_z)
        lab = Label(info, device="cpu")
        lab.num_t = 1
        labels = lab._run_frame_full_volume(0, fr, fr, None, thr)
        meta = dict(dim_res=dim_res, no_z=no_z, frangi_thresh=thr)
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, frangi=fr, labels=labels.astype(np.int32), min_area=np.int64(lab.min_area_pixels),
                            meta=np.asarray(json.dumps(meta)))
        print(f"{name}: shape={fr.shape} labels={int(labels.max())} min_area={lab.min_area_pixels}")


def network_cases():
    """The Network stage's array kernels (networking.py:261-296, :669-680, :758-797) executed on synthetic skeletons:
    random 26-connected walks labelled by object, with crossings (junctions), touching objects and border voxels."""
    from types import SimpleNamespace
    import scipy.ndimage as ndi
    ref_shim.load()
    from nellie.segmentation.networking import Network
    rng = np.random.default_rng(17)
    for name, shape, no_z in (("network3d", (24, 40, 48), False), ("network2d", (64, 72), True)):
        skel = np.zeros(shape, np.int32)
        nd = len(shape)
        for obj in range(1, 13):
            p = np.array([rng.integers(0, s) for s in shape])
            for _ in range(int(rng.integers(20, 90))):
                skel[tuple(p)] = obj
                p = np.clip(p + rng.integers(-1, 2, nd), 0, np.array(shape) - 1)
        me = SimpleNamespace(im_info=SimpleNamespace(no_z=no_z), low_memory=False, xp=np, ndi=ndi)
        cleaned = Network._remove_connected_label_pixels_impl(me, skel, np, ndi)
        pixel_class = Network._get_pixel_class_impl(me, skel, np, ndi)
        branch = Network._get_branch_skel_labels(me, pixel_class, force_cpu=True)
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, skel=skel, cleaned=np.asarray(cleaned).astype(np.int32),
                            pixel_class=np.asarray(pixel_class).astype(np.uint8), branch=np.asarray(branch).astype(np.int32),
                            meta=np.asarray(json.dumps(dict(no_z=no_z))))
        print(f"{name}: shape={shape} skeleton voxels={int((skel > 0).sum())} removed={int(((cleaned == 0) & (skel > 0)).sum())} "
              f"junctions={int((pixel_class == 4).sum())} branches={int(branch.max())}")


def _blob_case(shape, seed, n_blobs, r_max):
    """Labelled solid balls / discs (radius 2..r_max) + an intensity frame with a few bright spots per object."""
    rng = np.random.default_rng(seed)
    grids = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    labels = np.zeros(shape, np.int32)
    for k in range(1, n_blobs + 1):
        c = [rng.uniform(0, s) for s in shape]
        r = rng.uniform(2.0, r_max)
        d2 = sum((g - ci) ** 2 for g, ci in zip(grids, c))
        labels[d2 <= r * r] = k
    raw = (rng.random(shape) * 40.0 + 200.0 * (labels > 0) * rng.random(shape)).astype(np.float32)
    raw = np.round(raw).astype(np.uint16)          # integer intensities: ties between neighbouring peaks do occur
    return raw, labels


def marker_cases():
    """Markers stage (mocap_marking.py:648-703 ``_run_frame_impl``, full-volume branch) executed by the unmodified
    reference on the inputs of existing fixtures (raw + the reference's own labels / frangi) and on labelled blobs that
    are thicker than 2 * max_radius_px (the distance clamp of :447 is active)."""
    ref_shim.load()
    from nellie.segmentation.mocap_marking import Markers

    def run(name, raw, labels, dim_res, no_z, frangi=None, **kw):
        info = ref_shim.im_info_for(raw.shape, dim_res, no_z)
        m = Markers(info, num_t=1, device="cpu", **kw)
        m.im_memmap = raw[None]
        m.label_memmap = labels[None]
        m.im_frangi_memmap = None if frangi is None else frangi[None]
        m.shape = m.label_memmap.shape
        m._set_default_sigmas()
        marker, distance, border = m._run_frame_impl(0, low_memory=False)
        meta = dict(dim_res=dim_res, no_z=no_z, kwargs=kw)
        return dict(marker=np.asarray(marker, np.uint8), distance=np.asarray(distance, np.float32),
                    border=np.asarray(border, np.uint8), sigmas=np.asarray(m.sigmas, np.float64),
                    meta=np.asarray(json.dumps(meta)))

    def from_fixture(name, parent, **kw):
        z = np.load(os.path.join(GOLDEN_DIR, f"{parent}.npz"))
        meta = json.loads(str(z["meta"]))
        out = run(name, z["raw"], z["labels"], meta["dim_res"], meta["no_z"],
                  frangi=z["frangi"] if kw.get("use_im") == "frangi" else None, **kw)
        out["parent"] = np.asarray(parent)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"{name}.npz"), **out)
        print(f"{name}: markers={int(out['marker'].sum())} border={int(out['border'].sum())} "
              f"max distance={float(out['distance'].max()):.4f} sigmas={out['sigmas'].round(3).tolist()}")

    from_fixture("markers_sample_crop", "sample_crop")
    from_fixture("markers_phantom3d_iso", "phantom3d_iso")
    from_fixture("markers_phantom3d_aniso_frangi", "phantom3d_aniso", use_im="frangi", num_sigma=3)
    from_fixture("markers_phantom2d", "phantom2d", peak_min_distance=3)
    for name, shape, no_z, dim_res, kw in (
            ("markers_blobs3d", (30, 56, 64), False, {"X": 0.2, "Y": 0.2, "Z": 0.3, "T": 1.0}, {}),
            ("markers_blobs2d", (120, 140), True, {"X": 0.2, "Y": 0.2, "Z": None, "T": 1.0}, {"max_radius_um": 1.5})):
        raw, labels = _blob_case(shape, 31 + len(shape), 12, 15.0 if not no_z else 22.0)
        out = run(name, raw, labels, dim_res, no_z, **kw)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"{name}.npz"), raw=raw, labels=labels, **out)
        print(f"{name}: markers={int(out['marker'].sum())} border={int(out['border'].sum())} "
              f"max distance={float(out['distance'].max()):.4f} clamp={2.0 * (kw.get('max_radius_um', 1.0) / 0.2)}")


def hu_cases():
    """HuMomentTracking._get_frame_features_impl (hu_tracking.py:585-680) executed by the unmodified reference on the
    marker / distance frames of the Markers fixtures, in its dense mode and (low_memory=True) its streaming mode."""
    ref_shim.load()
    from nellie.tracking.hu_tracking import HuMomentTracking
    for name, src in (("hu_sample_crop", "markers_sample_crop"), ("hu_phantom3d_iso", "markers_phantom3d_iso"),
                      ("hu_phantom2d", "markers_phantom2d"), ("hu_blobs3d", "markers_blobs3d")):
        z = np.load(os.path.join(GOLDEN_DIR, f"{src}.npz"))
        meta = json.loads(str(z["meta"]))
        extra = {}
        if "parent" in z.files:
            p = np.load(os.path.join(GOLDEN_DIR, f"{str(z['parent'])}.npz"))
            raw, frangi = p["raw"], p["frangi"]
        else:
            raw = z["raw"]
            frangi = (z["distance"] * np.float32(0.013)).astype(np.float32)      # stand-in response, all values < 1
            extra["frangi"] = frangi
        info = ref_shim.im_info_for(raw.shape, meta["dim_res"], meta["no_z"])
        info.no_t = False
        info.shape = (2,) + raw.shape
        out = dict(source=np.asarray(src), meta=np.asarray(json.dumps(dict(dim_res=meta["dim_res"], no_z=meta["no_z"]))), **extra)
        for tag, low in (("dense", False), ("stream", True)):
            tr = HuMomentTracking(info, num_t=2, device="cpu", low_memory=low)
            tr.im_memmap, tr.im_frangi_memmap = raw[None], frangi[None]
            tr.im_distance_memmap, tr.im_marker_memmap = z["distance"][None], z["marker"][None]
            tr.shape = (1,) + raw.shape
            ff = tr._get_frame_features_impl(0)
            out[f"coords_{tag}"] = np.asarray(ff.coords_voxel)
            out[f"phys_{tag}"] = np.asarray(ff.coords_phys)
            out[f"stats_{tag}"] = np.asarray(ff.stats)
            out[f"hu_{tag}"] = np.asarray(ff.hu)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"{name}.npz"), **out)
        print(f"{name}: markers={out['coords_dense'].shape[0]} stats={out['stats_dense'].dtype} hu dense={out['hu_dense'].dtype} "
              f"stream={out['hu_stream'].dtype} dense==stream stats: {np.array_equal(out['stats_dense'], out['stats_stream'])} "
              f"hu max|dense-stream|={np.abs(out['hu_dense'] - out['hu_stream']).max():.3g}")


def standin_skeleton(mask):
    """Deterministic thin subset of a mask standing in for skimage.morphology.skeletonize (scikit-image is not in this
    image): the ridge of the Euclidean distance transform.  Every step AFTER the thinning is what the fixtures pin."""
    import scipy.ndimage as ndi
    mask = np.asarray(mask, bool)
    d = ndi.distance_transform_edt(mask)
    return (d == ndi.maximum_filter(d, size=3)) & mask


def standin_skeleton_half(mask):
    """The same ridge with nothing in the lower-X half of the frame: objects that live there end up without a skeleton
    voxel, which is the case networking.py:315-392 (_add_missing_skeleton_labels) exists for."""
    sk = standin_skeleton(mask)
    sk[..., : mask.shape[-1] // 2] = False
    return sk


def network_frame_cases():
    """Network._run_frame_backend (networking.py:825-851) executed by the unmodified reference on the labels / Frangi frames
    of existing fixtures, with ``skimage.morphology.skeletonize`` replaced by ``standin_skeleton`` (the one substitution;
    scikit-image is absent here), plus the two host steps on their own (_add_missing_skeleton_labels, _relabel_objects)."""
    import sys
    ref_shim.load()
    sys.modules["skimage"].morphology = sys.modules["skimage.morphology"]
    from nellie.segmentation import networking as ref_net
    ref_net.morph = sys.modules["skimage.morphology"]
    for name, parent in (("network_frame_sample_crop", "sample_crop"), ("network_frame_phantom3d_aniso", "phantom3d_aniso"),
                         ("network_frame_phantom2d", "phantom2d"), ("network_frame_cfg3", "phantom3d_cfg3"),
                         ("network_frame_cfg3_half", "phantom3d_cfg3"), ("network_frame_phantom2d_half", "phantom2d")):
        sys.modules["skimage.morphology"].skeletonize = standin_skeleton_half if name.endswith("_half") else standin_skeleton
        z = np.load(os.path.join(GOLDEN_DIR, f"{parent}.npz"))
        meta = json.loads(str(z["meta"]))
        from oracle.pipeline import tie_free
        labels, frangi = z["labels"], tie_free(z["frangi"])       # unique maxima per object: see tie_free
        dim_res = dict(meta["dim_res"])
        if meta["no_z"]:
            dim_res["Z"] = dim_res.get("Z") or 1.0
        info = ref_shim.im_info_for(labels.shape, dim_res, meta["no_z"])
        net = ref_net.Network(info, num_t=1, device="cpu")
        net.label_memmap, net.im_frangi_memmap = labels[None], frangi[None]
        net.shape = net.label_memmap.shape
        skel0 = net._skeletonize(labels)
        cleaned = net._remove_connected_label_pixels(skel0, force_cpu=True)
        added = net._add_missing_skeleton_labels(np.array(cleaned, copy=True), labels, frangi)
        branch, pixel_class, relabelled = net._run_frame_backend(0)
        out = dict(parent=np.asarray(parent), skeleton=np.asarray(skel0 > 0), cleaned=np.asarray(cleaned, np.int32),
                   added=np.asarray(added, np.int32), branch=np.asarray(branch, np.int32),
                   pixel_class=np.asarray(pixel_class, np.uint8), relabelled=np.asarray(relabelled, np.uint32),
                   scaling=np.asarray(net.scaling, np.float64),
                   meta=np.asarray(json.dumps(dict(dim_res=dim_res, no_z=meta["no_z"]))))
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"{name}.npz"), **out)
        n_missing = int(np.setdiff1d(np.unique(labels), np.unique(cleaned)).size)
        print(f"{name}: objects={int(labels.max())} skeleton voxels={int((skel0 > 0).sum())} objects without skeleton={n_missing} "
              f"branches={int(branch.max())} junction voxels={int((pixel_class == 4).sum())} relabelled voxels={int((relabelled > 0).sum())}")


def label_nbins_cases():
    """Label with histogram_nbins != 256 (labelling.py:23-35, :440-465), executed by the unmodified reference on the raw /
    Frangi frames of existing fixtures: thresholds and labels only (the inputs stay in the parent fixture)."""
    _, Label = ref_shim.load()
    for name, parent, kw in (("label_nbins64_iso", "phantom3d_iso", dict(histogram_nbins=64)),
                             ("label_nbins1000_iso", "phantom3d_iso", dict(histogram_nbins=1000)),
                             ("label_nbins100_sample", "sample_crop", dict(histogram_nbins=100)),
                             ("label_nbins33_2d", "phantom2d", dict(histogram_nbins=33)),
                             ("label_nbins64_u16_otsu", "phantom3d_u16_otsu", dict(histogram_nbins=64, otsu_thresh_intensity=True)),
                             ("label_nbins500_f32_otsu", "phantom3d_f32_otsu", dict(histogram_nbins=500, otsu_thresh_intensity=True))):
        z = np.load(os.path.join(GOLDEN_DIR, f"{parent}.npz"))
        meta = json.loads(str(z["meta"]))
        info = ref_shim.im_info_for(z["raw"].shape, meta["dim_res"], meta["no_z"])
        lab = Label(info, device="cpu", **kw)
        lab.num_t = 1
        it, ft = lab._compute_frame_thresholds(z["raw"], z["frangi"])
        labels = lab._run_frame_full_volume(0, z["raw"], z["frangi"], it, ft)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"{name}.npz"), parent=np.asarray(parent),
                            meta=np.asarray(json.dumps(dict(label_kwargs=kw))),
                            intensity_thresh=np.float64(np.nan if it is None else it),
                            frangi_thresh=np.float64(np.nan if ft is None else ft), labels=np.asarray(labels, np.int32))
        print(f"{name}: it={it} ft={ft:.8g} labels={int(labels.max())} (256 bins: ft={float(z['frangi_thresh']):.8g})")


def main():
    from nellie_b200.phantoms import tubular_phantom_np
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    # (1) crop of the reference's own sample file, frame 0, its OME pixel sizes (BASELINE config #1)
    vol = ref_shim.read_sample_frame(0)
    save_case("sample_crop", np.ascontiguousarray(vol[:, 56:152, 96:208]), ref_shim.SAMPLE_DIM_RES, False)
    # (2) isotropic 3-D phantom, default 5 sigmas
    iso = {"X": 0.1, "Y": 0.1, "Z": 0.1, "T": 1.0}
    save_case("phantom3d_iso", tubular_phantom_np((28, 60, 68), seed=21, n_tubes=6), iso, False)
    # (3) anisotropic 3-D phantom with explicit sigma list (the config-#3 mechanism)
    aniso = {"X": 0.1, "Y": 0.1, "Z": 0.2, "T": 1.0}
    save_case("phantom3d_aniso", tubular_phantom_np((20, 52, 76), seed=22, n_tubes=5), aniso, False,
              sigmas=[1.0, 1.4, 1.8, 2.2])
    # (4) 2-D phantom: 2x2 closed-form eigenvalues + LoG blobness path (BASELINE config #4 shape class)
    d2 = {"X": 0.1, "Y": 0.1, "Z": None, "T": 1.0}
    save_case("phantom2d", tubular_phantom_np((150, 170), seed=23, n_tubes=7), d2, True)
    # (5) > 1e6 voxels so the threshold lattice has strides > 1 (filtering.py:328-340); stored as uint8
    big = tubular_phantom_np((40, 128, 200), seed=24, n_tubes=30)
    big8 = np.clip(np.round(big / 2.0), 0, 255).astype(np.uint8)
    save_case("phantom3d_strided", big8, iso, False, sigmas=[1.0, 1.6])
    # (6) power-of-two pixel size (BASELINE config #2: dim_res 0.125, radii 0.25 .. 0.675 um -> sigmas 1.0 .. 1.6):
    #     every finite-difference divisor is a power of two (division mode POW2 of the CUDA kernels)
    p2 = {"X": 0.125, "Y": 0.125, "Z": 0.125, "T": 1.0}
    save_case("phantom3d_pow2", tubular_phantom_np((32, 64, 72), seed=25, n_tubes=7), p2, False,
              filter_kwargs={"min_radius_um": 0.25, "max_radius_um": 0.675})
    # (7) BASELINE config #3 in small: isotropic 0.1 um, the bench's explicit six sigmas, > 1.2e6 voxels (lattice
    #     strides > 1, Z radii 3-4); stored as uint8 like (5)
    c3 = tubular_phantom_np((48, 160, 160), seed=26, n_tubes=40)
    c3 = np.clip(np.round(c3 / 2.0), 0, 255).astype(np.uint8)
    save_case("phantom3d_cfg3", c3, iso, False, sigmas=[1.0, 1.4, 1.8, 2.2, 2.6, 3.0])
    # (8) Filter._run_frame(t, mask=False) (filtering.py:910-933): no Frobenius gate
    save_case("phantom3d_nomask", tubular_phantom_np((20, 40, 48), seed=27, n_tubes=4), iso, False, run_mask=False)
    # (9) / (10) Label's intensity gate (labelling.py:457-465, :511-556) with otsu_thresh_intensity=True: a uint16 frame
    #     (numpy bins integer samples with float64 edges and compares raw > thresh in float64) and the same data as
    #     float32 (float32 edges, float32 comparison)
    u16 = np.clip(np.round(tubular_phantom_np((24, 56, 64), seed=28, n_tubes=6) * 37.0), 0, 65535).astype(np.uint16)
    save_case("phantom3d_u16_otsu", u16, aniso, False, label_kwargs={"otsu_thresh_intensity": True})
    save_case("phantom3d_f32_otsu", (u16.astype(np.float32) * np.float32(0.731)), aniso, False,
              label_kwargs={"otsu_thresh_intensity": True})
    label_only_cases()
    network_cases()
    marker_cases()
    hu_cases()
    network_frame_cases()
    label_nbins_cases()


if __name__ == "__main__":
    main()
