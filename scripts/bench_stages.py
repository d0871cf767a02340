"""Timings of the stages either side of Filter -> Label (Network, Markers, HuMoment features) on one synthetic
frame, device-resident, CUDA events.  Prints ONE JSON line.  `bench.py` runs this in a subprocess after its own timed region
(N = 1 only) and embeds the line as `hierarchy_stages`; it can be run by hand as well:

    python scripts/bench_stages.py [--size 384] [--reps 5]

The frame goes through the repo's own Filter and Label first (so the label field is what the stages see in production).
Skeletonization is a scikit-image host call in the reference and here (not timed, not available in this image): the Network
steps run on a stand-in skeleton, the ridge of the distance transform, computed with torch ops.
Algorithmic bytes per voxel: mask + border + distance transform 19, one Markers scale 43 (five blur passes of
8 B + the fused response / maximum test).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=384)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--device", type=int, default=int(os.environ.get("LOCAL_RANK", "0")))
    args = ap.parse_args()

    import torch
    import torch.nn.functional as F

    from types import SimpleNamespace

    from nellie_b200 import Filter, Label
    from nellie_b200.hu_tracking import HuFeatureEngine
    from nellie_b200.mocap_marking import Markers
    from nellie_b200.networking import Network
    from nellie_b200.phantoms import tubular_phantom

    torch.cuda.set_device(args.device)
    dev = torch.device("cuda", args.device)
    n = args.size
    shape = (n, n, n)
    dim_res = {"X": 0.1, "Y": 0.1, "Z": 0.1, "T": 1.0}
    info = SimpleNamespace(no_t=True, no_z=False, shape=(1,) + shape, axes="TZYX", dim_res=dim_res)
    raw = tubular_phantom(shape, seed=5000, device=dev)
    times = {}

    def timed(name, fn, reps=args.reps):
        out = fn()                                        # warm-up (allocations, first launch)
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            out = fn()
        b.record()
        torch.cuda.synchronize(dev)
        times[name] = a.elapsed_time(b) / reps
        return out

    filt = Filter(info, device="b200", cuda_device=dev)
    filt._get_t()
    filt._set_default_sigmas()
    frangi = timed("filter.frame", lambda: filt.filter_frame_device(raw), reps=2).clone()
    lab = Label(info, device="b200", cuda_device=dev)
    labels = timed("label.frame", lambda: lab.label_frame_device(frangi, raw)[0], reps=2).clone()
    n_objects = int(labels.max().item())

    # ---- Markers (mocap_marking.py:648-703) ----
    mk = Markers(info, device="b200", cuda_device=dev)
    mk._set_default_sigmas()
    meng = mk._engine_for(shape)
    raw_f = raw.contiguous()
    timed("markers.mask_border_distance", lambda: meng.distance_and_border(labels))
    checks = {}
    timed("markers.peaks_all_scales_two_step", lambda: meng.peaks(meng.distance, fused=False))
    two_step = (meng.peak.clone(), meng.best.clone())
    timed("markers.peaks_all_scales", lambda: meng.peaks(meng.distance))
    checks["peaks_fused_equals_two_step"] = bool(torch.equal(meng.peak, two_step[0]) and torch.equal(meng.best, two_step[1]))
    timed("markers.suppress", lambda: meng.suppress(meng.peak, raw_f))
    timed("markers.frame", lambda: meng.run_frame(labels, raw_f))
    n_markers = int(meng.marker.sum().item())
    n_scales = len(meng.sigmas)
    # size-independent properties of the reference's definitions (mocap_marking.py:419-450, :595-606)
    fg = labels > 0
    checks["markers_are_peaks_inside_the_mask"] = bool(((meng.marker == 0) | ((meng.peak != 0) & fg)).all().item())
    checks["border_is_outside_the_mask_and_touches_it"] = bool(
        ((meng.border == 0) | (~fg & (F.max_pool3d(fg[None, None].float(), 3, 1, 1)[0, 0] > 0))).all().item())
    checks["distance_positive_exactly_on_the_mask"] = bool(torch.equal(meng.distance > 0, fg))
    checks["distance_at_most_the_clamp"] = bool((meng.distance <= meng.clamp).all().item())

    # ---- HuMoment features (hu_tracking.py:585-680) ----
    hu = HuFeatureEngine(shape, False, dev)
    timed("hu.transform_frangi", lambda: hu.transform_frangi(frangi))
    timed("hu.max_distance", lambda: hu.max_distance(meng.distance))
    res = timed("hu.features_streaming", lambda: hu.frame_features(raw_f, 0, frangi, meng.distance, meng.marker, 0), reps=2)
    if n_markers <= 20000:          # the dense mode walks the zero-padded ROI cube per marker: parity feature, slow by design
        timed("hu.features_dense", lambda: hu.frame_features(raw_f, 0, frangi, meng.distance, meng.marker, 1e18), reps=1)

    # ---- Network, device steps (networking.py:825-851) on a stand-in skeleton ----
    net = Network(info, device="b200", cuda_device=dev, skeletonize=lambda m: m)
    neng = net._engine()
    d = meng.distance
    ridge = (d == F.max_pool3d(d[None, None], 3, 1, 1)[0, 0]) & (labels > 0)
    skel0 = (labels * ridge).to(torch.int32).contiguous()
    cleaned = timed("network.remove_connected", lambda: net._remove_connected_label_pixels(skel0))
    added = timed("network.add_missing", lambda: neng.add_missing(cleaned.clone(), labels, frangi, n_objects))
    skel_pre = timed("network.skeleton_labels", lambda: neng.skeleton_labels(added, labels))
    pixel_class = timed("network.pixel_class", lambda: net._get_pixel_class(skel_pre))
    branch = timed("network.branch_labels", lambda: net._get_branch_skel_labels(pixel_class))
    relabelled = timed("network.relabel_objects", lambda: neng.relabel(branch, labels, n_objects), reps=2)
    # networking.py:485-577: seeds keep their own branch label; only object voxels of seeded objects are relabelled
    seeds = (branch > 0) & fg
    checks["relabel_keeps_seed_labels"] = bool(torch.equal(relabelled[seeds], branch[seeds]))
    checks["relabel_stays_inside_objects"] = bool(((relabelled == 0) | fg).all().item())
    checks["branch_labels_only_on_the_skeleton"] = bool(((branch == 0) | (skel_pre > 0)).all().item())

    vox = float(n) ** 3
    peak = 6530.3
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:  # noqa: BLE001
        pass

    def frac(bytes_per_voxel, ms):
        return bytes_per_voxel * vox / (ms * 1e-3) / 1e9 / peak

    line = {
        "what": "stages around Filter -> Label on one device-resident frame, CUDA events, mean of %d runs" % args.reps,
        "frame": list(shape), "foreground": float((labels > 0).float().mean().item()), "objects": n_objects,
        "markers": n_markers, "marker_scales": n_scales, "branches": int(branch.max().item()),
        "skeleton": "stand-in: ridge of the distance transform (the thinning is a scikit-image host call in the reference)",
        "relabelled_voxels": int((relabelled != 0).sum().item()), "relabel_crop_voxels_over_frame": neng.crop_voxels / vox,
        "property_checks": checks,
        "ms": {k: round(v, 4) for k, v in times.items()},
        "voxels_per_s": {"markers.frame": vox / (times["markers.frame"] * 1e-3),
                         "network.device_steps": vox / (1e-3 * sum(v for k, v in times.items() if k.startswith("network.")))},
        "hbm_frac_of_peak": {"markers.mask_border_distance (19 B/voxel)": frac(19.0, times["markers.mask_border_distance"]),
                             "markers.peaks_all_scales (43 B/voxel and scale)": frac(43.0 * n_scales, times["markers.peaks_all_scales"])},
        "peak_gbs": peak,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
