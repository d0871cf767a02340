"""Quick GPU check of the Markers kernels (the checks of tests/markers_checks.py, most valuable first), logging every
result as it is obtained:  python scripts/markers_quick.py  ->  gpurun_out/markers_quick.log"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "markers_quick.log"), "a")
T0 = time.time()


def log(msg):
    LOG.write(f"[{time.time() - T0:6.1f}s] {msg}\n")
    LOG.flush()
    os.fsync(LOG.fileno())
    print(msg, flush=True)


log("start")
import torch  # noqa: E402

log(f"torch imported, cuda={torch.cuda.is_available()}")
import markers_checks as K  # noqa: E402
from nellie_b200 import _cabi  # noqa: E402

be = K.Backend(_cabi.load(), "cuda")
log("library loaded")


def run(name, fn, *args):
    try:
        fn(*args)
        torch.cuda.synchronize()
        log(f"PASS {name}")
    except Exception as exc:  # noqa: BLE001
        log(f"FAIL {name}: {type(exc).__name__}: {str(exc)[:300]}")


if "--mirror" in sys.argv:
    # third call: the mirror classes themselves on the GPU (host arrays in, files out) and the blur kernels at the large
    # radii / derivative taps the Markers scales use
    import pathlib
    import tempfile

    import numpy as np
    import scipy.ndimage as ndi

    import hu_checks as HK
    from nellie_b200 import Markers
    from nellie_b200.hu_tracking import HuMomentFeatures

    tmp = pathlib.Path(tempfile.mkdtemp())
    run("mirror: reference's own marker tests", K.check_mirror_class_replays_reference_tests, Markers)
    run("mirror: helpers + run() on files (+ T shards)", K.check_mirror_class_helpers_and_run_on_files, Markers, tmp / "a")
    run("mirror: Markers.run() -> HuMomentFeatures on files", HK.check_markers_then_hu_on_files, Markers, HuMomentFeatures, tmp / "b")

    def gauss_case(shape, sigma):
        import ctypes as C
        from nellie_b200.engine import gaussian_taps
        from nellie_b200.engine2d import gaussian_taps_order2
        rng = np.random.default_rng(int(sigma * 100) + shape[2])
        x = (rng.random(shape, dtype=np.float32) * 20.0).astype(np.float32)
        a = torch.from_numpy(x).cuda()
        b = torch.empty_like(a)
        v = _cabi.Vol.whole(*shape)
        for order, (w, r) in ((0, gaussian_taps(sigma, 4.0)), (2, gaussian_taps_order2(sigma, 4.0))):
            for axis in range(3):
                _cabi.call("nb200_gauss_axis", C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.byref(v), axis,
                           w.ctypes.data_as(C.POINTER(C.c_double)), r, be.stream())
                ref = ndi.gaussian_filter1d(x, sigma, axis=axis, order=order, truncate=4.0, mode="reflect")
                assert np.array_equal(b.cpu().numpy(), ref), (order, axis, r)

    for sigma in (2.2, 2.87, 3.4, 4.4):
        run(f"gauss_axis radius {int(4 * sigma + 0.5)} orders 0/2", gauss_case, (21, 45, 68), sigma)
    log("done")
    sys.exit(0)

if "--hu" in sys.argv:
    # second call of the round: the Hu feature kernels first, then timings of both stages on a 256^3 frame
    import json

    import numpy as np

    import hu_checks as HK

    hb = HK.Backend(_cabi.load(), "cuda")
    # every log-Hu entry of these fixtures is well conditioned (tests/test_hu_cpu.py evaluates the extended-precision
    # criterion); skipping that CPU work keeps this call inside the last GPU seconds of the round
    HK.well_conditioned = lambda i, f, d, marker, sc, no_z, dense: np.ones((int((marker > 0).sum()), 6 if no_z else 18), bool)
    for case in HK.HU_CASES:
        run(f"hu fixture {case}", HK.check_fixture, hb, case)
    for dt in (np.float32, np.uint16, np.uint8):
        run(f"hu stats {dt.__name__}", HK.check_stats_and_bounds, hb, (14, 30, 33), dt)
    run("hu transforms", HK.check_frame_transforms, hb, (7, 20, 33))
    try:
        from nellie_b200.hu_tracking import HuFeatureEngine
        from nellie_b200.mocap_marking import MarkerEngine
        shape = (256, 256, 256)
        g = torch.Generator(device="cuda").manual_seed(1)
        noise = torch.rand(shape, device="cuda", generator=g)
        import torch.nn.functional as F
        sm = F.avg_pool3d(noise[None, None], 9, 1, 4)[0, 0]
        labels = (sm > sm.flatten()[::97].quantile(0.90)).to(torch.int32)
        raw = (noise * 1000).round()
        eng = MarkerEngine(shape, False, [1.0, 1.4667, 1.9333, 2.4, 2.8667], 1.0, 10.0, 2, "cuda")
        times = {}

        def timed(name, fn, reps=3):
            fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                out = fn()
            b.record()
            torch.cuda.synchronize()
            times[name] = a.elapsed_time(b) / reps
            return out

        timed("markers.distance_and_border", lambda: eng.distance_and_border(labels))
        timed("markers.peaks(5 scales)", lambda: eng.peaks(eng.distance))
        timed("markers.suppress", lambda: eng.suppress(eng.peak, raw))
        timed("markers.frame", lambda: eng.run_frame(labels, raw))
        hu = HuFeatureEngine(shape, False, "cuda")
        frangi = (sm * labels).contiguous()
        timed("hu.transform_frangi", lambda: hu.transform_frangi(frangi))
        timed("hu.max_distance", lambda: hu.max_distance(eng.distance))
        timed("hu.frame_features(dense)", lambda: hu.frame_features(raw, 16, frangi, eng.distance, eng.marker, 5e9))
        timed("hu.frame_features(stream)", lambda: hu.frame_features(raw, 16, frangi, eng.distance, eng.marker, 0))
        info = dict(shape=shape, foreground=float((labels > 0).float().mean()), markers=int(eng.marker.sum()),
                    ms={k: round(v, 4) for k, v in times.items()})
        log("TIMING " + json.dumps(info))
    except Exception as exc:  # noqa: BLE001
        log(f"FAIL timing: {type(exc).__name__}: {str(exc)[:300]}")
    log("done")
    sys.exit(0)

for case in ["markers_phantom3d_iso", "markers_sample_crop", "markers_blobs3d", "markers_phantom2d",
             "markers_phantom3d_aniso_frangi", "markers_blobs2d"]:
    run(f"fixture {case}", K.check_fixture, be, case)
for shape, clamp in K.EDT_CASES:
    run(f"edt {shape} {clamp}", K.check_edt_and_border, be, shape, clamp)
for shape, z_res in K.PEAK_CASES:
    run(f"peaks {shape} {z_res}", K.check_peaks_and_nms, be, shape, z_res)
log("done")
