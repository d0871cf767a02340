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


for case in ["markers_phantom3d_iso", "markers_sample_crop", "markers_blobs3d", "markers_phantom2d",
             "markers_phantom3d_aniso_frangi", "markers_blobs2d"]:
    run(f"fixture {case}", K.check_fixture, be, case)
for shape, clamp in K.EDT_CASES:
    run(f"edt {shape} {clamp}", K.check_edt_and_border, be, shape, clamp)
for shape, z_res in K.PEAK_CASES:
    run(f"peaks {shape} {z_res}", K.check_peaks_and_nms, be, shape, z_res)
log("done")
