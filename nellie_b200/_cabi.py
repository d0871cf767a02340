"""ctypes binding of ``csrc/libnellie_b200.so`` (the C ABI declared in ``include/nellie_b200.h``).

There is deliberately no CPU fallback: if the shared library is missing or a call fails, the
product path raises.  (``nellie_b200.build.build()`` compiles it in-tree with nvcc.)
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

HIST_WORDS = 3 + 256
SP_WORDS = 20
HS_WORDS = 8
SP_GAMMA, SP_GAMMA_SQ, SP_FROB_THR, SP_FROB_CUT, SP_MAX_ABS, SP_SKIP, SP_STATUS, SP_TRI, SP_OTSU, SP_UNSAFE = range(10)
SP_FROBSQ_MIN, SP_AMBIG, SP_FS_LO, SP_FS_HI, SP_ZT_C, SP_DELTA = range(10, 16)
HS_MAX_ABS_BITS, HS_MAX_FROBSQ_BITS, HS_MIN_NZ_COMPL, HS_MAX_G_BITS, HS_APPROX_MAX_BITS, HS_FALLBACK = range(6)
DIV_IEEE, DIV_FAST, DIV_POW2 = 0, 1, 2
TF_NONE, TF_DIV, TF_LOG10 = 0, 1, 2
SELECT_WORDS = 2048 + 8
STATE_WORDS = HIST_WORDS + HS_WORDS
FOLD_MINMAX, FOLD_BINS = 0, 1

ERR_OOM = -3


class Vol(C.Structure):
    """``nb200_vol``: Z-window geometry shared by the 3-D kernels."""
    _fields_ = [("nz_buf", C.c_int), ("ny", C.c_int), ("nx", C.c_int), ("zc0", C.c_int), ("zc1", C.c_int),
                ("zg_off", C.c_int), ("nz_glob", C.c_int)]

    @classmethod
    def whole(cls, nz, ny, nx):
        return cls(nz, ny, nx, 0, nz, 0, nz)


class NellieB200Error(RuntimeError):
    pass


class NellieB200OutOfMemory(MemoryError):
    """Raised for NB200_ERR_OOM so the reference's retry ladder (adaptive_run.is_oom_error) sees it."""


_p = C.c_void_p
_ll = C.c_longlong
_SIGS = {
    "nb200_abi_version": ([], C.c_int),
    "nb200_last_error": ([], C.c_char_p),
    "nb200_sm_count": ([], C.c_int),
    "nb200_gauss_axis": ([_p, _p, C.POINTER(Vol), C.c_int, C.POINTER(C.c_double), C.c_int, _p], C.c_int),
    "nb200_gauss_yx": ([_p, _p, C.POINTER(Vol), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, _p], C.c_int),
    "nb200_lattice_sample": ([_p, C.POINTER(Vol), C.c_int, C.c_int, C.c_int, _p, _p], C.c_int),
    "nb200_strided_sample": ([_p, _ll, _ll, _ll, _p, C.c_float, _p, _p], C.c_int),
    "nb200_hist_reset": ([_p, _p], C.c_int),
    "nb200_hist_minmax": ([_p, _ll, C.c_int, _p, _p, _p], C.c_int),
    "nb200_hist_bins": ([_p, _ll, C.c_int, _p, _p, _p], C.c_int),
    "nb200_finalize_gamma": ([_p, _p, _p], C.c_int),
    "nb200_finalize_frob": ([_p, _p, C.c_double, C.c_double, _p, _p], C.c_int),
    "nb200_finalize_max_abs": ([_p, _p, _p], C.c_int),
    "nb200_finalize_label_threshold": ([_p, C.c_int, _p, _p], C.c_int),
    "nb200_hist_bins_f64": ([_p, _ll, _p, _p], C.c_int),
    "nb200_finalize_otsu_f64": ([_p, _p, _p], C.c_int),
    "nb200_hessian_stats": ([_p, C.POINTER(Vol), C.POINTER(C.c_float), C.c_int, _p, C.c_int, C.c_int, C.c_int, _p, _p, _p],
                            C.c_int),
    "nb200_hessian_stats_code": ([_p, C.POINTER(Vol), C.POINTER(C.c_float), C.c_int, _p, C.c_int, C.c_int, C.c_int, _p, _p,
                                  _p, _p], C.c_int),
    "nb200_hessian_stats_redo": ([_p, C.POINTER(Vol), C.POINTER(C.c_float), C.c_int, _p, C.c_int, C.c_int, C.c_int, _p, _p,
                                  _p, _p], C.c_int),
    "nb200_frangi_sparse": ([_p, _p, _p, C.POINTER(Vol), C.POINTER(C.c_float), C.c_int, C.c_float, C.c_float, _p, _p, _ll, _p, _p],
                            C.c_int),
    "nb200_frangi_sparse_gated": ([_p, _p, _p, C.POINTER(Vol), C.POINTER(C.c_float), C.c_int, C.c_float, C.c_float, _p, _p, _ll, _p,
                                   _p], C.c_int),
    "nb200_hessian_stats_ambig": ([_p, C.POINTER(Vol), C.POINTER(C.c_float), C.c_int, _p, C.c_int, C.c_int, C.c_int, _p, _p, _p],
                                  C.c_int),
    "nb200_hessian_fast_workspace_bytes": ([], C.c_size_t),
    "nb200_hessian_stats_fast": ([_p, C.POINTER(Vol), C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int, _p, _p, _p, _p],
                                 C.c_int),
    "nb200_finalize_frob_fast": ([_p, _p, C.c_double, C.c_double, C.c_double, C.c_int, _p, _p], C.c_int),
    "nb200_finalize_frob_resolve": ([_p, _p, _p], C.c_int),
    "nb200_pixel_class": ([_p, C.c_int, C.c_int, C.c_int, _p, _p], C.c_int),
    "nb200_branch_labels": ([_p, C.c_int, C.c_int, C.c_int, _p, _p, _p, _p], C.c_int),
    "nb200_remove_connected_label_pixels": ([_p, C.c_int, C.c_int, C.c_int, _p, _p], C.c_int),
    "nb200_markers_mask_border": ([_p, C.c_int, C.c_int, C.c_int, _p, _p, _p], C.c_int),
    "nb200_markers_edt": ([_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _p, _p, _p], C.c_int),
    "nb200_markers_log_response": ([_p, _p, _p, _ll, C.c_float, _p, _p], C.c_int),
    "nb200_markers_peak_update": ([_p, _p, _p, C.c_int, C.c_int, C.c_int, _p, _p, _p], C.c_int),
    "nb200_markers_peak_update_fused": ([_p, _p, _p, C.c_float, _p, _p, C.c_int, C.c_int, C.c_int, _p, _p, _p], C.c_int),
    "nb200_markers_nms": ([_p, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p], C.c_int),
    "nb200_hu_frangi_transform": ([_p, _ll, _p, _p, _p], C.c_int),
    "nb200_hu_distance_max": ([_p, C.c_int, C.c_int, C.c_int, _p, _p], C.c_int),
    "nb200_hu_bounds": ([_p, _ll, C.c_int, _p, C.c_int, C.c_int, C.c_int, _p, _p, _p], C.c_int),
    "nb200_hu_roi_stats": ([_p, C.c_int, C.c_int, C.c_int, _p, _ll, C.c_int, C.c_int, C.c_int, _p, _p], C.c_int),
    "nb200_hu_log_moments": ([_p, C.c_int, C.c_int, C.c_int, _p, _ll, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p, _p], C.c_int),
    "nb200_network_add_missing": ([_p, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p, _p], C.c_int),
    "nb200_network_skeleton_labels": ([_p, _p, _ll, _p, _p], C.c_int),
    "nb200_network_object_boxes": ([_p, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p, _p], C.c_int),
    "nb200_network_relabel": ([_p, _p, C.c_int, C.c_int, C.c_int, _p, _ll, _ll, _p, C.POINTER(_ll), C.POINTER(C.c_double),
                               _p, _p, _p, _p, _p], C.c_int),
    "nb200_histn_workspace_bytes": ([C.c_int], C.c_size_t),
    "nb200_histn_threshold": ([_p, _ll, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p, _p, _p], C.c_int),
    "nb200_remove_edges": ([_p, C.c_int, C.c_int, C.c_int, C.c_int, _p], C.c_int),
    "nb200_fold_records": ([_p, C.c_int, C.c_int, _p, _p], C.c_int),
    "nb200_fold_records_n": ([_p, C.c_int, C.c_int, C.c_int, _p, _p], C.c_int),
    "nb200_frangi_fast": ([_p, _p, C.POINTER(Vol), C.POINTER(C.c_float), C.c_int, C.c_float, C.c_float, _p, _p, _p], C.c_int),
    "nb200_hessian_components": ([_p, C.POINTER(Vol), C.POINTER(C.c_float), C.c_int, _p, _p], C.c_int),
    "nb200_divisor_mode": ([C.c_float, C.POINTER(C.c_int), _p], C.c_int),
    "nb200_hstats_reset": ([_p, _p], C.c_int),
    "nb200_frangi_accumulate": ([_p, _p, C.POINTER(Vol), C.POINTER(C.c_float), C.c_int, C.c_float, C.c_float, _p, _p],
                                C.c_int),
    "nb200_frangi_accumulate_2d": ([_p, _p, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_float, _p, _p], C.c_int),
    "nb200_hessian_stats_2d": ([_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int, _p, _p, _p], C.c_int),
    "nb200_percentile": ([_p, _ll, C.c_double, _p, _p, _p], C.c_int),
    "nb200_finalize_opening": ([_p, _p, C.POINTER(Vol), _p, _p], C.c_int),
    "nb200_finalize_opening_2d": ([_p, _p, C.c_int, C.c_int, _p, _p], C.c_int),
    "nb200_log2d_accumulate": ([_p, _p, _p, C.c_float, C.c_int, _ll, _p, _p], C.c_int),
    "nb200_log2d_combine": ([_p, _p, _ll, _p, _p, _p], C.c_int),
    "nb200_label_workspace_bytes": ([C.c_int, C.c_int, C.c_int], C.c_size_t),
    "nb200_label_frame": ([_p, _p, C.c_int, C.c_float, _p, C.c_int, C.c_int, C.c_int, _ll, C.c_int, _p, _p, _p, _p],
                          C.c_int),
    "nb200_ccl_label": ([_p, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p, _p, _p], C.c_int),
}

_lib = None


def lib_path() -> str:
    # NB200_LIB: an alternative build of the same library (kernel experiments: scripts/build_variant.sh)
    return os.environ.get("NB200_LIB") or _build.LIB


def load():
    """Load (once) and return the ctypes library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise NellieB200Error(
            f"{path} is missing: build it with `python -m nellie_b200.build` (nvcc, sm_100a). "
            "nellie_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (argtypes, restype) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc == 0:
        return
    msg = load().nb200_last_error().decode("utf-8", "replace")
    if rc == ERR_OOM:
        raise NellieB200OutOfMemory(f"CUDA out of memory in {what}: {msg}")
    raise NellieB200Error(f"{what} failed ({rc}): {msg}")


def call(name: str, *args):
    check(getattr(load(), name)(*args), name)
