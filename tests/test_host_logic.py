"""CPU-only checks: host scalar logic vs the oracle, C ABI surface, device math compiled for the host."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sigma_schedule_and_strides_match_oracle():
    from nellie_b200.engine import FilterParams, sample_strides
    from oracle import pipeline as P
    for dim_res, no_z in [({"X": 0.0655, "Y": 0.0655, "Z": 0.25, "T": 4.5}, False),
                          ({"X": 0.1, "Y": 0.1, "Z": 0.1, "T": 1}, False),
                          ({"X": 0.125, "Y": 0.125, "Z": 0.125, "T": 1}, False),
                          ({"X": 0.1, "Y": 0.1, "Z": None, "T": 1}, True),
                          ({"X": 0.5, "Y": 0.5, "Z": 1.0, "T": 1}, False)]:
        for kw in [{}, dict(min_radius_um=0.25, max_radius_um=0.675), dict(min_radius_um=1.0, max_radius_um=1.0)]:
            p = FilterParams(dim_res=dim_res, no_z=no_z, **kw)
            s = P.FrameSpec(dim_res=dim_res, no_z=no_z, **kw)
            assert p.sigma_list() == P.sigma_schedule(s)
            prev = 0.0
            for sg in p.sigma_list():
                assert p.delta_sigma_vec(prev, sg) == P.delta_sigma_vector(s, prev, sg)
                prev = sg
    for shape in [(17, 192, 279), (512, 512, 512), (1024, 1024, 1024), (2048, 2048), (40, 128, 200), (3, 5)]:
        assert sample_strides(shape, 10 ** 6) == P.sample_strides(shape, 10 ** 6)
    # BASELINE config #2 derivation: exactly 4 sigmas 1.0..1.6
    p = FilterParams(dim_res={"X": 0.125, "Y": 0.125, "Z": 0.125}, min_radius_um=0.25, max_radius_um=0.675)
    assert np.allclose(p.sigma_list(), [1.0, 1.2, 1.4, 1.6])


def test_gaussian_taps_match_scipy_kernel():
    from scipy.ndimage import _filters
    from nellie_b200.engine import gaussian_taps
    for sd in [0.441, 0.5, 1.25, 1.908, 2.294, 3.0]:
        w, r = gaussian_taps(sd, 3.0)
        ref = _filters._gaussian_kernel1d(sd, 0, int(3.0 * sd + 0.5))
        assert r == (len(ref) - 1) // 2
        assert np.array_equal(w, ref[r:]) and np.array_equal(w, ref[:r + 1][::-1])


def test_order2_taps_match_scipy_kernel():
    from scipy.ndimage import _filters
    from nellie_b200.engine2d import gaussian_taps_order2
    for sd in [1.25, 1.667, 2.083, 2.5, 2.917]:
        w, r = gaussian_taps_order2(sd, 4.0)
        ref = _filters._gaussian_kernel1d(sd, 2, int(4.0 * sd + 0.5))[::-1]
        assert r == (len(ref) - 1) // 2
        assert np.array_equal(w, ref[r:]) and np.array_equal(w, ref[:r + 1][::-1])


def test_cabi_exports_every_declared_symbol():
    """The shared library loads and exports every function include/nellie_b200.h declares."""
    from nellie_b200 import _cabi, build
    build.build()
    header = open(os.path.join(ROOT, "include", "nellie_b200.h")).read()
    declared = set(re.findall(r"\b(nb200_[a-z0-9_]+)\s*\(", header))
    declared -= {"nb200_vol"}
    lib = C.CDLL(_cabi.lib_path())
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert declared == set(_cabi._SIGS), declared ^ set(_cabi._SIGS)
    assert _cabi.load().nb200_abi_version() == 1


def test_product_has_no_cpu_fallback_and_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "nellie_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert not re.search(r"^\s*(import|from)\s+(scipy|cupy)", src, re.M), f


@pytest.fixture(scope="module")
def hostmath():
    so = os.path.join(ROOT, "oracle", "_build", "devmath_host.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC", "-o", so,
                    os.path.join(ROOT, "oracle", "devmath_host.cpp")], check=True)
    return C.CDLL(so)


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def test_np_log10f_is_bit_exact(hostmath):
    """numpy's float32 log10 (SVML on AVX-512 hosts, which the golden fixtures were produced on) restated in devmath.cuh:
    identical bits on uniform, log-uniform and subnormal inputs.  On a host without AVX-512 numpy falls back to libm
    log10f and this comparison does not apply."""
    import ctypes as C
    if "avx512f" not in open("/proc/cpuinfo").read().lower():
        pytest.skip("numpy uses libm log10f on this host")
    rng = np.random.default_rng(5)
    xs = [rng.uniform(1e-6, 1.0, 500_000).astype(np.float32),
          (np.float32(10.0) ** rng.uniform(-44.5, 38.0, 500_000).astype(np.float32)),
          np.array([1.0, 0.75, 1.5, 1.4999999, 0.99999994, 1.0000001, 1e-45, 1.1754942e-38, 1.17549435e-38, 3.4e38], np.float32)]
    for x in xs:
        x = np.ascontiguousarray(x[np.isfinite(x) & (x > 0)])
        y = np.empty_like(x)
        hostmath.hm_log10f(x.ctypes.data_as(C.POINTER(C.c_float)), y.ctypes.data_as(C.POINTER(C.c_float)), C.c_long(x.size))
        assert np.array_equal(y, np.log10(x)), int((y != np.log10(x)).sum())


def test_np_expf_is_bit_exact(hostmath):
    rng = np.random.default_rng(0)
    x = np.concatenate([-rng.random(400_000) * 30, -rng.random(200_000) * 1e-3,
                        -np.exp(rng.uniform(-40, 5, 400_000)), [-0.0, 0.0, -np.inf, -104.0, -88.0, -87.5, np.nan]]
                       ).astype(np.float32)
    y = np.empty_like(x)
    hostmath.hm_expf(_vp(x), _vp(y), C.c_long(x.size))
    with np.errstate(all="ignore"):
        ref = np.exp(x)
    assert np.array_equal(y.view(np.uint32)[~np.isnan(ref)], ref.view(np.uint32)[~np.isnan(ref)])
    assert np.isnan(y[np.isnan(ref)]).all()
    # the straight-line variant used by the vesselness (x <= 0 or NaN), incl. the gradual-underflow tail
    neg = x[~(x > 0)]
    y2 = np.empty_like(neg)
    hostmath.hm_expf_nonpos(_vp(neg), _vp(y2), C.c_long(neg.size))
    with np.errstate(all="ignore"):
        ref2 = np.exp(neg)
    ok = ~np.isnan(ref2)
    assert np.array_equal(y2.view(np.uint32)[ok], ref2.view(np.uint32)[ok])
    tail = np.linspace(-104.5, -86.0, 200001).astype(np.float32)
    y3 = np.empty_like(tail)
    hostmath.hm_expf_nonpos(_vp(tail), _vp(y3), C.c_long(tail.size))
    assert np.array_equal(y3.view(np.uint32), np.exp(tail).view(np.uint32))


def _eig_ref(h6):
    H = np.empty((len(h6), 3, 3), np.float32)
    H[:, 0, 0], H[:, 0, 1], H[:, 0, 2], H[:, 1, 1], H[:, 1, 2], H[:, 2, 2] = h6.T
    H[:, 1, 0], H[:, 2, 0], H[:, 2, 1] = H[:, 0, 1], H[:, 0, 2], H[:, 1, 2]
    ev = np.linalg.eigvalsh(H)
    return np.take_along_axis(ev, np.argsort(np.abs(ev), axis=1), axis=1)


def test_eig3_matches_eigvalsh_rounded_to_f32(hostmath):
    rng = np.random.default_rng(1)
    n = 300_000
    cases = [rng.standard_normal((n, 6)).astype(np.float32) * 1000]
    t = rng.standard_normal((n, 6)).astype(np.float32)
    t[:, [1, 2, 4]] *= 1e-3
    cases.append(t)
    t = rng.standard_normal((n, 6)).astype(np.float32) * 1e-2
    t[:, [0, 3, 5]] += 5
    cases.append(t)
    z = np.zeros((64, 6), np.float32)
    z[::2, [0, 3, 5]] = 3
    cases.append(z)
    for h6 in cases:
        h6 = np.ascontiguousarray(h6)
        out = np.empty((len(h6), 3), np.float32)
        hostmath.hm_eig3(_vp(h6), _vp(out), C.c_long(len(h6)), C.c_int(2))
        assert np.array_equal(out, _eig_ref(h6))
    # near-degenerate tubes (lambda2 ~ lambda3): the cubic is ill-conditioned there; stay within 1 ulp
    v = rng.standard_normal((n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    lam = rng.uniform(-1000, -10, (n, 1, 1))
    M = lam * (np.eye(3)[None] - v[:, :, None] * v[:, None, :]) + rng.standard_normal((n, 3, 3)) * 0.01
    M = (M + M.transpose(0, 2, 1)) / 2
    h6 = np.ascontiguousarray(np.stack([M[:, 0, 0], M[:, 0, 1], M[:, 0, 2], M[:, 1, 1], M[:, 1, 2], M[:, 2, 2]], 1),
                              dtype=np.float32)
    out = np.empty((n, 3), np.float32)
    hostmath.hm_eig3(_vp(h6), _vp(out), C.c_long(n), C.c_int(2))
    ref = _eig_ref(h6)
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(out.astype(np.float64) - ref) <= 1.2e-7 * scale).all()
    assert (out != ref).any(axis=1).mean() < 2e-3


def test_vesselness_matches_oracle_bitwise(hostmath):
    from oracle import pipeline as P
    rng = np.random.default_rng(2)
    ev = (rng.standard_normal((500_000, 3)) * 50).astype(np.float32)
    ev = np.take_along_axis(ev, np.argsort(np.abs(ev), axis=1), axis=1)
    ev[:10] = 0
    spec = P.FrameSpec(dim_res={"X": 1, "Y": 1, "Z": 1})
    for gamma_sq in [2.0 * 104.2 ** 2, 3.7, 1e-9]:
        ref = P.vesselness(ev.copy(), spec, gamma_sq)
        out = np.empty(len(ev), np.float32)
        hostmath.hm_vesselness3(_vp(np.ascontiguousarray(ev)), _vp(out), C.c_long(len(ev)), C.c_float(0.5),
                                C.c_float(0.5), C.c_float(gamma_sq))
        assert np.array_equal(out, ref)
    spec2 = P.FrameSpec(dim_res={"X": 1, "Y": 1}, no_z=True)
    ev2 = np.ascontiguousarray(ev[:, 1:])
    ref = P.vesselness(ev2.copy(), spec2, 77.0)
    out = np.empty(len(ev2), np.float32)
    hostmath.hm_vesselness2(_vp(ev2), _vp(out), C.c_long(len(ev2)), C.c_float(0.5), C.c_float(77.0))
    assert np.array_equal(out, ref)


def test_provably_zero_tests_never_reject_a_nonzero_response(hostmath):
    """K2 flags a voxel / K3 drops a candidate only when the reference's response is exactly zero: eigenvalues from
    numpy's eigvalsh (f64), rounded to f32, sorted by magnitude, zeroed when l2 > 0 or l3 > 0 (filtering.py:759-761).
    Random matrices at many scales, tube-like and plate-like spectra, and matrices sitting right at the margins."""
    rng = np.random.default_rng(7)
    n = 200_000
    cases = []
    for scale in (1e-6, 1.0, 3e3, 1e8):
        cases.append(rng.standard_normal((n, 6)).astype(np.float32) * np.float32(scale))
    # rotated diagonal spectra: two negative + one small eigenvalue of either sign (tubes), near-degenerate ones
    lam = np.stack([rng.uniform(-1, 0, n), rng.uniform(-1, 0, n), rng.normal(0, 1e-4, n)], 1)
    lam2 = np.stack([rng.normal(0, 1e-5, n), rng.normal(0, 1e-5, n), rng.uniform(-1, 1, n)], 1)
    for L in (lam, lam2, lam * 1e4):
        q, _ = np.linalg.qr(rng.standard_normal((n, 3, 3)))
        H = np.einsum("nij,nj,nkj->nik", q, L, q)
        cases.append(np.stack([H[:, 0, 0], H[:, 0, 1], H[:, 0, 2], H[:, 1, 1], H[:, 1, 2], H[:, 2, 2]], 1).astype(np.float32))
    h6 = np.ascontiguousarray(np.concatenate(cases))
    out = np.empty(len(h6), np.uint8)
    hostmath.hm_zero_tests(_vp(h6), _vp(out), C.c_long(len(h6)))
    ev = _eig_ref(h6)                                   # float32 eigenvalues sorted by |.|
    nonzero = ~((ev[:, 2] > 0) | (ev[:, 1] > 0))        # the reference keeps a response only for these
    flagged = out != 0
    assert flagged.any() and (~flagged).any()
    assert not (flagged & nonzero).any(), int((flagged & nonzero).sum())
    # the tests are worth having: most matrices with a zero response are caught without an eigen-solve
    assert (flagged & ~nonzero).sum() > 0.8 * (~nonzero).sum()


def test_library_sass_uses_tma_and_packed_f32():
    """Static guard (no GPU): the built sm_100a library must still contain the TMA plane loads of the Hessian march
    (UTMALDG.3D + mbarrier transactions), the packed float32x2 arithmetic, and a blur without DFMA contraction."""
    import shutil
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    from nellie_b200 import build
    if not (os.path.exists(cuobjdump) and os.path.exists(build.LIB)):
        pytest.skip("cuobjdump or the built library is missing")
    sass = subprocess.run([cuobjdump, "-sass", build.LIB], capture_output=True, text=True).stdout
    assert "UTMALDG.3D" in sass and "SYNCS.ARRIVE.TRANS64" in sass
    assert "FFMA2" in sass and "FMUL2" in sass
    # per-function check of the blur kernels: DADD / DMUL only
    blocks = sass.split("Function : ")
    gauss = [b for b in blocks if b.startswith("_ZN") and ("gauss_z_vec" in b.split("\n")[0] or "gauss_yx_tile" in b.split("\n")[0])]
    assert gauss, "blur kernels not found in the library"
    for b in gauss:
        assert "DADD" in b and "DMUL" in b and "DFMA" not in b, b.split("\n")[0]
    # the Voronoi stage of the Network relabel reproduces scipy's float64 comparisons (tie-breaking between equidistant
    # seeds): products and sums must stay separate instructions there as well
    voronoi = [b for b in blocks if "voronoi_stage_kernel" in b.split("\n")[0]]
    assert voronoi, "voronoi_stage_kernel not found in the library"
    for b in voronoi:
        assert "DMUL" in b and "DADD" in b and "DFMA" not in b, b.split("\n")[0]


def test_run_ladder_matches_the_reference_rules(monkeypatch):
    """run() of the stage classes (filtering.py:1033-1076 / adaptive_run.py:103-141): the B200 rung raises loudly by
    default; with fallback='reference' an OOM or GPU-unavailable failure moves on to the reference's own classes
    (cpu / high, then cpu / low), any other exception aborts at once."""
    import sys
    import types
    from types import SimpleNamespace

    from nellie_b200 import Filter, Label, adaptive

    assert adaptive.is_oom_error(MemoryError()) and adaptive.is_oom_error(RuntimeError("CUDA out of memory. Tried ..."))
    assert adaptive.is_gpu_unavailable_error(RuntimeError("GPU backend requested but CUDA is not available."))
    assert not adaptive.is_oom_error(ValueError("bad shape")) and not adaptive.is_gpu_unavailable_error(ValueError("x"))

    info = SimpleNamespace(no_t=True, no_z=False, shape=(1, 8, 16, 16), axes="TZYX",
                           dim_res={"X": 0.1, "Y": 0.1, "Z": 0.1, "T": 1.0}, im_path="raw", pipeline_paths={})
    calls = []

    class FakeRef:
        def __init__(self, im_info, device=None, low_memory=False, **kw):
            self.args = (device, low_memory, kw)

        def run(self, *a, **k):
            calls.append(self.args[:2])
            if not self.args[1]:
                raise MemoryError("host out of memory")          # cpu / high-memory fails, cpu / low-memory works

    for modname, cls in (("filtering", "Filter"), ("labelling", "Label")):
        mod = types.ModuleType(f"nellie.segmentation.{modname}")
        setattr(mod, cls, FakeRef)
        monkeypatch.setitem(sys.modules, f"nellie.segmentation.{modname}", mod)
    monkeypatch.setitem(sys.modules, "nellie", types.ModuleType("nellie"))
    monkeypatch.setitem(sys.modules, "nellie.segmentation", types.ModuleType("nellie.segmentation"))

    def no_gpu(*a, **k):
        raise RuntimeError("GPU backend requested but CUDA is not available.")

    for Stage in (Filter, Label):
        calls.clear()
        st = Stage(info, device="b200")                            # default: no ladder below the B200 rung
        monkeypatch.setattr(st, "_run_b200", no_gpu)
        with pytest.raises(RuntimeError, match="CUDA is not available"):
            st.run()
        assert calls == []
        st = Stage(info, device="b200", fallback="reference")
        monkeypatch.setattr(st, "_run_b200", no_gpu)
        st.run()
        assert calls == [("cpu", False), ("cpu", True)]
        calls.clear()
        st = Stage(info, device="b200", fallback="reference")
        monkeypatch.setattr(st, "_run_b200", lambda *a, **k: (_ for _ in ()).throw(ValueError("not a memory problem")))
        with pytest.raises(ValueError):
            st.run()
        assert calls == []


def test_gate_threshold_reproduces_numpy_comparison_semantics():
    """labelling.py:419, :550: ``frame > thresh`` is a float64 comparison for integer frames (and for float32 frames against
    a float64 numpy scalar), a float32 comparison for float32 frames against a Python float; the device compares float32
    values against ONE float32 number, which must therefore be chosen per case."""
    from nellie_b200.labelling import gate_threshold_f32 as g
    rng = np.random.default_rng(0)
    x = np.arange(0, 65536, dtype=np.uint16)
    for t in list(rng.uniform(0, 65535, 300)) + [999.99999999, 1000.0, 1000.00000001, 0.0, 65535.0, -3.5]:
        assert np.array_equal(x > t, x.astype(np.float32) > np.float32(g(t, np.uint16))), t
    xf = rng.uniform(0, 2, 100000).astype(np.float32)
    for t in rng.uniform(0, 2, 300):
        assert np.array_equal(xf > np.float64(t), xf > np.float32(g(np.float64(t), np.float32)))
        assert np.array_equal(xf > float(t), xf > np.float32(g(float(t), np.float32)))
        assert np.array_equal(xf > np.float32(t), xf > np.float32(g(np.float32(t), np.float32)))


def test_stage_engines_refuse_a_host_device():
    """The engine classes of the widened stages take ``lib=`` / ``device=`` so that the tests can inject the host-emulated
    kernel library; without an injected library a host device is an error (no CPU path in the product)."""
    from nellie_b200.hu_tracking import HuFeatureEngine
    from nellie_b200.mocap_marking import MarkerEngine
    from nellie_b200.networking import NetworkEngine
    for make in (lambda: MarkerEngine((4, 5, 6), False, [1.0], 1.0, 5.0, 2, "cpu"),
                 lambda: HuFeatureEngine((4, 5, 6), False, "cpu"),
                 lambda: NetworkEngine(False, (1.0, 1.0, 1.0), "cpu")):
        with pytest.raises(RuntimeError, match="no CPU path"):
            make()
