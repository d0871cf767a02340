"""GPU parity of the Label path: golden vectors of the executed reference, the reference's own unit
cases (tests/test_labelling.py), and scipy.ndimage on random masks through the raw C ABI."""
import ctypes as C
from types import SimpleNamespace

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _info(shape, dim_res, no_z):
    return SimpleNamespace(no_t=True, no_z=no_z, shape=(1,) + tuple(shape), axes="TYX" if no_z else "TZYX",
                           dim_res=dim_res)


@pytest.mark.parametrize("name", ["label3d", "label2d"])
def test_label_only_golden(name):
    from nellie_b200 import Label
    g = load_golden(name)
    lab = Label(_info(g["frangi"].shape, g["meta"]["dim_res"], g["meta"]["no_z"]), device="b200")
    lab.num_t = 1
    assert lab.min_area_pixels == int(g["min_area"])
    labels = lab._run_frame_full_volume(0, g["frangi"], g["frangi"], None, g["meta"]["frangi_thresh"])
    assert labels.dtype == np.int32
    assert np.array_equal(labels, g["labels"])


@pytest.mark.parametrize("name", ["sample_crop", "phantom3d_iso", "phantom3d_aniso", "phantom2d", "phantom3d_strided", "phantom3d_pow2",
                                  "phantom3d_cfg3", "phantom3d_nomask"])
def test_label_on_reference_frangi(name):
    from nellie_b200 import Label
    g = load_golden(name)
    lab = Label(_info(g["raw"].shape, g["meta"]["dim_res"], g["meta"]["no_z"]), device="b200")
    lab.num_t = 1
    assert lab.min_area_pixels == int(g["min_area"])
    it, ft = lab._compute_frame_thresholds(g["raw"], g["frangi"])
    assert it is None
    # the device transforms the samples with numpy's own float32 log10 (SVML restated in devmath.cuh) and the host
    # applies 10 ** np.float32(.) to the two scalars like labelling.py:452-455: the threshold is the reference's, bit
    # for bit, and so are the labels derived from it
    assert ft == float(g["frangi_thresh"]), (ft, float(g["frangi_thresh"]))
    labels = lab._run_frame_full_volume(0, g["raw"], g["frangi"], None, ft)
    assert np.array_equal(labels, g["labels"])


@pytest.mark.parametrize("name", ["phantom3d_u16_otsu", "phantom3d_f32_otsu"])
def test_label_intensity_otsu_on_reference_frangi(name):
    """Label(otsu_thresh_intensity=True) against the executed reference (labelling.py:457-465, :511-556): the gate
    threshold of a uint16 frame is numpy's FLOAT64 bin centre (integer samples are binned with float64 edges) and the gate
    compares in float64; a float32 frame gets float32 edges, centre and comparison.  Both must come out bit for bit, and
    with them the gated Frangi threshold and the labels."""
    import torch
    from nellie_b200 import Label
    g = load_golden(name)
    lab = Label(_info(g["raw"].shape, g["meta"]["dim_res"], g["meta"]["no_z"]), device="b200", otsu_thresh_intensity=True)
    lab.num_t = 1
    it, ft = lab._compute_frame_thresholds(g["raw"], g["frangi"])
    assert type(it) is (np.float64 if g["raw"].dtype.kind in "iu" else np.float32)
    assert float(it) == float(g["intensity_thresh"]), (it, float(g["intensity_thresh"]))
    assert ft == float(g["frangi_thresh"]), (ft, float(g["frangi_thresh"]))
    labels = lab._run_frame_full_volume(0, g["raw"], g["frangi"], it, ft)
    assert np.array_equal(labels, g["labels"])
    # the device-resident path (Label.run's per-frame call) with the frame in its native dtype
    raw_t = torch.from_numpy(g["raw"].astype(np.int32) if g["raw"].dtype == np.uint16 else g["raw"]).cuda()
    dev_labels, ft2 = lab.label_frame_device(torch.from_numpy(g["frangi"]).cuda(), raw_t)
    assert ft2 == float(g["frangi_thresh"])
    assert np.array_equal(dev_labels.cpu().numpy(), g["labels"])


def _tiny_info():
    return _info((5, 5), {"X": 1.0, "Y": 1.0, "Z": None, "T": 1.0}, True)


def test_reference_unit_case_label_ids_reset_per_frame():
    # reference tests/test_labelling.py:25-53
    from nellie_b200 import Label
    lab = Label(_tiny_info(), num_t=2, device="b200")
    original = np.zeros((5, 5), np.float32)
    original[1:4, 1:4] = 1.0
    frangi = original.copy()
    for t in range(2):
        labels = lab._run_frame_full_volume(t, original, frangi, intensity_thresh=None, frangi_thresh=0.5)
        assert labels is not None and labels.max() == 1 and set(np.unique(labels)) <= {0, 1}


def test_reference_unit_case_masking_does_not_mutate_inputs():
    # reference tests/test_labelling.py:56-77
    from nellie_b200 import Label
    lab = Label(_tiny_info(), num_t=1, device="b200")
    original = np.zeros((5, 5), np.float32)
    original[1:4, 1:4] = 1.0
    frangi = original.copy()
    o0, f0 = original.copy(), frangi.copy()
    labels = lab._run_frame_full_volume(0, original, frangi, intensity_thresh=0.5, frangi_thresh=0.5)
    assert labels is not None
    assert np.array_equal(original, o0) and np.array_equal(frangi, f0)


def _ccl(mask, full):
    import torch
    from nellie_b200 import _cabi
    lib = _cabi.load()
    m = torch.from_numpy(mask.astype(np.uint8)).cuda()
    nz, ny, nx = (1,) * (3 - mask.ndim) + mask.shape
    ws = torch.empty(lib.nb200_label_workspace_bytes(nz, ny, nx), dtype=torch.uint8, device="cuda")
    out = torch.empty(mask.shape, dtype=torch.int32, device="cuda")
    n = torch.zeros(1, dtype=torch.int64, device="cuda")
    _cabi.call("nb200_ccl_label", C.c_void_p(m.data_ptr()), nz, ny, nx, int(full), C.c_void_p(out.data_ptr()),
               C.c_void_p(ws.data_ptr()), C.c_void_p(n.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    return out.cpu().numpy(), int(n.item())


@pytest.mark.parametrize("shape,density,full", [
    ((37, 45, 70), 0.05, True), ((37, 45, 70), 0.25, True), ((16, 33, 129), 0.6, True), ((37, 45, 70), 0.3, False),
    ((200, 333), 0.4, True), ((200, 333), 0.55, False), ((1, 1, 5), 0.5, True), ((3, 3, 3), 1.0, True),
    ((64, 128, 256), 0.02, True), ((64, 128, 256), 0.9, False),
])
def test_ccl_matches_scipy_label(shape, density, full):
    import scipy.ndimage as ndi
    rng = np.random.default_rng(hash((shape, density, full)) % (2 ** 32))
    mask = rng.random(shape) < density
    structure = np.ones((3,) * mask.ndim, bool) if full else None
    ref, nref = ndi.label(mask, structure=structure)
    got, ngot = _ccl(mask, full)
    assert ngot == nref
    assert np.array_equal(got, ref.astype(np.int32))


def _seam_cases():
    """Structures aimed at the tile scheme (32 x 8 x 8 tiles in 3-D, 32 x 64 in 2-D): diagonals that cross tile corners,
    blocks that contain whole tiles (the all-full fast path), thin sheets on tile faces."""
    m = np.zeros((40, 50, 150), bool)
    t = np.arange(40)
    m[t, t, t] = True                                   # body diagonal through tile corners
    m[t, t + 5, 149 - t] = True                         # the other way in x
    m[t, 49 - t, 3 * t + 20] = False
    m[t[:-1], 49 - t[:-1], 31 + (t[:-1] % 2)] = True    # zig-zag over the x = 31 | 32 strip boundary
    m[8:32, 16:40, 32:128] = True                       # 3 x 3 x 3 whole tiles
    m[16:24, 24:32, 64:96] = False                      # one whole tile of background inside
    m[7, :, 40:100] = True                              # sheet just below a Z face
    m[:, 7, 100:140] ^= True                            # sheet just below a Y face
    m2 = np.zeros((130, 100), bool)
    u = np.arange(100)
    m2[u, u] = True
    m2[u + 30, 99 - u] = True
    m2[60:70, :] = True
    m2[64, 20:80] = False
    return [m, ~m, m2, ~m2]


@pytest.mark.parametrize("case", range(4))
@pytest.mark.parametrize("full", [True, False])
def test_ccl_tile_seams(case, full):
    import scipy.ndimage as ndi
    mask = _seam_cases()[case]
    ref, nref = ndi.label(mask, structure=np.ones((3,) * mask.ndim, bool) if full else None)
    got, ngot = _ccl(mask, full)
    assert ngot == nref
    assert np.array_equal(got, ref.astype(np.int32))


def test_label_frame_with_whole_tiles_of_foreground_and_cavities():
    """Solid blocks larger than a tile (the all-full / all-empty tile fast paths), a cavity that is a whole background tile,
    a cavity open to the frame border, a block flush with the frame border."""
    from nellie_b200 import Label
    from oracle import pipeline as P
    rng = np.random.default_rng(5)
    field = (0.2 * rng.random((48, 64, 160))).astype(np.float32)
    field[4:44, 6:60, 10:150] = 0.9
    field[16:24, 24:32, 64:96] = 0.1             # exactly one tile: enclosed
    field[30:33, 40:44, 0:70] = 0.1              # channel open to x = 0
    field[0:10, 0:12, 100:160] = 0.9             # flush with three faces of the frame
    field[2:6, 3:7, 110:130] = 0.1               # enclosed cavity next to the border
    field[rng.random(field.shape) < 0.002] = 1.0
    dim_res = {"X": 0.2, "Y": 0.2, "Z": 0.25, "T": 1.0}
    spec = P.FrameSpec(dim_res=dim_res, no_z=False)
    stages = {}
    ref = P.label_frame(field, spec, 0.5, stages=stages)
    assert (stages["filled"] != (field > 0.5)).any()
    lab = Label(_info(field.shape, dim_res, False), device="b200")
    lab.num_t = 1
    got = lab._run_frame_full_volume(0, field, field, None, 0.5)
    assert np.array_equal(got, ref)


def test_ccl_empty_and_full():
    got, n = _ccl(np.zeros((9, 10, 11), bool), True)
    assert n == 0 and not got.any()
    got, n = _ccl(np.ones((9, 10, 67), bool), True)
    assert n == 1 and (got == 1).all()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_label_frame_matches_oracle_random_blobs(seed):
    """Whole _get_labels chain (fill holes, size filter, smoothing, relabel) against the oracle."""
    import scipy.ndimage as ndi
    from nellie_b200 import Label
    from oracle import pipeline as P
    rng = np.random.default_rng(seed)
    field = ndi.gaussian_filter(rng.standard_normal((48, 80, 96)), 2.0 + seed).astype(np.float32)
    field = (field - field.min()) / (field.max() - field.min())
    field[rng.random(field.shape) < 0.003] = 1.0
    field[10:22, 20:40, 30:60] = 0.9            # a solid block ...
    field[14:18, 26:34, 38:52] = 0.1            # ... with an enclosed cavity (filled by binary_fill_holes)
    field[15, 29, 36:40] = 0.1                  # and one that tunnels... stays inside, still enclosed
    field[30:34, 50:60, 10:30] = 0.9
    field[31:33, 53:57, 0:20] = 0.1             # a cavity open to the frame border: must NOT be filled
    dim_res = {"X": 0.2, "Y": 0.2, "Z": 0.25, "T": 1.0}
    spec = P.FrameSpec(dim_res=dim_res, no_z=False)
    stages = {}
    ref = P.label_frame(field, spec, 0.5, stages=stages)
    assert (stages["filled"] != (field > 0.5)).any(), "case should exercise fill-holes"
    lab = Label(_info(field.shape, dim_res, False), device="b200")
    lab.num_t = 1
    got = lab._run_frame_full_volume(0, field, field, None, 0.5)
    assert np.array_equal(got, ref)
    # intensity-gated variant (labelling.py:550-552)
    raw = rng.random(field.shape).astype(np.float32)
    ref2 = P.label_frame(field, spec, 0.5, raw=raw, intensity_thresh=0.3)
    got2 = lab._run_frame_full_volume(0, raw, field, 0.3, 0.5)
    assert np.array_equal(got2, ref2)


def test_sharded_labeller_on_one_gpu_equals_the_reference_labels(tmp_path):
    """nellie_b200/sharded_label.py with the CUDA local-CCL callback (world size 1: no seam, but the whole
    fill-holes / size filter / majority / numbering chain on device tensors) against the executed reference."""
    import torch
    import torch.distributed as dist
    from nellie_b200.sharded_label import ZShardedLabeller, cuda_local_label
    g = load_golden("label3d")
    created = False
    if not dist.is_initialized():
        store = dist.FileStore(str(tmp_path / "store"), 1)
        dist.init_process_group("nccl", store=store, rank=0, world_size=1)
        created = True
    try:
        thr = np.float32(g["meta"]["frangi_thresh"])
        mask = torch.from_numpy(g["frangi"] > thr).cuda()
        nz, ny, nx = mask.shape
        lab = ZShardedLabeller(0, nz, nz, ny, nx, cuda_local_label)
        got = lab.label(mask, int(g["min_area"])).cpu().numpy()
        assert got.dtype == np.int32
        assert np.array_equal(got, g["labels"])
    finally:
        if created:
            dist.destroy_process_group()


def test_z_sharded_label_on_two_gpus():
    """Needs >= 2 GPUs (skipped on a single-GPU box): torchrun scripts/zshard_label_check.py — the Z-sharded labeller
    against the single-GPU kernel on a slab pair, and Label(z_shard=(rank, world)).run() on files against Label.run()."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29543",
                        os.path.join(root, "scripts", "zshard_label_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
