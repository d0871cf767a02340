"""Checks of the Network stage's device steps shared by the CPU run (csrc/network.cu host-emulated, the three label.cu
kernels replaced by the oracle — they are GPU-tested in tests/test_network_gpu.py) and the GPU run (tests/test_zz_network_gpu.py)."""
import json
import os

import numpy as np
import scipy.ndimage as ndi
import torch

from conftest import GOLDEN_DIR

FRAME_CASES = ["network_frame_sample_crop", "network_frame_phantom3d_aniso", "network_frame_phantom2d", "network_frame_cfg3",
               "network_frame_cfg3_half", "network_frame_phantom2d_half"]


def load_frame_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    p = np.load(os.path.join(GOLDEN_DIR, f"{str(z['parent'])}.npz"))
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(str(z["meta"]))
    from oracle.pipeline import tie_free
    d["labels"], d["frangi"] = p["labels"], tie_free(p["frangi"])
    return d


def engine(lib, device, no_z, scaling):
    from nellie_b200.networking import NetworkEngine
    return NetworkEngine(no_z, scaling, device, lib=lib)


def t(a, device):
    return torch.from_numpy(np.array(a, copy=True, order="C")).to(device)        # never aliases the caller's array


def check_host_steps_on_fixture(lib, device, name):
    """_add_missing_skeleton_labels and _relabel_objects against what the executed reference produced."""
    g = load_frame_case(name)
    eng = engine(lib, device, g["meta"]["no_z"], tuple(g["scaling"]))
    labels = t(g["labels"].astype(np.int32), device)
    max_label = int(g["labels"].max())
    added = eng.add_missing(t(g["cleaned"], device), labels, t(g["frangi"].astype(np.float32), device), max_label)
    assert np.array_equal(added.cpu().numpy(), g["added"])
    pre = eng.skeleton_labels(added, labels).cpu().numpy()
    assert np.array_equal(pre, (g["added"] > 0) * g["labels"])
    out = eng.relabel(t(g["branch"], device), labels, max_label).cpu().numpy().view(np.uint32)
    assert np.array_equal(out, g["relabelled"])
    assert eng.crop_voxels > 0


def check_relabel_against_scipy(lib, device, trials=24, seed=1):
    """Random labelled blobs, random seeds with random branch ids (so equidistant seeds with DIFFERENT labels are common),
    isotropic / anisotropic / random voxel sizes, 2-D and 3-D, objects without seeds and missing ids."""
    from oracle import pipeline as P
    rng = np.random.default_rng(seed)
    for trial in range(trials):
        no_z = trial % 3 == 0
        shape = tuple(int(v) for v in rng.integers(6, 28, 2 if no_z else 3))
        grids = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
        labels = np.zeros(shape, np.int32)
        for k in range(1, 9):
            c = [rng.uniform(0, s) for s in shape]
            r = rng.uniform(1.5, 9)
            labels[sum((g - ci) ** 2 for g, ci in zip(grids, c)) <= r * r] = k
        if trial % 5 == 0:
            labels[labels == 3] = 0
        branch = ((rng.random(shape) < 0.06) * rng.integers(1, 6, shape)).astype(np.int32)
        branch[labels == 0] = 0
        if trial % 4 == 0:
            branch[labels == 2] = 0
        scaling = (0.25, 0.0655, 0.0655)[-len(shape):] if trial % 2 else tuple(float(v) for v in rng.uniform(0.1, 1.0, len(shape)))
        if trial % 7 == 0:
            scaling = (1.0,) * len(shape)
        eng = engine(lib, device, no_z, scaling)
        if trial % 2:
            eng.max_crop_voxels = 300                 # several groups of objects, a big object alone in its group
        out = eng.relabel(t(branch, device), t(labels, device), int(labels.max())).cpu().numpy().view(np.uint32)
        assert np.array_equal(out, P.network_relabel_objects(branch, labels, scaling)), (trial, shape, scaling)
    # nothing to do: no objects / no seeds at all
    eng = engine(lib, device, False, (1.0, 1.0, 1.0))
    z = np.zeros((4, 5, 6), np.int32)
    assert not eng.relabel(t(z, device), t(z, device), 0).cpu().numpy().any()
    assert not eng.relabel(t(z, device), t(z + 1, device), 1).cpu().numpy().any()


def check_add_missing_against_oracle(lib, device):
    from oracle import pipeline as P
    rng = np.random.default_rng(5)
    for shape in ((9, 20, 22), (30, 31)):
        labels = ndi.label(rng.random(shape) < 0.25)[0].astype(np.int32)
        frangi = rng.random(shape).astype(np.float32) * (labels > 0)
        skel = np.where(rng.random(shape) < 0.05, labels, 0).astype(np.int32)
        skel[labels % 3 == 0] = 0                                       # every third object loses its skeleton
        eng = engine(lib, device, len(shape) == 2, (1.0,) * len(shape))
        got = eng.add_missing(t(skel, device), t(labels, device), t(frangi, device), int(labels.max())).cpu().numpy()
        assert np.array_equal(got, P.network_add_missing(skel, labels, frangi))
        assert (got != skel).sum() > 3


def check_add_missing_tie_rule(lib, device):
    """Equal maxima inside one object: scipy picks by an unstable sort (arbitrary); the kernel takes the first voxel in
    raster order that holds the maximum."""
    labels = np.zeros((6, 7, 8), np.int32)
    labels[1:5, 2:6, 1:7] = 1
    labels[0, 0, :3] = 2
    frangi = np.zeros(labels.shape, np.float32)
    frangi[labels == 1] = 0.25
    frangi[2, 3, 4] = frangi[3, 2, 5] = frangi[4, 5, 6] = 0.75
    frangi[0, 0, 1] = 0.5
    skel = np.zeros_like(labels)
    eng = engine(lib, device, False, (1.0, 1.0, 1.0))
    got = eng.add_missing(t(skel, device), t(labels, device), t(frangi, device), 2).cpu().numpy()
    assert got[2, 3, 4] == 1 and got[0, 0, 1] == 2 and (got > 0).sum() == 2


def check_stage_class_on_fixture(make_network, name, tmp_path=None):
    """Network._run_frame (and run() on files when tmp_path is given) against the executed reference's frame outputs,
    the skeleton mask of the fixture standing in for skimage's thinning on both sides."""
    from types import SimpleNamespace
    g = load_frame_case(name)
    no_z = g["meta"]["no_z"]
    info = SimpleNamespace(no_t=True, no_z=no_z, shape=(1,) + g["labels"].shape, axes="TYX" if no_z else "TZYX",
                           dim_res=g["meta"]["dim_res"])
    net = make_network(info, num_t=1, skeletonize=lambda mask: g["skeleton"])
    net.label_memmap, net.im_frangi_memmap = g["labels"][None], g["frangi"][None]
    net.shape = net.label_memmap.shape
    branch, pixel_class, relabelled = net._run_frame(0)
    assert branch.dtype == np.int32 and np.array_equal(branch, g["branch"])
    assert pixel_class.dtype == np.uint8 and np.array_equal(pixel_class, g["pixel_class"])
    assert relabelled.dtype == np.uint32 and np.array_equal(relabelled, g["relabelled"])
    assert np.array_equal(net._add_missing_skeleton_labels(g["cleaned"], g["labels"], g["frangi"]), g["added"])
    assert np.array_equal(net._relabel_objects(g["branch"], g["labels"]), g["relabelled"])
    if tmp_path is None:
        return
    from nellie_b200.imio import StackInfo
    raws = np.stack([g["frangi"], g["frangi"]])                          # the raw image is not read by the stage
    sinfo = StackInfo.from_array(raws, "TYX" if no_z else "TZYX", g["meta"]["dim_res"], str(tmp_path))
    for key in ("im_instance_label", "im_preprocessed", "im_skel", "im_pixel_class", "im_skel_relabelled"):
        sinfo.create_output_path(key)
    sinfo.allocate_memory(sinfo.pipeline_paths["im_instance_label"], dtype="int32", data=np.stack([g["labels"], g["labels"]]).astype(np.int32))
    sinfo.allocate_memory(sinfo.pipeline_paths["im_preprocessed"], dtype="float32", data=raws)
    for shard in ((0, 2), (1, 2)):
        make_network(sinfo, skeletonize=lambda mask: g["skeleton"], t_shard=shard).run()
    for tt in range(2):
        assert np.array_equal(sinfo.get_memmap(sinfo.pipeline_paths["im_skel"])[tt], g["branch"])
        assert np.array_equal(sinfo.get_memmap(sinfo.pipeline_paths["im_pixel_class"])[tt], g["pixel_class"])
        assert np.array_equal(sinfo.get_memmap(sinfo.pipeline_paths["im_skel_relabelled"])[tt], g["relabelled"])
