"""T-sharded Markers and Network stages in TWO processes on shared output files (gloo, CPU; kernels host-emulated).

Frames are independent: rank r runs the frames t with t % world == r and there is no data-path collective; the
one thing the ranks have to agree on is the output file, which rank 0 alone creates (sharding.allocate_shared_output) while
the others wait at a barrier and then map it.  Both ranks call run() — the protocol a torchrun launch uses on GPUs."""
import ctypes as C
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import markers_checks as MK
import network_checks as NK
from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _emu(unit, prefixes):
    from nellie_b200 import _cabi
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", f"{unit}.so"))
    for name, (argtypes, restype) in _cabi._SIGS.items():
        if name.startswith(prefixes):
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = argtypes, restype
    return lib


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nellie_b200 import mocap_marking as M, networking as N
        from nellie_b200.imio import StackInfo
        from oracle import pipeline as P
        mlib = _emu("markers_host", ("nb200_markers_", "nb200_gauss_"))
        nlib = _emu("network_host", ("nb200_network_",))
        g = NK.load_frame_case("network_frame_phantom3d_aniso")
        shape = g["labels"].shape
        T = 3
        labs = np.stack([np.roll(g["labels"], 3 * t, axis=2) for t in range(T)]).astype(np.int32)
        frs = np.stack([np.roll(g["frangi"], 3 * t, axis=2) for t in range(T)])
        skels = [np.roll(g["skeleton"], 3 * t, axis=2) for t in range(T)]
        raws = (frs * 1000).astype(np.float32)
        dim_res = g["meta"]["dim_res"]
        # every rank describes the same stack; the raw file and the inputs are written by rank 0 before the stages start
        if rank == 0:
            info = StackInfo.from_array(raws, "TZYX", dim_res, out_dir)
            info.allocate_memory(info.pipeline_paths["im_instance_label"], dtype="int32", data=labs)
            info.allocate_memory(info.pipeline_paths["im_preprocessed"], dtype="float32", data=frs)
        dist.barrier()
        if rank != 0:                                            # describe the files rank 0 wrote
            name = StackInfo.output_name("stack", "TZYX", dim_res, 0, 0, T - 1)
            info = StackInfo(os.path.join(out_dir, "nellie_necessities", name + ".ome.tif"), "TZYX", tuple(raws.shape),
                             dict(dim_res), out_dir, name, np.dtype(np.float32))

        class EmuMarkers(M.Markers):
            def _torch_device(self):
                return torch.device("cpu")

            def _engine_for(self, frame_shape):
                if not self.sigmas:
                    self._set_default_sigmas()
                if self._engine is None:
                    self._engine = M.MarkerEngine(tuple(frame_shape), False, tuple(float(s) for s in self.sigmas), self.z_ratio,
                                                  self.max_radius_px, self.peak_min_distance, "cpu", lib=mlib)
                return self._engine

        class EmuNetwork(N.Network):
            @property
            def device(self):
                return torch.device("cpu")

            def _engine(self):
                if self._net is None:
                    self._net = N.NetworkEngine(False, self.scaling, "cpu", lib=nlib)
                return self._net

            def _skeletonize(self, label_frame):
                t = [k for k in range(T) if np.array_equal(label_frame, labs[k])][0]
                return np.asarray(label_frame) * skels[t]

            def _get_pixel_class(self, skel):
                return torch.from_numpy(P.network_pixel_class(skel.numpy(), False).astype(np.uint8))

            def _get_branch_skel_labels(self, pixel_class):
                return torch.from_numpy(P.network_branch_labels(pixel_class.numpy(), False).astype(np.int32))

            def _remove_connected_label_pixels(self, skel):
                return torch.from_numpy(P.network_remove_connected(skel.numpy(), False).astype(np.int32))

        EmuMarkers(info, t_shard=(rank, world)).run()
        EmuNetwork(info, t_shard=(rank, world)).run()
        dist.barrier()
        if rank == 0:
            spec = P.MarkerSpec(dim_res=dim_res)
            for t in range(T):
                m, d, b = P.marker_frame(raws[t], labs[t], spec)
                assert np.array_equal(info.get_memmap(info.pipeline_paths["im_marker"])[t], m), t
                assert np.array_equal(info.get_memmap(info.pipeline_paths["im_distance"])[t], d), t
                assert np.array_equal(info.get_memmap(info.pipeline_paths["im_border"])[t], b), t
                br, pc, rl = P.network_frame(labs[t], frs[t], skels[t], tuple(g["scaling"]), False)
                assert np.array_equal(info.get_memmap(info.pipeline_paths["im_skel"])[t], br), t
                assert np.array_equal(info.get_memmap(info.pipeline_paths["im_pixel_class"])[t], pc), t
                assert np.array_equal(info.get_memmap(info.pipeline_paths["im_skel_relabelled"])[t], rl), t
            open(os.path.join(out_dir, "ok"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_markers_and_network_t_sharded_in_two_processes(tmp_path):
    import subprocess
    for unit in ("markers_host", "network_host"):                 # built by __graft_entry__.build(); rebuild if missing
        so = os.path.join(ROOT, "oracle", "_build", f"{unit}.so")
        if not os.path.exists(so):
            os.makedirs(os.path.dirname(so), exist_ok=True)
            subprocess.run(["g++", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC",
                            f"-DNB200_HOST_EMU=\"{os.path.join(ROOT, 'oracle', 'cuda_emu.h')}\"", "-x", "c++",
                            os.path.join(ROOT, "oracle", f"{unit}.cpp"), "-o", so], check=True)
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok")
