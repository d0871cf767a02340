"""2+ GPU check of the Z-sharded Label (run by tests/test_label_gpu.py::test_z_sharded_label when >= 2 GPUs are visible):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        scripts/zshard_label_check.py

Every rank labels its slab with nellie_b200.sharded_label.ZShardedLabeller (CUDA local CCL + seam merge); rank 0
labels the whole frame with the single-GPU kernel (nb200_label_frame through LabelEngine) and compares."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from nellie_b200.labelling import LabelEngine
from nellie_b200.phantoms import tubular_phantom
from nellie_b200.sharded_label import ZShardedLabeller, cuda_local_label
from nellie_b200.sharding import z_partition

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
shape = (96, 160, 192)
field = tubular_phantom(shape, seed=11, device=dev)              # same seed on every rank: same frame
thr, min_area = 250.0, 62
z0, z1 = z_partition(shape[0], world)[rank]
lab = ZShardedLabeller(z0, z1 - z0, shape[0], shape[1], shape[2], cuda_local_label)
mine = lab.label(field[z0:z1] > thr, min_area)
ok = torch.ones(1, device=dev)
eng = LabelEngine(shape, False, min_area, 1_000_000, dev)
full = eng.label(field, thr)
ok[0] = float(torch.equal(mine, full[z0:z1]))
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print("labels", int(full.max()), "sharded == single GPU:", bool(ok.item()), flush=True)

# ---- the same through the stage class on files: Label(..., z_shard=(rank, world)).run() ----
import tempfile

import numpy as np

from nellie_b200 import Filter, Label, imio

root = [tempfile.mkdtemp() if rank == 0 else None]
dist.broadcast_object_list(root, src=0)
dim_res = {"X": 0.1, "Y": 0.1, "Z": 0.2, "T": 1.0}
frames = np.stack([tubular_phantom((40, 96, 128), seed=70 + t, device=dev).cpu().numpy() for t in range(2)])
info = [None]
if rank == 0:
    info[0] = imio.StackInfo.from_array(frames, "TZYX", dim_res, root[0], "p")
    Filter(info[0], device="b200").run()
    Label(info[0], device="b200").run()
    want = imio.read_tiff(info[0].pipeline_paths["im_instance_label"]).copy()
dist.broadcast_object_list(info, src=0)
dist.barrier()
Label(info[0], device="b200", z_shard=(rank, world)).run()
dist.barrier()
ok2 = torch.ones(1, device=dev)
if rank == 0:
    got = imio.read_tiff(info[0].pipeline_paths["im_instance_label"])
    same = bool(np.array_equal(got, want))
    print("Label.run(z_shard) == Label.run():", same, "labels per frame", [int(w.max()) for w in want], flush=True)
    ok2[0] = float(same)
dist.all_reduce(ok2, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if bool(ok.item()) and bool(ok2.item()) else 1)
