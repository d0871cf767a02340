"""torchrun script: Z-sharded Filter output must equal the single-GPU output bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from nellie_b200.engine import FilterParams, FrangiEngine3D
from nellie_b200.phantoms import tubular_phantom
from nellie_b200.sharding import ZShardedFilter

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
ok = True
for shape, dim_res, sigmas in [((64, 96, 160), {"X": 0.1, "Y": 0.1, "Z": 0.1, "T": 1}, [1.0, 1.4, 1.8, 2.2, 2.6, 3.0]),
                               ((48, 130, 150), {"X": 0.1, "Y": 0.1, "Z": 0.25, "T": 1}, None)]:
    params = FilterParams(dim_res=dim_res, no_z=False, sigmas=sigmas)
    full = tubular_phantom(shape, seed=9, device=dev, n_tubes=12)
    zs = ZShardedFilter(shape, params, dev)
    slab = zs.slab_of(full)
    out = zs.filter_frame(slab).clone()                      # eager launches
    out2 = zs.filter_frame(slab).clone()                     # captured as a CUDA graph (kernels + NCCL), then replayed
    out3 = zs.filter_frame(slab).clone()                     # replay
    assert zs.engine.use_graph and len(zs.engine._graphs) == 1, "graph capture did not happen"
    graph_ok = torch.equal(out, out2) and torch.equal(out, out3)
    sp = zs.engine.sigma_records()
    zs.engine._graphs.clear()
    ref_eng = FrangiEngine3D(shape, params, device=dev)
    ref = ref_eng.filter_frame(full)
    sp_ref = ref_eng.sigma_records()
    same = torch.equal(out, ref[zs.z0:zs.z1]) and graph_ok
    same_sp = bool((sp[:, :6] == sp_ref[:, :6]).all())
    print(f"rank {rank}/{world} shape {shape}: slab [{zs.z0},{zs.z1}) identical={same} scalars identical={same_sp} "
          f"nonzero={int((out > 0).sum())}", flush=True)
    ok = ok and same and same_sp
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
