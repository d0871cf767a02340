"""Per-sigma K2/K3 timing for one setting of NB200_STREAM_CTAS (profiling helper)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import DIM_RES_CFG3, SIGMAS_CFG3
from nellie_b200.engine import FilterParams, FrangiEngine3D
from nellie_b200.phantoms import tubular_phantom
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
eng = FrangiEngine3D((n, n, n), FilterParams(dim_res=DIM_RES_CFG3, no_z=False, sigmas=SIGMAS_CFG3), device=dev)
frame = tubular_phantom((n, n, n), seed=3, device=dev)
for _ in range(2):
    eng.filter_frame(frame)
eng.profile = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    eng.filter_frame(frame)
e1.record()
torch.cuda.synchronize()
print("NB200_STREAM_CTAS", os.environ.get("NB200_STREAM_CTAS"), "ms/step", e0.elapsed_time(e1) / 3)
for name in ("nb200_hessian_stats_code", "nb200_frangi_sparse"):
    evs = [(a, b) for nm, a, b in eng.profile if nm == name]
    print(name, [round(float(np.mean([a.elapsed_time(b) for a, b in evs[i::6]])), 3) for i in range(6)])
