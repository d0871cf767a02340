"""Run the Filter path once on a 512^3 phantom (profiling driver for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import DIM_RES_CFG3, SIGMAS_CFG3
from nellie_b200.engine import FilterParams, FrangiEngine3D
from nellie_b200.phantoms import tubular_phantom
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
params = FilterParams(dim_res=DIM_RES_CFG3, no_z=False, sigmas=SIGMAS_CFG3[:int(os.environ.get("NSIG", "6"))])
eng = FrangiEngine3D((n, n, n), params, device=dev)
print("div_mode", eng.div_mode)
frame = tubular_phantom((n, n, n), seed=3, device=dev)
for _ in range(int(os.environ.get("REPS", "1"))):
    out = eng.filter_frame(frame)
torch.cuda.synchronize()
rec = eng.sigma_records()
print("gamma", rec[:, 0], "skip", rec[:, 5])
print("alive frac", float((eng.acc >= 0).float().mean()), "nonzero out", float((out > 0).float().mean()))
