"""One 2048^2 frame through the 2-D Filter path with eager launches (profiling driver for ncu: 2-D kernels + thresholds)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nellie_b200.engine import FilterParams
from nellie_b200.engine2d import FrangiEngine2D
from nellie_b200.phantoms import tubular_phantom
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda", 0)
eng = FrangiEngine2D((n, n), FilterParams(dim_res={"X": 0.1, "Y": 0.1, "T": 1.0}, no_z=True), device=dev)
eng.use_graph = False
frame = tubular_phantom((n, n), seed=4000, device=dev, n_tubes=200)
for _ in range(2):
    out = eng.filter_frame(frame)
torch.cuda.synchronize()
print("nonzero", int((out > 0).sum()), "max", float(out.max()))
