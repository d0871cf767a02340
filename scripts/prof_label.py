"""One Label pass on a 512^3 Frangi frame (profiling driver for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
from nellie_b200 import Filter, Label
from nellie_b200.phantoms import tubular_phantom
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
dim = {"X": 0.1, "Y": 0.1, "Z": 0.1, "T": 1.0}
info = SimpleNamespace(no_t=False, no_z=False, shape=(2, n, n, n), axes="TZYX", dim_res=dim)
flt = Filter(info, device="b200"); flt._get_t(); flt._set_default_sigmas()
lab = Label(info, device="b200")
raw = tubular_phantom((n,) * 3, seed=5000, device=dev)
fr = flt.filter_frame_device(raw).clone()
torch.cuda.synchronize()
for _ in range(int(os.environ.get("REPS", "2"))):
    labels, ft = lab.label_frame_device(fr, raw)
torch.cuda.synchronize()
print("labels", int(labels.max()), "thr", ft, "fg frac", float((labels > 0).float().mean()))
