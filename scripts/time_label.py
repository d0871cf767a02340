"""CUDA-event time of one Label frame (512^3 by default): thresholds + nb200_label_frame through Label.label_frame_device."""
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
for _ in range(3):
    labels, ft = lab.label_frame_device(fr, raw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    labels, ft = lab.label_frame_device(fr, raw)
e1.record()
torch.cuda.synchronize()
h = int(torch.hash_tensor(labels).item()) if hasattr(torch, "hash_tensor") else int(labels.to(torch.int64).sum().item())
print("label ms/frame", e0.elapsed_time(e1) / 10,
      "labels", int(labels.max()), "sum", int(labels.to(torch.int64).sum().item()), "hash", h, flush=True)
