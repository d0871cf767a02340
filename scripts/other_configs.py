"""Timing of the BASELINE configs that are parity cases rather than bench lines (#2, #4, #5), device-resident."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import numpy as np, torch
from nellie_b200 import Filter, Label
from nellie_b200.engine import FilterParams, FrangiEngine3D
from nellie_b200.phantoms import tubular_phantom

dev = torch.device("cuda", 0)


def timed(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# config #2: 512^3, min/max radius giving 4 sigmas [1.0, 1.2, 1.4, 1.6]
dim2 = {"X": 0.125, "Y": 0.125, "Z": 0.125, "T": 1.0}
p2 = FilterParams(dim_res=dim2, no_z=False, min_radius_um=0.25, max_radius_um=0.675)
eng2 = FrangiEngine3D((512,) * 3, p2, device=dev)
f2 = tubular_phantom((512,) * 3, seed=2, device=dev)
ms = timed(lambda: eng2.filter_frame(f2))
print(f"cfg2 512^3 {len(eng2.sigmas)} sigmas {eng2.sigmas}: {ms:.2f} ms/frame = {512**3 / ms / 1e6:.2f} Gvoxel/s")

# config #5: 512^3 frame, 5 sigmas, Filter then Label (device resident)
dim5 = {"X": 0.1, "Y": 0.1, "Z": 0.1, "T": 1.0}
info5 = SimpleNamespace(no_t=False, no_z=False, shape=(16, 512, 512, 512), axes="TZYX", dim_res=dim5)
flt = Filter(info5, device="b200")
flt._get_t(); flt._set_default_sigmas()
lab = Label(info5, device="b200")
raw5 = tubular_phantom((512,) * 3, seed=5000, device=dev)
msf = timed(lambda: flt.filter_frame_device(raw5))
fr = flt.filter_frame_device(raw5).clone()
msl = timed(lambda: lab.label_frame_device(fr, raw5))
labels, ft = lab.label_frame_device(fr, raw5)
print(f"cfg5 512^3 {len(flt.sigmas)} sigmas: filter {msf:.2f} ms + label {msl:.2f} ms per frame "
      f"({int(labels.max())} labels) = {512**3 / (msf + msl) / 1e6:.2f} Gvoxel/s")

# config #4: 2048^2 2-D frames incl. LoG
dim4 = {"X": 0.1, "Y": 0.1, "T": 1.0}
info4 = SimpleNamespace(no_t=False, no_z=True, shape=(256, 2048, 2048), axes="TYX", dim_res=dim4)
f4 = Filter(info4, device="b200")
f4._get_t(); f4._set_default_sigmas()
from nellie_b200.phantoms import tubular_phantom_np
img = torch.from_numpy(tubular_phantom_np((1, 2048, 2048), seed=4000, n_tubes=200)[0]).to(dev)
ms4 = timed(lambda: f4.filter_frame_device(img), reps=10)
t0 = time.perf_counter()
for _ in range(10):
    f4.filter_frame_device(img)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 10 * 1e3
print(f"cfg4 2048^2 {len(f4.sigmas)} sigmas: {ms4:.3f} ms/frame device, {wall:.3f} ms/frame wall, "
      f"{f4._engine.launches // 13} calls/frame = {2048**2 / ms4 / 1e6:.2f} Gpixel/s")
