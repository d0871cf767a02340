"""Experiment: blur of sigma i+1 on a side stream under K2/K3 of sigma i (engine.overlap_blur), with the march kernels
at 2 or 1 CTA/SM (NB200_FAST_CTAS).  python scripts/overlap_test.py [size]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nellie_b200.engine import FilterParams, FrangiEngine3D
from nellie_b200.phantoms import tubular_phantom
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
eng = FrangiEngine3D((n, n, n), FilterParams(dim_res={"X": 0.1, "Y": 0.1, "Z": 0.1, "T": 1.0}, sigmas=[1.0, 1.4, 1.8, 2.2, 2.6, 3.0]), device="cuda")
frame = tubular_phantom((n, n, n), seed=3, device="cuda")
for overlap in (False, True):
    eng.overlap_blur = overlap
    for _ in range(2):
        out = eng.filter_frame(frame)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = eng.filter_frame(frame)
    e1.record()
    torch.cuda.synchronize()
    print(f"overlap_blur={overlap} ctas={os.environ.get('NB200_FAST_CTAS', '2')}: {e0.elapsed_time(e1) / 3:.2f} ms/frame, nonzero {int((out > 0).sum())}")
