"""Per-sigma timing + candidate counters of the fast Hessian path (diagnostics; not a bench).
    python scripts/prof_fast.py [size] [nsigma]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nellie_b200.engine import FilterParams, FrangiEngine3D  # noqa: E402
from nellie_b200.phantoms import tubular_phantom  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 6
sig = [1.0, 1.4, 1.8, 2.2, 2.6, 3.0][:ns]
eng = FrangiEngine3D((n, n, n), FilterParams(dim_res={"X": 0.1, "Y": 0.1, "Z": 0.1, "T": 1.0}, sigmas=sig), device="cuda")
frame = tubular_phantom((n, n, n), seed=3, device="cuda")
eng.filter_frame(frame)
torch.cuda.synchronize()
eng.diag = torch.zeros(8, dtype=torch.int64, device="cuda")
eng.profile = []
eng.load_frame(frame)
eng.acc.zero_()
src = eng.cur
last = torch.zeros(8, dtype=torch.int64)
for i in range(len(sig)):
    src = eng._blur_sigma(i, src, {})
    eng._analyse_sigma(i, eng.gauss[src])
    torch.cuda.synchronize()
    d = eng.diag.cpu()
    alive = int((eng.acc >= 0).sum())
    print(f"sigma {sig[i]}: candidates {int(d[0] - last[0])} ({100.0 * float(d[0] - last[0]) / n ** 3:.2f} %), "
          f"uncertain kills {int(d[1] - last[1])}, survivors {int(d[2] - last[2])} "
          f"({100.0 * float(d[2] - last[2]) / n ** 3:.2f} %), alive after {100.0 * alive / n ** 3:.1f} %")
    last = d.clone()
rec = eng.sigma_records()
print("unsafe", rec[:, 9], "skip", rec[:, 5], "fs_min", rec[:, 10], "lo", rec[:, 12], "hi", rec[:, 13], "zc", rec[:, 14], "delta", rec[:, 15],
      "max_abs", rec[:, 4])
for name, (cnt, ms) in sorted(eng.profile_summary().items(), key=lambda kv: -kv[1][1]):
    print(f"{name:32s} {cnt:3d} {ms / cnt:9.3f} ms/launch")
