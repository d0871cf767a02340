"""CUDA-event time of nb200_ccl_label alone (pack + tile + border + roots + numbering + labels) on a noisy 512^3 mask:
the voxels above a cut (26-connectivity) and their complement (6-connectivity)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nellie_b200 import _cabi
from nellie_b200.phantoms import tubular_phantom
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
lib = _cabi.load()
raw = tubular_phantom((n,) * 3, seed=5000, device=dev)
fg = (raw > 160.0).to(torch.uint8).contiguous()          # tubes above the noisy background
bg = (1 - fg).contiguous()
ws = torch.empty(int(lib.nb200_label_workspace_bytes(n, n, n)), dtype=torch.uint8, device=dev)
labels = torch.empty((n,) * 3, dtype=torch.int32, device=dev)
cnt = torch.zeros(1, dtype=torch.int64, device=dev)
P = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
out = []
for name, m, full in (("fg26", fg, 1), ("bg6", bg, 0)):
    for _ in range(2):
        _cabi.check(lib.nb200_ccl_label(P(m), n, n, n, full, P(labels), P(ws), P(cnt), st), "ccl")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lib.nb200_ccl_label(P(m), n, n, n, full, P(labels), P(ws), P(cnt), st)
    e1.record()
    torch.cuda.synchronize()
    out.append(f"{name}: {e0.elapsed_time(e1) / 5:.3f} ms ({int(cnt.item())} comps, frac {float(m.float().mean()):.3f})")
print(" | ".join(out), flush=True)
