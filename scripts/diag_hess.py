import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
import numpy as np, torch
from conftest import load_golden, spec_from_meta
from oracle import pipeline as P
from nellie_b200 import _cabi
from nellie_b200.engine import FilterParams, FrangiEngine3D
name = sys.argv[1] if len(sys.argv) > 1 else "phantom3d_aniso"
g = load_golden(name); spec = spec_from_meta(g["meta"]); tr = []
P.frangi_frame(g["raw"], spec, trace=tr)
params = FilterParams(dim_res=g["meta"]["dim_res"], no_z=False, sigmas=g["meta"].get("explicit_sigmas"))
eng = FrangiEngine3D(g["raw"].shape, params, device="cuda")
print("div_mode", eng.div_mode, "fd", eng.fd)
t = tr[0]
gauss = torch.from_numpy(t["gauss"]).cuda()
comp, frob_sq, max_abs, frob = P.hessian(t["gauss"], spec)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
own = eng.vol()
eng.samples.zero_()
_cabi.call("nb200_hstats_reset", C.c_void_p(eng.hstats.data_ptr()), st)
_cabi.call("nb200_hessian_stats", C.c_void_p(gauss.data_ptr()), C.byref(own), eng._fd_c, eng.div_mode, None, 1, 1, 1,
           C.c_void_p(eng.samples.data_ptr()), C.c_void_p(eng.hstats.data_ptr()), st)
torch.cuda.synchronize()
got = eng.samples.cpu().numpy().reshape(t["gauss"].shape)
ref = np.sqrt(frob_sq)
bad = np.argwhere(got != ref)
print("mismatching voxels", len(bad), "of", got.size)
print("first", bad[:12].tolist())
if len(bad):
    for ax, nm in enumerate("zyx"):
        vals, cnt = np.unique(bad[:, ax], return_counts=True)
        print(nm, dict(zip(vals.tolist()[:20], cnt.tolist()[:20])))
hs = eng.hstats.cpu().numpy()
print("max_abs", np.array([hs[0]], dtype=np.uint32).view(np.float32)[0], max_abs)
out6 = torch.zeros((6,) + t["gauss"].shape, dtype=torch.float32, device="cuda")
for mode in (0, 1):
    _cabi.call("nb200_hessian_components", C.c_void_p(gauss.data_ptr()), C.byref(own), eng._fd_c, mode,
               C.c_void_p(out6.data_ptr()), st)
    torch.cuda.synchronize()
    o = out6.cpu().numpy()
    for j, nm in enumerate(["hxx", "hxy", "hxz", "hyy", "hyz", "hzz"]):
        d = o[j] != comp[nm]
        msg = ""
        if d.any():
            b = np.argwhere(d)[0]
            msg = f" first {b.tolist()} got {o[j][tuple(b)]!r} ref {comp[nm][tuple(b)]!r}"
        print("mode", mode, nm, "mismatches", int(d.sum()), msg)
b = np.argwhere(got != ref)
print("n bad", len(b))
f32 = np.float32
c = {k: v.astype(np.float64) for k, v in comp.items()}
def r(x): return x.astype(np.float32).astype(np.float64)
sq = {k: r(v * v) for k, v in c.items()}
cands = {
  "ref_order": r(r(r(sq["hxx"] + sq["hyy"]) + sq["hzz"]) + r(2.0 * r(r(sq["hxy"] + sq["hxz"]) + sq["hyz"]))),
  "fma_diag_last": r(r(r(sq["hxx"] + sq["hyy"]) + sq["hzz"]) + 2.0 * r(r(sq["hxy"] + sq["hxz"]) + sq["hyz"])),
  "fma_yy": r(r(r(c["hyy"] * c["hyy"] + sq["hxx"]) + sq["hzz"]) + r(2.0 * r(r(sq["hxy"] + sq["hxz"]) + sq["hyz"]))),
  "fma_xx": r(r(c["hzz"] * c["hzz"] + r(sq["hxx"] + sq["hyy"])) + r(2.0 * r(r(sq["hxy"] + sq["hxz"]) + sq["hyz"]))),
  "fma_hxz": r(r(r(sq["hxx"] + sq["hyy"]) + sq["hzz"]) + r(2.0 * r(r(c["hxz"] * c["hxz"] + sq["hxy"]) + sq["hyz"]))),
  "fma_hyz": r(r(r(sq["hxx"] + sq["hyy"]) + sq["hzz"]) + r(2.0 * r(c["hyz"] * c["hyz"] + r(sq["hxy"] + sq["hxz"])))),
}
gsq = got.astype(np.float64)
for nm, v in cands.items():
    print(nm, "matches sqrt:", int((np.sqrt(v.astype(np.float32)) == got).sum()), "of", got.size)
for bb in b[:4]:
    print(tuple(bb), repr(got[tuple(bb)]), repr(ref[tuple(bb)]))
