"""Per-sigma diagnostics of the 3-D Filter path on the bench phantom: kernel times, alive / response fractions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import DIM_RES_CFG3, SIGMAS_CFG3
from nellie_b200.engine import FilterParams, FrangiEngine3D
from nellie_b200.phantoms import tubular_phantom
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
frame = tubular_phantom((n, n, n), seed=3, device=dev)
for ns in range(1, len(SIGMAS_CFG3) + 1):
    params = FilterParams(dim_res=DIM_RES_CFG3, no_z=False, sigmas=SIGMAS_CFG3[:ns])
    eng = FrangiEngine3D((n, n, n), params, device=dev)
    eng.load_frame(frame); eng.run_sigmas(); torch.cuda.synchronize()
    eng.profile = []
    eng.load_frame(frame); eng.run_sigmas()
    torch.cuda.synchronize()
    per = {}
    for name, e0, e1 in eng.profile:
        per.setdefault(name, []).append(e0.elapsed_time(e1))
    alive = float((eng.acc >= 0).float().mean()); nz = float((eng.acc > 0).float().mean())
    a4 = (eng.acc >= 0).view(n, n, n // 4, 4).any(-1).float().mean().item()
    print(f"nsig={ns} alive={alive:.4f} alive_groups4={a4:.4f} response>0={nz:.4f} "
          f"K3={per['nb200_frangi_accumulate'][-1]:.2f}ms K2={per['nb200_hessian_stats'][-1]:.2f}ms "
          f"gauss={sum(per['nb200_gauss_axis'][-3:]):.2f}ms", flush=True)
    del eng
    torch.cuda.empty_cache()
