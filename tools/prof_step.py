"""Run a few fused Jacobi plume steps (for ncu captures of the stencil kernels)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
os.environ["FLUIDNET_B200_GRAPHS"] = "0"
import bench
from fluidnet_cxx_b200.lib import fluid, simulate
res = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
wl = dict(res=(1, res, res), method="jacobi", jacobi_iters=8)
mconf = bench.plume_mconf(8, "jacobi")
U_np, rho_np = bench.synthetic_state_numpy(1, res, res, 0)
bd = {k: torch.zeros(1, c, 1, res, res, device="cuda") for k, c in (("p", 1), ("U", 2), ("flags", 1), ("density", 1))}
bench.init_state(fluid, wl, mconf, bd, U_np, rho_np, lambda a: torch.from_numpy(a).cuda())
for _ in range(3):
    simulate(mconf, bd, None, "jacobi")
torch.cuda.synchronize()
print("ok")
