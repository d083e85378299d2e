"""A few Jacobi steps of a plume workload for ncu (profiles the stencil kernels).
    python tools/prof_step.py [res] [steps] [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
import torch
import bench
res = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 16
from fluidnet_cxx_b200.lib import fluid
sim = importlib.import_module("fluidnet_cxx_b200.lib.simulate")
wl = dict(bench.WORKLOADS["plume4096_jacobi100"], res=(1, res, res), jacobi_iters=iters)
mconf = bench.workload_mconf(wl)
U_np, rho_np = bench.synthetic_state_numpy(1, res, res, 0)
bd = {"p": torch.zeros(1, 1, 1, res, res, device="cuda"), "U": torch.zeros(1, 2, 1, res, res, device="cuda"),
      "flags": torch.zeros(1, 1, 1, res, res, device="cuda"), "density": torch.zeros(1, 1, 1, res, res, device="cuda")}
bench.init_state(fluid, wl, mconf, bd, U_np, rho_np, lambda a: torch.from_numpy(a).cuda())
os.environ["FLUIDNET_B200_GRAPHS"] = "0"
for _ in range(steps):
    sim.simulate(mconf, bd, None, "jacobi")
torch.cuda.synchronize()
print("done")
