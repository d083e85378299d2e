"""Golden vectors of the two `simulate.py` side branches: addViscosity
(viscosity.py:7-70) and correctScalar (advection.py:9-12), from the REFERENCE on CPU (oracle/_ref), plus
a viscous simulate() step.  Run in the build container:  python tests/golden/make_golden_extras.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), HERE]
import ref_loader  # noqa: E402
from make_golden import random_case, t  # noqa: E402


def main():
    torch.set_num_threads(4)
    reflib = ref_loader.load()
    fluid = reflib.fluid
    out = {}
    for name, seed, H, W in (("a", 3, 24, 31), ("b", 4, 40, 36)):
        flags, U, rho, p = random_case(seed, H, W, "obstacle", 3, 0.5, with_empty=(name == "b"))
        out[f"{name}/flags"], out[f"{name}/U"], out[f"{name}/rho"] = flags, U, rho
        for visc in (0.05, 1.3):
            Uv = t(U).clone()
            fluid.addViscosity(0.1, Uv, t(flags), visc)
            out[f"{name}/visc_{visc}"] = Uv.numpy()
        div = fluid.velocityDivergence(t(U), t(flags))
        r2 = t(rho).clone()
        fluid.correctScalar(0.1, r2, div, t(flags))
        out[f"{name}/div"] = div.numpy()
        out[f"{name}/corrected"] = r2.numpy()
    # one viscous, scalar-corrected Jacobi step of lib.simulate on a plume state (simulate.py:67-69,79-81)
    from make_golden import plume_state
    mconf, bd = plume_state(reflib, 48, {"simMethod": "jacobi", "jacobiIter": 12, "viscosity": 0.2, "correctScalar": True})
    g = torch.Generator().manual_seed(9)
    bd["U"] = bd["U"] + 0.3 * torch.randn(bd["U"].shape, generator=g)
    for k in ("U", "density", "flags", "UBC", "UBCInvMask", "densityBC", "densityBCInvMask"):
        out[f"sim/{k}0"] = bd[k].numpy().copy()
    for it in range(1, 3):
        reflib.simulate(mconf, bd, None, "jacobi")
        for k in ("p", "U", "density"):
            out[f"sim/step{it}_{k}"] = bd[k].numpy().copy()
    np.savez_compressed(os.path.join(HERE, "extras_2d.npz"), **out)
    print("extras_2d.npz", os.path.getsize(os.path.join(HERE, "extras_2d.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
