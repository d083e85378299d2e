"""Golden vectors of the Rayleigh-Taylor configuration (BASELINE.json configs[2] at fixture size):
pytorch/rayleighTaylor.py:140-167 state (emptyDomain + createRayleighTaylorBCs, periodic-y seam) advanced
by the REFERENCE's own lib.simulate on CPU (patched build oracle/_ref), with the shipped ScaleNet and with
the Jacobi solver.  Run in the build container:  python tests/golden/make_golden_rt.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import ref_loader  # noqa: E402

RT = {"dt": 0.5, "maccormackStrength": 0.6, "sampleOutsideFluid": False, "buoyancyScale": 1.0, "gravityScale": 0,
      "viscosity": 0, "correctScalar": False, "operatingDensity": 0.0, "gravityVec": {"x": 0.0, "y": 1.0, "z": 0.0},
      "pTol": 0.0, "jacobiIter": 34, "rho1": -0.01, "rho2": 0.01, "perturbThickness": 100, "perturbAmplitude": 0.01,
      "height": 0.5, "periodic-y": True, "periodic-x": False}


def rt_state(reflib, resY, resX, mconf):
    bd = {"p": torch.zeros(1, 1, 1, resY, resX), "U": torch.zeros(1, 2, 1, resY, resX),
          "flags": torch.zeros(1, 1, 1, resY, resX), "density": torch.zeros(1, 1, 1, resY, resX)}
    reflib.fluid.emptyDomain(bd["flags"])
    reflib.fluid.createRayleighTaylorBCs(bd, mconf, rho1=mconf["rho1"], rho2=mconf["rho2"])
    return bd


def main():
    torch.set_num_threads(4)
    reflib, net, mconf_net = ref_loader.load_scalenet()
    out = {}
    for method in ("convnet", "jacobi"):
        m = dict(mconf_net)
        m.update(RT)
        m["simMethod"] = method
        net.mconf = m
        net.scale.mconf = m
        bd = rt_state(reflib, 64, 48, m)
        # a seeded velocity perturbation so that the first steps are not trivially zero
        g = torch.Generator().manual_seed(5)
        bd["U"] = bd["U"] + 0.05 * torch.randn(bd["U"].shape, generator=g)
        out[f"{method}/U0"] = bd["U"].numpy().copy()
        out[f"{method}/density0"] = bd["density"].numpy().copy()
        out[f"{method}/flags"] = bd["flags"].numpy().copy()
        with torch.no_grad():
            for it in range(1, 4):
                reflib.simulate(m, bd, net, method)
                for k in ("p", "U", "density"):
                    out[f"{method}/step{it}_{k}"] = bd[k].numpy().copy()
    np.savez_compressed(os.path.join(HERE, "rt64_periodic.npz"), **out)
    print("rt64_periodic.npz", os.path.getsize(os.path.join(HERE, "rt64_periodic.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
