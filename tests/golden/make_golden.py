#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own ATen CPU path.

Runs only in the build container (needs oracle/_ref, built from /root/reference
by oracle/build_ref.py).  The committed .npz files are what travels: they pin
the C oracle (tests/test_oracle_golden.py, CPU) and the CUDA path (tests -m gpu).

    python oracle/build_ref.py && python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_loader  # noqa: E402

FLUID, OBST, EMPTY = 1.0, 2.0, 4.0


def make_flags(rng, H, W, border, nboxes, with_empty=False):
    f = np.full((1, 1, 1, H, W), FLUID, np.float32)
    if border == "obstacle":
        f[..., 0, :] = OBST; f[..., -1, :] = OBST; f[..., :, 0] = OBST; f[..., :, -1] = OBST
    for _ in range(nboxes):
        h, w = rng.randint(1, 5), rng.randint(1, 5)
        y, x = rng.randint(1, H - 1 - h), rng.randint(1, W - 1 - w)
        f[0, 0, 0, y:y + h, x:x + w] = OBST
    if with_empty:
        for _ in range(3):
            h, w = rng.randint(1, 4), rng.randint(1, 4)
            y, x = rng.randint(1, H - 1 - h), rng.randint(1, W - 1 - w)
            f[0, 0, 0, y:y + h, x:x + w] = EMPTY
    return f


def random_case(seed, H, W, border, nboxes, vscale, with_empty=False):
    rng = np.random.RandomState(seed)
    flags = make_flags(rng, H, W, border, nboxes, with_empty)
    U = (rng.randn(1, 2, 1, H, W) * vscale).astype(np.float32)
    rho = rng.rand(1, 1, 1, H, W).astype(np.float32)
    p = rng.randn(1, 1, 1, H, W).astype(np.float32)
    return flags, U, rho, p


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).clone()


def gen_ops(fluid):
    out = {}
    cases = [
        # name, seed, H, W, border, nboxes, vscale, dt, with_empty
        ("a", 0, 20, 24, "obstacle", 3, 0.5, 0.1, False),
        ("b", 1, 24, 20, "obstacle", 4, 8.0, 0.4, False),   # long back-traces, many blocked rays
        ("c", 2, 16, 28, "fluid", 2, 6.0, 0.5, False),     # fluid border: rays exit the domain (case 1)
        ("d", 3, 22, 22, "obstacle", 0, 2.0, 0.25, True),   # Empty cells
        ("e", 4, 12, 40, "fluid", 5, 20.0, 0.3, True),      # traces far outside the grid
        ("f", 5, 33, 17, "obstacle", 6, 3.0, 1.0, False),
    ]
    for name, seed, H, W, border, nboxes, vscale, dt, we in cases:
        flags, U, rho, p = random_case(seed, H, W, border, nboxes, vscale, we)
        g = {"flags": flags, "U": U, "rho": rho, "p": p, "dt": np.float32(dt)}
        tf, tU, tr, tp = t(flags), t(U), t(rho), t(p)
        for so in (False, True):
            for m in ("maccormackFluidNet", "eulerFluidNet"):
                try:
                    r = fluid.advectScalar(dt, tr.clone(), tU.clone(), tf.clone(), m, 1, so, 0.6).numpy()
                except RuntimeError as e:      # the reference asserts (e.g. "case 1 exited bounds")
                    print(f"  [{name}] advectScalar {m} so={so}: reference raised: {str(e)[:60]}")
                    continue
                g[f"advectScalar_{m}_{int(so)}"] = r
        for m in ("maccormackFluidNet", "eulerFluidNet"):
            g[f"advectVelocity_{m}"] = fluid.advectVelocity(dt, tU.clone(), tU.clone(), tf.clone(), m, 1, 0.6).numpy()
        # orig != U (viscous branch of simulate.py:92-94)
        orig = (U * 0.7 + 0.1).astype(np.float32)
        g["orig2"] = orig
        g["advectVelocity_orig2"] = fluid.advectVelocity(dt, t(orig), tU.clone(), tf.clone(),
                                                         "maccormackFluidNet", 1, 0.75).numpy()
        grav = torch.tensor([0.3, -0.25, 0.0])
        g["gravity"] = grav.numpy()
        g["addBuoyancy"] = fluid.addBuoyancy(tU.clone(), tf, tr, grav, 0.05, dt).numpy()
        g["addGravity"] = fluid.addGravity(tU.clone(), tf, grav, dt).numpy()
        g["setWallBcs"] = fluid.setWallBcs(tU.clone(), tf).numpy()
        div = fluid.velocityDivergence(tU.clone(), tf)
        g["velocityDivergence"] = div.numpy()
        Uu = tU.clone()
        fluid.velocityUpdate(pressure=tp, U=Uu, flags=tf)
        g["velocityUpdate"] = Uu.numpy()
        pj, res = fluid.solveLinearSystemJacobi(tf, div, False, 0.0, 17)
        g["jacobi17_p"], g["jacobi17_res"] = pj.numpy(), np.float32(res.item())
        pj, res = fluid.solveLinearSystemJacobi(tf, div, False, 0.0, 4)
        g["jacobi4_p"], g["jacobi4_res"] = pj.numpy(), np.float32(res.item())
        ptol = float(g["jacobi17_res"]) * 3.0    # terminates early on the residual test
        pj, res = fluid.solveLinearSystemJacobi(tf, div, False, ptol, 500)
        g["jacobi_tol"], g["jacobi_tol_p"], g["jacobi_tol_res"] = np.float32(ptol), pj.numpy(), np.float32(res.item())
        g["flagsToOccupancy"] = fluid.flagsToOccupancy(tf).numpy()
        for k, v in g.items():
            out[f"{name}/{k}"] = v
    return out


def plume_state(reflib, res, mconf_over=None):
    """plume.py:127-160 initial state + mconf from plumeConfig.yaml."""
    import yaml
    with open(os.path.join(ref_loader.REF_DIR, "plumeConfig.yaml")) as f:
        sim = yaml.load(f, Loader=yaml.SafeLoader)
    mconf = dict(sim)
    mconf.update({"sampleOutsideFluid": False})
    mconf.update(mconf_over or {})
    p = torch.zeros(1, 1, 1, res, res)
    U = torch.zeros(1, 2, 1, res, res)
    flags = torch.zeros(1, 1, 1, res, res)
    density = torch.zeros(1, 1, 1, res, res)
    reflib.fluid.emptyDomain(flags)
    bd = {"p": p, "U": U, "flags": flags, "density": density}
    reflib.fluid.createPlumeBCs(bd, mconf["injectionDensity"], mconf["injectionVelocity"], mconf["sourceRadius"])
    return mconf, bd


def gen_plume_jacobi(reflib):
    """BASELINE.json configs[0]: 128x128 plume, Jacobi 28 it., single-step + short multi-step."""
    mconf, bd = plume_state(reflib, 128, {"simMethod": "jacobi", "jacobiIter": 28, "pTol": 0.0})
    out = {}
    for k in ("UBC", "UBCInvMask", "densityBC", "densityBCInvMask", "flags"):
        out[f"init/{k}"] = bd[k].numpy().copy()
    snaps = (1, 2, 8, 24)
    with torch.no_grad():
        for it in range(1, max(snaps) + 1):
            reflib.simulate(mconf, bd, None, "jacobi")
            if it in snaps:
                for k in ("p", "U", "density"):
                    out[f"step{it}/{k}"] = bd[k].numpy().copy()
    return out


def gen_cnn(reflib, net):
    out = {}
    for name, seed, H, W in (("a", 0, 32, 32), ("b", 1, 48, 40), ("c", 2, 36, 52)):
        flags, U, rho, p = random_case(seed, H, W, "obstacle", 3, 0.5)
        data = torch.cat((t(p), t(U), t(flags), t(rho)), 1)
        with torch.no_grad():
            pp, UU = net(data)
            # also the bare MultiScaleNet on a random 2-channel input
            x = torch.from_numpy(np.random.RandomState(seed + 100).randn(1, 2, H, W).astype(np.float32))
            y = net.multiScale(x)
        out[f"{name}/flags"], out[f"{name}/U"], out[f"{name}/rho"], out[f"{name}/p"] = flags, U, rho, p
        out[f"{name}/p_out"], out[f"{name}/U_out"] = pp.numpy(), UU.numpy()
        out[f"{name}/msn_x"], out[f"{name}/msn_y"] = x.numpy(), y.numpy()
    return out


def gen_plume_cnn(reflib, net, mconf_net):
    """BASELINE.json configs[1] shape at a fixture-sized grid (64x64): plume + ScaleNet, 3 steps."""
    mconf, bd = plume_state(reflib, 64, {"simMethod": "convnet"})
    m = dict(mconf_net); m.update(mconf)
    net.mconf = m; net.scale.mconf = m
    out = {}
    with torch.no_grad():
        for it in range(1, 4):
            reflib.simulate(m, bd, net, "convnet")
            for k in ("p", "U", "density"):
                out[f"step{it}/{k}"] = bd[k].numpy().copy()
    return out


def main():
    torch.set_num_threads(4)
    reflib, net, mconf_net = ref_loader.load_scalenet()
    np.savez_compressed(os.path.join(HERE, "ops_2d.npz"), **gen_ops(reflib.fluid))
    np.savez_compressed(os.path.join(HERE, "plume128_jacobi28.npz"), **gen_plume_jacobi(reflib))
    np.savez_compressed(os.path.join(HERE, "scalenet_forward.npz"), **gen_cnn(reflib, net))
    np.savez_compressed(os.path.join(HERE, "plume64_convnet.npz"), **gen_plume_cnn(reflib, net, mconf_net))
    # the shipped weights as a flat fp32 archive (state-dict names under `multiScale.`)
    wdir = os.path.join(ROOT, "fluidnet_cxx_b200", "data")
    os.makedirs(wdir, exist_ok=True)
    sd = {k[len("multiScale."):]: v.numpy() for k, v in net.state_dict().items() if k.startswith("multiScale.")}
    np.savez(os.path.join(wdir, "scalenet_weights.npz"), **sd)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
