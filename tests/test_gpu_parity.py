"""GPU parity tests: the sm_100a path (called through the C-ABI via lib.fluid) against
 (1) the committed outputs of the reference's own ATen CPU path (tests/golden/*.npz),
 (2) the C oracle (oracle/fluid_oracle.c) on seeded inputs, 2-D and 3-D,
 (3) size-independent properties at BASELINE.json's full sizes.

Bar: fp32 bit equality (+0 == -0) for every stencil / advection / Jacobi output; 1e-5 relative
for the Jacobi residual norm (a reduction whose summation order is unspecified)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, n_mismatch

pytestmark = pytest.mark.gpu

OPS = load_golden("ops_2d")
CASES = sorted(OPS)


@pytest.fixture(scope="module")
def fluid():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from fluidnet_cxx_b200.lib import fluid as f
    return f


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def host(t):
    return t.detach().cpu().numpy()


# ---------------------------------------------------------------------------------------------
# (1) golden vectors produced by the reference itself
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("method", ["maccormackFluidNet", "eulerFluidNet"])
@pytest.mark.parametrize("so", [0, 1])
def test_advect_scalar_golden(fluid, case, method, so):
    g = OPS[case]
    key = f"advectScalar_{method}_{so}"
    if key not in g:
        pytest.skip("reference itself raised on this input")
    src, U, flags = cu(g["rho"]), cu(g["U"]), cu(g["flags"])
    keep = (src.clone(), U.clone(), flags.clone())
    out = fluid.advectScalar(float(g["dt"]), src, U, flags, method, 1, bool(so), 0.6)
    assert n_mismatch(host(out), g[key]) == 0
    # inputs untouched, output is a new tensor (SURVEY §8b ownership)
    assert torch.equal(src, keep[0]) and torch.equal(U, keep[1]) and torch.equal(flags, keep[2])
    assert out.data_ptr() != src.data_ptr()


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("method", ["maccormackFluidNet", "eulerFluidNet"])
def test_advect_velocity_golden(fluid, case, method):
    g = OPS[case]
    U, flags = cu(g["U"]), cu(g["flags"])
    out = fluid.advectVelocity(float(g["dt"]), U, U, flags, method, 1, 0.6)
    assert n_mismatch(host(out), g[f"advectVelocity_{method}"]) == 0
    out = fluid.advectVelocity(float(g["dt"]), cu(g["orig2"]), U, flags, "maccormackFluidNet", 1, 0.75)
    assert n_mismatch(host(out), g["advectVelocity_orig2"]) == 0


@pytest.mark.parametrize("case", CASES)
def test_small_stencils_golden(fluid, case):
    g = OPS[case]
    dt = float(g["dt"])
    flags, rho, p = cu(g["flags"]), cu(g["rho"]), cu(g["p"])
    grav = torch.from_numpy(g["gravity"])
    U = cu(g["U"])
    r = fluid.addBuoyancy(U, flags, rho, grav, 0.05, dt)
    assert r is U and n_mismatch(host(U), g["addBuoyancy"]) == 0
    U = cu(g["U"])
    assert n_mismatch(host(fluid.addGravity(U, flags, grav, dt)), g["addGravity"]) == 0
    U = cu(g["U"])
    assert n_mismatch(host(fluid.setWallBcs(U, flags)), g["setWallBcs"]) == 0
    U = cu(g["U"])
    assert n_mismatch(host(fluid.velocityDivergence(U, flags)), g["velocityDivergence"]) == 0
    assert fluid.velocityUpdate(pressure=p, U=U, flags=flags) is None
    assert n_mismatch(host(U), g["velocityUpdate"]) == 0
    assert n_mismatch(host(fluid.flagsToOccupancy(flags)), g["flagsToOccupancy"]) == 0
    assert torch.equal(flags, cu(g["flags"]))  # flags never modified


@pytest.mark.parametrize("case", CASES)
def test_jacobi_golden(fluid, case):
    g = OPS[case]
    flags, div = cu(g["flags"]), cu(g["velocityDivergence"])
    for it in (4, 17):
        p, res = fluid.solveLinearSystemJacobi(flags, div, False, 0.0, it)
        assert n_mismatch(host(p), g[f"jacobi{it}_p"]) == 0
        assert res.dim() == 0
        assert abs(res.item() - g[f"jacobi{it}_res"]) <= 1e-5 * abs(g[f"jacobi{it}_res"])
    # tolerance-terminated run (generic one-iteration-per-launch path with the device-side test)
    p, res = fluid.solveLinearSystemJacobi(flags, div, False, float(g["jacobi_tol"]), 500)
    assert n_mismatch(host(p), g["jacobi_tol_p"]) == 0
    assert abs(res.item() - g["jacobi_tol_res"]) <= 1e-5 * abs(g["jacobi_tol_res"])
    with pytest.raises(RuntimeError):
        fluid.solveLinearSystemJacobi(flags, div, False, 0.0, 0)


def test_plume128_jacobi28_golden(fluid):
    """BASELINE.json configs[0]: 128x128 plume, Jacobi 28 -- the reference's own state after
    1, 2, 8 and 24 steps; the fused step and the op-by-op sequence must both reproduce it."""
    import importlib
    sim = importlib.import_module("fluidnet_cxx_b200.lib.simulate")
    G = load_golden("plume128_jacobi28")
    mconf = plume_mconf()
    for mode in ("fused", "ops"):
        bd = plume_state(fluid, 128, mconf)
        for k in ("UBC", "UBCInvMask", "densityBC", "densityBCInvMask", "flags"):
            assert n_mismatch(host(bd[k]), G["init"][k]) == 0, k
        for it in range(1, 25):
            if mode == "fused":
                sim.simulate(mconf, bd, None, "jacobi")
            else:
                sim._simulate_ops(mconf, bd, None, "jacobi", float(mconf["dt"]), False)
            if f"step{it}" in G:
                for k in ("p", "U", "density"):
                    assert n_mismatch(host(bd[k]), G[f"step{it}"][k]) == 0, (mode, it, k)


def plume_mconf(**over):
    m = {"dt": 0.1, "maccormackStrength": 0.6, "sampleOutsideFluid": False, "buoyancyScale": 0.25,
         "gravityScale": 0, "viscosity": 0, "correctScalar": False, "operatingDensity": 0.0,
         "gravityVec": {"x": 0, "y": -1, "z": 0}, "pTol": 0.0, "jacobiIter": 28, "simMethod": "jacobi",
         "injectionDensity": 0.1, "injectionVelocity": 2, "sourceRadius": 0.145}
    import os
    import yaml
    from conftest import ROOT
    y = os.path.join(ROOT, "tests", "golden", "plumeConfig.yaml")
    if os.path.exists(y):
        with open(y) as f:
            m.update(yaml.safe_load(f))
        m.update({"sampleOutsideFluid": False, "simMethod": "jacobi", "jacobiIter": 28, "pTol": 0.0})
    m.update(over)
    return m


def plume_state(fluid, res, mconf, depth=1):
    nc = 3 if depth > 1 else 2
    bd = {"p": torch.zeros(1, 1, depth, res, res, device="cuda"),
          "U": torch.zeros(1, nc, depth, res, res, device="cuda"),
          "flags": torch.zeros(1, 1, depth, res, res, device="cuda"),
          "density": torch.zeros(1, 1, depth, res, res, device="cuda")}
    fluid.emptyDomain(bd["flags"])
    fluid.createPlumeBCs(bd, mconf["injectionDensity"], mconf["injectionVelocity"], mconf["sourceRadius"])
    return bd


# ---------------------------------------------------------------------------------------------
# (2) C oracle on seeded inputs (larger 2-D, and 3-D where the reference has no runnable path)
# ---------------------------------------------------------------------------------------------
def random_case(seed, D, H, W, border, nboxes, vscale, with_empty=False):
    rng = np.random.RandomState(seed)
    is3d = D > 1
    f = np.full((1, 1, D, H, W), 1.0, np.float32)
    if border == "obstacle":
        f[..., 0, :] = 2; f[..., -1, :] = 2; f[..., :, 0] = 2; f[..., :, -1] = 2
        if is3d:
            f[:, :, 0] = 2; f[:, :, -1] = 2
    for _ in range(nboxes):
        h, w = rng.randint(1, 6), rng.randint(1, 6)
        y, x = rng.randint(1, H - 1 - h), rng.randint(1, W - 1 - w)
        if is3d:
            d = rng.randint(1, 4); z = rng.randint(1, D - 1 - d)
            f[0, 0, z:z + d, y:y + h, x:x + w] = 2
        else:
            f[0, 0, 0, y:y + h, x:x + w] = 2
    if with_empty:
        for _ in range(3):
            y, x = rng.randint(1, H - 4), rng.randint(1, W - 4)
            f[0, 0, :, y:y + 2, x:x + 3] = 4
    nc = 3 if is3d else 2
    U = (rng.randn(1, nc, D, H, W) * vscale).astype(np.float32)
    rho = rng.rand(1, 1, D, H, W).astype(np.float32)
    p = rng.randn(1, 1, D, H, W).astype(np.float32)
    return f, U, rho, p


ORACLE_CASES = [
    # seed, D, H, W, border, nboxes, vscale, dt, empty
    (10, 1, 96, 160, "obstacle", 12, 3.0, 0.3, False),
    (11, 1, 130, 70, "fluid", 6, 9.0, 0.5, True),
    (12, 1, 257, 131, "obstacle", 30, 1.0, 1.0, True),
    (13, 12, 20, 28, "obstacle", 6, 2.0, 0.4, False),
    (14, 16, 18, 14, "fluid", 4, 5.0, 0.5, True),
    (15, 9, 33, 21, "obstacle", 10, 1.0, 1.0, False),
]


@pytest.mark.parametrize("case", ORACLE_CASES, ids=lambda c: f"s{c[0]}_{c[1]}x{c[2]}x{c[3]}")
def test_ops_vs_oracle(fluid, oracle, case):
    seed, D, H, W, border, nboxes, vscale, dt, we = case
    f, U, rho, p = random_case(seed, D, H, W, border, nboxes, vscale, we)
    is3d = D > 1
    tf, tU, tr, tp = cu(f), cu(U), cu(rho), cu(p)
    for so in (False, True):
        for m in ("maccormackFluidNet", "eulerFluidNet"):
            ref = oracle.advectScalar(dt, rho, U, f, m, 1, so, 0.6)
            if oracle.advectScalar.last_errors:
                continue   # the reference would have asserted on this input
            got = fluid.advectScalar(dt, tr, tU, tf, m, 1, so, 0.6)
            assert n_mismatch(host(got), ref) == 0, (m, so)
    for m in ("maccormackFluidNet", "eulerFluidNet"):
        assert n_mismatch(host(fluid.advectVelocity(dt, tU, tU, tf, m, 1, 0.8)),
                          oracle.advectVelocity(dt, U, U, f, m, 1, 0.8)) == 0, m
    grav = [0.3, -0.25, 0.15 if is3d else 0.0]
    assert n_mismatch(host(fluid.addBuoyancy(tU.clone(), tf, tr, grav, 0.05, dt)),
                      oracle.addBuoyancy(U, f, rho, grav, 0.05, dt)) == 0
    assert n_mismatch(host(fluid.addGravity(tU.clone(), tf, grav, dt)), oracle.addGravity(U, f, grav, dt)) == 0
    assert n_mismatch(host(fluid.setWallBcs(tU.clone(), tf)), oracle.setWallBcs(U, f)) == 0
    div = oracle.velocityDivergence(U, f)
    assert n_mismatch(host(fluid.velocityDivergence(tU, tf)), div) == 0
    Uu = tU.clone()
    fluid.velocityUpdate(tp, Uu, tf)
    assert n_mismatch(host(Uu), oracle.velocityUpdate(p, U, f)) == 0
    for iters in (1, 7, 19):
        pr, rr = oracle.solveLinearSystemJacobi(f, div, is3d, 0.0, iters)
        pg, rg = fluid.solveLinearSystemJacobi(tf, cu(div), is3d, 0.0, iters)
        assert n_mismatch(host(pg), pr) == 0, iters
        assert abs(rg.item() - rr) <= 1e-5 * abs(rr) + 1e-30


def test_batched(fluid, oracle):
    """B > 1 (the C++ is batch-generic)."""
    parts = [random_case(20 + b, 1, 40, 56, "obstacle", 5, 2.0) for b in range(3)]
    f, U, rho, p = (np.concatenate([q[i] for q in parts], 0) for i in range(4))
    tf, tU, tr = cu(f), cu(U), cu(rho)
    assert n_mismatch(host(fluid.advectScalar(0.3, tr, tU, tf)), oracle.advectScalar(0.3, rho, U, f)) == 0
    assert n_mismatch(host(fluid.advectVelocity(0.3, tU, tU, tf)), oracle.advectVelocity(0.3, U, U, f)) == 0
    div = oracle.velocityDivergence(U, f)
    pg, rg = fluid.solveLinearSystemJacobi(tf, cu(div), False, 0.0, 11)
    pr, rr = oracle.solveLinearSystemJacobi(f, div, False, 0.0, 11)
    assert n_mismatch(host(pg), pr) == 0 and abs(rg.item() - rr) <= 1e-5 * rr


@pytest.mark.parametrize("D,res,boxes,vscale", [(24, 24, False, 0.8), (14, 36, True, 2.5)])
def test_step3d_vs_oracle(fluid, oracle, D, res, boxes, vscale):
    """3-D plume-like step (configs[4] semantics at a size the oracle finishes in seconds):
    fused step == op-by-op oracle sequence; second case: non-cubic grid, obstacle boxes inside the
    volume, velocities of several cells per step (line traces that stop at obstacles)."""
    import importlib
    sim = importlib.import_module("fluidnet_cxx_b200.lib.simulate")
    H = W = res
    mconf = plume_mconf(jacobiIter=9)
    bd = plume_state(fluid, W, mconf, depth=D)
    if boxes:
        bd["flags"][:, :, 4:9, 10:17, 20:27] = 2.0
        bd["flags"][:, :, 2:12, 25:30, 5:9] = 2.0
    rng = np.random.RandomState(3)
    bd["U"] = cu(rng.randn(1, 3, D, H, W) * vscale)
    bd["density"] = cu(rng.rand(1, 1, D, H, W))
    st = {k: host(v) for k, v in bd.items()}
    for _ in range(2):
        sim.simulate(mconf, bd, None, "jacobi")
        st = oracle_step(oracle, mconf, st)
    for k in ("p", "U", "density"):
        assert n_mismatch(host(bd[k]), st[k]) == 0, k


def oracle_step(orc, mconf, st):
    """simulate.py:28-171 (jacobi branch) restated on the oracle's per-op functions."""
    dt = float(mconf["dt"])
    f = st["flags"]
    rho = orc.advectScalar(dt, st["density"], st["U"], f, "maccormackFluidNet", 1, mconf["sampleOutsideFluid"],
                           mconf["maccormackStrength"])
    U = orc.advectVelocity(dt, st["U"], st["U"], f, "maccormackFluidNet", 1, mconf["maccormackStrength"])

    def cv(U, rho):
        if "UBC" in st:
            U = orc.setConstVals(U, st["UBCInvMask"], st["UBC"])
            rho = orc.setConstVals(rho, st["densityBCInvMask"], st["densityBC"])
        return U, rho
    U, rho = cv(U, rho)
    g = mconf["gravityVec"]
    grav = (np.array([g["x"], g["y"], g["z"]], np.float32) * np.float32(-mconf["buoyancyScale"])).astype(np.float32)
    U = orc.addBuoyancy(U, f, rho, grav, mconf["operatingDensity"], dt)
    U = orc.setWallBcs(U, f)
    U, rho = cv(U, rho)
    div = orc.velocityDivergence(U, f)
    p, _ = orc.solveLinearSystemJacobi(f, div, U.shape[1] == 3, 0.0, mconf["jacobiIter"])
    U = orc.velocityUpdate(p, U, f)
    U = orc.setWallBcs(U, f)
    U, rho = cv(U, rho)
    out = dict(st)
    out.update({"p": p, "U": U, "density": rho})
    return out


# ---------------------------------------------------------------------------------------------
# (3) properties at full size (BASELINE.json configs[3]: 4096x4096 plume, Jacobi 100)
# ---------------------------------------------------------------------------------------------
def test_full_size_properties(fluid):
    import importlib
    sim = importlib.import_module("fluidnet_cxx_b200.lib.simulate")
    res = 4096
    mconf = plume_mconf(jacobiIter=100)
    bd = plume_state(fluid, res, mconf)
    torch.manual_seed(0)
    bd["U"] = torch.randn_like(bd["U"]) * 0.5
    bd["density"] = torch.rand_like(bd["density"])
    flags0 = bd["flags"].clone()
    U0 = bd["U"].clone()
    # temporally blocked Jacobi == one-iteration-per-launch Jacobi, bit for bit
    div = fluid.velocityDivergence(U0, flags0)
    p_blk, r_blk = fluid.solveLinearSystemJacobi(flags0, div, False, 0.0, 100)
    p_gen, r_gen = fluid.solveLinearSystemJacobi(flags0, div, False, 1e-30, 100)   # residual test never fires
    assert torch.equal(p_blk, p_gen)
    assert abs(r_blk.item() - r_gen.item()) <= 1e-5 * r_gen.item()
    # p = 0 on the border ring; projection reduces the divergence
    assert p_blk[..., 0, :].abs().max() == 0 and p_blk[..., :, 0].abs().max() == 0
    assert p_blk[..., -1, :].abs().max() == 0 and p_blk[..., :, -1].abs().max() == 0
    U1 = U0.clone()
    fluid.velocityUpdate(p_blk, U1, flags0)
    # border ring untouched by the update (tfluids.cpp:1055-1061)
    assert torch.equal(U1[..., 0, :], U0[..., 0, :]) and torch.equal(U1[..., :, -1], U0[..., :, -1])
    d0 = fluid.velocityDivergence(U0, flags0)[..., 2:-2, 2:-2].pow(2).mean()
    d1 = fluid.velocityDivergence(U1, flags0)[..., 2:-2, 2:-2].pow(2).mean()
    assert d1 < 0.5 * d0
    # full fused step == op-by-op step, flags never modified, advected border = inlet BCs only
    bd2 = {k: v.clone() for k, v in bd.items()}
    sim.simulate(mconf, bd, None, "jacobi")
    sim._simulate_ops(mconf, bd2, None, "jacobi", float(mconf["dt"]), False)
    for k in ("p", "U", "density"):
        assert torch.equal(bd[k], bd2[k]), k
    assert torch.equal(bd["flags"], flags0)
    assert torch.isfinite(bd["U"]).all() and torch.isfinite(bd["p"]).all()


def test_full_size_properties_3d(fluid):
    """The same properties on the 256^3 grid of BASELINE.json configs[4] (3-D has no runnable reference: parity
    unpinned, so what can be held at full size is internal consistency): the vectorised 3-D Jacobi == the
    one-iteration generic kernel bit for bit, p = 0 on the border shell, the projection reduces the divergence,
    the fused 3-D step == the op-by-op sequence, flags untouched, everything finite."""
    import importlib
    sim = importlib.import_module("fluidnet_cxx_b200.lib.simulate")
    n = 256
    mconf = plume_mconf(jacobiIter=40)
    bd = plume_state(fluid, n, mconf, depth=n)
    torch.manual_seed(1)
    bd["U"] = torch.randn_like(bd["U"]) * 0.5
    bd["density"] = torch.rand_like(bd["density"])
    flags0, U0 = bd["flags"].clone(), bd["U"].clone()
    div = fluid.velocityDivergence(U0, flags0)
    p_vec, r_vec = fluid.solveLinearSystemJacobi(flags0, div, True, 0.0, 40)
    p_gen, r_gen = fluid.solveLinearSystemJacobi(flags0, div, True, 1e-30, 40)    # residual test never fires
    assert torch.equal(p_vec, p_gen)
    assert abs(r_vec.item() - r_gen.item()) <= 1e-5 * r_gen.item()
    del p_gen
    for face in (p_vec[:, :, 0], p_vec[:, :, -1], p_vec[..., 0, :], p_vec[..., -1, :], p_vec[..., 0], p_vec[..., -1]):
        assert face.abs().max() == 0
    U1 = U0.clone()
    fluid.velocityUpdate(p_vec, U1, flags0)
    d0 = div[..., 2:-2, 2:-2, 2:-2].pow(2).mean()
    d1 = fluid.velocityDivergence(U1, flags0)[..., 2:-2, 2:-2, 2:-2].pow(2).mean()
    assert d1 < 0.7 * d0
    del U1, p_vec, div
    bd2 = {k: v.clone() for k, v in bd.items()}
    sim.clear_graph_cache()
    sim.simulate(mconf, bd, None, "jacobi")
    sim._simulate_ops(mconf, bd2, None, "jacobi", float(mconf["dt"]), False)
    for k in ("p", "U", "density"):
        assert torch.equal(bd[k], bd2[k]), k
    assert torch.equal(bd["flags"], flags0)
    assert torch.isfinite(bd["U"]).all() and torch.isfinite(bd["p"]).all()
    sim.clear_graph_cache()


def test_viscosity_and_correct_scalar_golden():
    """fnx_add_viscosity / fnx_correct_scalar and the viscous, scalar-corrected branch of lib.simulate
    (simulate.py:67-69,79-81) against the reference's outputs, bit for bit."""
    import os
    from conftest import GOLDEN
    from fluidnet_cxx_b200.lib import fluid, simulate
    z = np.load(os.path.join(GOLDEN, "extras_2d.npz"))
    for name in ("a", "b"):
        f, U, rho = cu(z[f"{name}/flags"]), z[f"{name}/U"], z[f"{name}/rho"]
        for visc in (0.05, 1.3):
            Uv = cu(U)
            fluid.addViscosity(0.1, Uv, f, visc)
            assert n_mismatch(Uv.cpu().numpy(), z[f"{name}/visc_{visc}"]) == 0, (name, visc)
        r = cu(rho)
        fluid.correctScalar(0.1, r, cu(z[f"{name}/div"]), f)
        assert n_mismatch(r.cpu().numpy(), z[f"{name}/corrected"]) == 0, name
    mconf = plume_mconf(simMethod="jacobi", jacobiIter=12, viscosity=0.2, correctScalar=True)
    bd = {"p": torch.zeros(1, 1, 1, 48, 48, device="cuda")}
    for k in ("U", "density", "flags", "UBC", "UBCInvMask", "densityBC", "densityBCInvMask"):
        bd[k] = cu(z[f"sim/{k}0"])
    for it in range(1, 3):
        simulate(mconf, bd, None, "jacobi")
        for k in ("p", "U", "density"):
            assert n_mismatch(bd[k].cpu().numpy(), z[f"sim/step{it}_{k}"]) == 0, (it, k)


# ---------------------------------------------------------------------------------------------
# getCentered (the output / statistics step of plume.py:238-243; reference grid.py:7-30)
# ---------------------------------------------------------------------------------------------
def _get_centered_reference(U):
    """grid.py:7-30 restated on CPU tensors (pure slicing arithmetic: 0.5 * (a + b), last row / column /
    plane left at 0); when oracle/_ref is on the box the reference function itself is used instead."""
    try:
        import ref_loader
        if ref_loader.available():
            return ref_loader.load().fluid.getCentered(U)
    except Exception:      # noqa: BLE001 - fall through to the restatement
        pass
    is3d = U.size(2) > 1
    cx = torch.zeros_like(U[:, 0]); cy = torch.zeros_like(U[:, 0]); cz = torch.zeros_like(U[:, 0])
    cx[:, :, :, :-1] = 0.5 * (U[:, 0, :, :, 0:-1] + U[:, 0, :, :, 1:])
    cy[:, :, :-1, :] = 0.5 * (U[:, 1, :, 0:-1, :] + U[:, 1, :, 1:, :])
    if is3d:
        cz[:, :-1, :, :] = 0.5 * (U[:, 2, 0:-1, :, :] + U[:, 2, 1:, :, :])
    return torch.stack((cx, cy, cz), dim=1)


@pytest.mark.parametrize("shape", [(1, 2, 1, 33, 47), (2, 2, 1, 128, 128), (1, 3, 9, 17, 21), (1, 2, 1, 1024, 1024)])
def test_get_centered_vs_reference(fluid, shape):
    g = torch.Generator().manual_seed(shape[3])
    U = torch.randn(shape, generator=g)
    want = _get_centered_reference(U).numpy()
    got = host(fluid.getCentered(U.cuda()))
    assert got.shape == want.shape
    assert n_mismatch(got, want) == 0


def _plume_output_block(reffluid, bd):
    """plume.py:240-256 and :330-359 on CPU tensors, statement for statement (the drivers' output step):
    divergence, centred velocity, norm, NaN-filled obstacle cells, centred density / pressure gradients."""
    import numpy.ma as ma
    U, flags, density, pressure = bd["U"], bd["flags"], bd["density"], bd["p"]
    div = reffluid.velocityDivergence(U.clone(), flags.clone())
    vel = reffluid.getCentered(U.clone())
    mask = flags.eq(2)[:, 0].numpy().astype(float)
    out = {"div": div[:, 0].numpy()}
    for name, t in (("velx", vel[:, 0]), ("vely", vel[:, 1]), ("vel_norm", torch.norm(vel, dim=1)),
                    ("pressure", pressure[:, 0])):
        m = ma.array(t.numpy(), mask=mask)
        ma.set_fill_value(m, np.nan)
        out[name] = m.filled()
    b, _, d, h, w = pressure.shape
    for name, fld in (("gradRho", density), ("gradP", pressure)):
        c = fld.narrow(4, 1, w - 2).narrow(3, 1, h - 2)
        c = c.clone().expand(b, 2, d, h - 2, w - 2)
        m = c.clone()
        m[:, 0] = fld.narrow(4, 0, w - 2).narrow(3, 1, h - 2).squeeze(1)
        m[:, 1] = fld.narrow(4, 1, w - 2).narrow(3, 0, h - 2).squeeze(1)
        center = torch.zeros_like(vel)[:, 0:2].contiguous()
        center[:, 0:2, 0, 1:(h - 1), 1:(w - 1)] = reffluid.getCentered((c - m).contiguous())[:, 0:2, 0]
        out[name + "x"], out[name + "y"] = center[:, 0].numpy(), center[:, 1].numpy()
    return out


@pytest.mark.parametrize("hw", [(48, 64), (257, 131)])
def test_output_fields_vs_driver_output_block(fluid, hw):
    """fnx_output_fields (one kernel, one buffer, one device-to-host copy) against the drivers' own output block
    restated with the reference's velocityDivergence / getCentered (oracle/_ref when present, else this
    package's parity-pinned ops on the same tensors): every plane bit-exact, the norm to 1 ulp."""
    H, W = hw
    g = torch.Generator().manual_seed(H)
    bd = {"U": torch.randn(2, 2, 1, H, W, generator=g), "density": torch.rand(2, 1, 1, H, W, generator=g),
          "p": torch.randn(2, 1, 1, H, W, generator=g), "flags": torch.ones(2, 1, 1, H, W)}
    f = bd["flags"]
    f[..., 0, :] = 2; f[..., -1, :] = 2; f[..., :, 0] = 2; f[..., :, -1] = 2
    f[0, ..., H // 3:H // 3 + 7, W // 2:W // 2 + 9] = 2
    f[1, ..., 5:9, 3:30] = 2

    class _Ours:      # the same block driven by this package's (parity-pinned) ops, results moved to the CPU
        velocityDivergence = staticmethod(lambda U, fl: fluid.velocityDivergence(U.cuda(), fl.cuda()).cpu())
        getCentered = staticmethod(lambda U: fluid.getCentered(U.cuda()).cpu())
    reffluid = _Ours
    try:
        import ref_loader
        if ref_loader.available():
            reffluid = ref_loader.load().fluid
    except Exception:      # noqa: BLE001
        pass
    want = _plume_output_block(reffluid, bd)
    dev = {k: v.cuda() for k, v in bd.items()}
    arrays, pinned, dev_out = fluid.outputFieldsToHost(dev)
    assert set(arrays) == set(fluid.OUTPUT_PLANES)
    for name, ref in want.items():
        got = arrays[name][:, 0]
        ref = ref[:, 0] if ref.ndim == 4 else ref
        if name == "vel_norm":
            ok = np.isclose(got, ref, rtol=3e-7, atol=0, equal_nan=True)
            assert ok.all(), (name, int((~ok).sum()))
        else:
            assert n_mismatch(got, ref) == 0, name
    assert np.isnan(arrays["velx"][0, 0, 0, 0]) and not np.isnan(arrays["div"]).any()
    assert float(np.abs(arrays["velz"][~np.isnan(arrays["velz"])]).max()) == 0.0
    # unmasked variant, buffers reused
    arrays2, _, _ = fluid.outputFieldsToHost(dev, mask_obstacles=False, pinned=pinned, device_out=dev_out)
    assert not np.isnan(arrays2["velx"]).any()
    assert n_mismatch(arrays2["pressure"][:, 0], bd["p"][:, 0, 0].numpy()) == 0


# ---------------------------------------------------------------------------------------------
# (4) the fused 2-D step kernels (csrc/step2d.cu: forward pass staged in shared memory, tiled
# forces / divergence) against the op-by-op sequence AND the generic one-thread-per-cell kernels
# (FNX_STEP2D=0), on inputs that exercise every slow path: obstacles and Empty cells inside the
# domain (line-trace stops, fluid-aware interpolation), velocities of several cells per step
# (backward samples leaving the staged apron), odd sizes (partial tiles), both sampling modes,
# with and without imposed-value masks, convnet (no wall BCs, no divergence) and Jacobi variants.
# ---------------------------------------------------------------------------------------------
FUSED_CASES = [
    # seed, H, W, border, nboxes, vscale, dt, empty, sample_outside, masks
    (30, 96, 160, "obstacle", 12, 3.0, 0.3, False, False, True),
    (31, 130, 70, "fluid", 6, 9.0, 0.5, True, False, False),
    (32, 257, 131, "obstacle", 30, 1.0, 1.0, True, True, True),
    (33, 64, 64, "obstacle", 0, 0.5, 0.1, False, False, True),
    (34, 37, 203, "obstacle", 8, 20.0, 0.5, True, False, True),
    (35, 512, 384, "obstacle", 60, 2.0, 0.4, True, False, True),
    # mostly CLEAN tiles (interior fast path of k2_advect_clean) with a few obstacles / fast cells mixed in
    (36, 300, 420, "obstacle", 3, 1.0, 0.2, False, False, True),
    (37, 200, 330, "fluid", 2, 2.0, 0.3, True, True, True),
    (38, 1024, 1024, "obstacle", 5, 0.5, 0.1, False, False, False),
]


@pytest.mark.parametrize("case", FUSED_CASES, ids=lambda c: f"s{c[0]}_{c[1]}x{c[2]}")
def test_fused_2d_step_equals_ops_and_generic(fluid, case):
    import importlib
    import os
    sim = importlib.import_module("fluidnet_cxx_b200.lib.simulate")
    seed, H, W, border, nboxes, vscale, dt, we, so, masks = case
    f, U, rho, p = random_case(seed, 1, H, W, border, nboxes, vscale, we)
    mconf = plume_mconf(jacobiIter=11, dt=dt, sampleOutsideFluid=so, buoyancyScale=0.7, operatingDensity=0.05)
    base = {"p": cu(p), "U": cu(U), "flags": cu(f), "density": cu(rho)}
    if masks:
        rng = np.random.RandomState(seed + 100)
        inv = np.ones((1, 2, 1, H, W), np.float32); bc = np.zeros((1, 2, 1, H, W), np.float32)
        rinv = np.ones((1, 1, 1, H, W), np.float32); rbc = np.zeros((1, 1, 1, H, W), np.float32)
        rows = [0, 1, 2, 3, H // 2]
        for r in rows:
            sel = rng.rand(W) < 0.5
            inv[0, :, 0, r, sel] = 0; bc[0, 0, 0, r, sel] = 1.5; bc[0, 1, 0, r, sel] = -0.5
            rinv[0, 0, 0, r, sel] = 0; rbc[0, 0, 0, r, sel] = 0.8
        base.update({"UBC": cu(bc), "UBCInvMask": cu(inv), "densityBC": cu(rbc), "densityBCInvMask": cu(rinv)})
    sim.clear_graph_cache()
    runs = {}
    for mode in ("fused", "ops", "generic"):
        bd = {k: v.clone() for k, v in base.items()}
        for _ in range(2):
            if mode == "ops":
                sim._simulate_ops(mconf, bd, None, "jacobi", float(mconf["dt"]), False)
            else:
                # the library picks by grid size (fused kernels from ~1 M cells): force each one here
                os.environ["FNX_STEP2D"] = "0" if mode == "generic" else "1"
                try:
                    sim._simulate_fused(mconf, bd, None, "jacobi", float(mconf["dt"]), False)
                finally:
                    os.environ.pop("FNX_STEP2D", None)
        runs[mode] = {k: host(bd[k]) for k in ("p", "U", "density")}
    for k in ("p", "U", "density"):
        assert n_mismatch(runs["fused"][k], runs["ops"][k]) == 0, ("fused vs ops", k, n_mismatch(runs["fused"][k], runs["ops"][k]))
        assert n_mismatch(runs["fused"][k], runs["generic"][k]) == 0, ("fused vs generic", k)


def test_infinite_displacement_terminates(fluid, oracle):
    """A velocity whose displacement overflows fp32 (|u| dt > 1.8e19: length = inf, direction = delta/inf = 0) makes
    the reference's line-trace march stand still forever (calc_line_trace.cpp:310; there is no reference result).
    The kernels keep the start position instead -- this is the input that hung the 4-GPU run of round 1 once a
    freed CUDA-graph input buffer fed garbage to the advection.  Must terminate and agree with the oracle."""
    f, U, rho, p = random_case(41, 1, 70, 90, "obstacle", 4, 0.5, False)
    U[0, 0, 0, 30, 40] = 3.0e30
    U[0, 1, 0, 31, 41] = -2.5e31
    U[0, 0, 0, 50, 20] = np.inf
    tf, tU, tr = cu(f), cu(U), cu(rho)
    got = fluid.advectScalar(0.1, tr, tU, tf, "maccormackFluidNet", 1, False, 0.6)
    torch.cuda.synchronize()
    ref = oracle.advectScalar(0.1, rho, U, f, "maccormackFluidNet", 1, False, 0.6)
    assert n_mismatch(host(got), ref) == 0
    # and through the fused step kernels (generic path: these tiles are not "clean")
    import importlib
    sim = importlib.import_module("fluidnet_cxx_b200.lib.simulate")
    mconf = plume_mconf(jacobiIter=4)
    bd = {"p": cu(p), "U": tU.clone(), "flags": tf, "density": tr.clone()}
    bd2 = {k: v.clone() for k, v in bd.items()}
    sim._simulate_fused(mconf, bd, None, "jacobi", 0.1, False)
    sim._simulate_ops(mconf, bd2, None, "jacobi", 0.1, False)
    torch.cuda.synchronize()
    assert n_mismatch(host(bd["density"]), host(bd2["density"])) == 0
