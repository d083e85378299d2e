"""The slab-decomposed step on ONE GPU: several virtual ranks run fluidnet_cxx_b200.lib.distributed.
simulate_distributed in threads of this process (an in-process transport stands in for NCCL), each
launching the real row-window kernels.  Jacobi path: the gathered result equals the single-GPU fused
step bit for bit.  ScaleNet path: within the CNN tolerance (the split-fp16 activation scales depend
on the tensor-wide max, which differs between a window and the whole grid)."""
import importlib
import threading

import numpy as np
import pytest
import torch
import torch.distributed as dist

pytestmark = pytest.mark.gpu


class ThreadComm:
    """exchange / all_reduce / all_gather between virtual ranks living in threads of one process"""

    def __init__(self, world):
        self.world = world
        self.bar = threading.Barrier(world)
        self.box = {}

    def _sync(self):
        torch.cuda.synchronize()
        self.bar.wait()

    def exchange_rows(self, dec, sends):
        for peer, buf in sends.items():
            self.box[(dec.rank, peer)] = buf
        self._sync()
        out = {peer: self.box[(peer, dec.rank)].clone() for peer in sends}
        self._sync()
        return out

    def all_reduce(self, dec, t, op):
        self.box[("r", dec.rank)] = t.clone()
        self._sync()
        parts = torch.stack([self.box[("r", r)] for r in range(self.world)])
        t.copy_(parts.max(0).values if op == dist.ReduceOp.MAX else parts.sum(0))
        self._sync()

    def all_gather(self, dec, mine):
        self.box[("g", dec.rank)] = mine
        self._sync()
        parts = [self.box[("g", r)].clone() for r in range(self.world)]
        self._sync()
        return parts


def locked_ops():
    """Virtual ranks share one process, hence one set of per-device scratch workspaces: run each
    per-window op (its kernels + the workspace it owns) under a lock, to completion."""
    from fluidnet_cxx_b200.lib.distributed import CudaLocalOps
    lock = threading.Lock()

    class LockedOps(CudaLocalOps):
        pass

    def wrap(name):
        inner = getattr(CudaLocalOps, name)

        def call(self, *a, **kw):
            with lock:
                out = inner(self, *a, **kw)
                torch.cuda.synchronize()
                return out
        return call
    for name in ("advect_forces_div", "jacobi", "jacobi_resid", "project", "set_const", "cnn"):
        setattr(LockedOps, name, wrap(name))
    return LockedOps      # one instance per virtual rank: each rank owns its pool of output buffers


def run_virtual(world, ghost, mconf, state, net, method, steps):
    from fluidnet_cxx_b200.lib.distributed import SlabDecomposition, simulate_distributed
    comm = ThreadComm(world)
    ops_cls = locked_ops()
    H = state["flags"].shape[3]
    results, errors = [None] * world, []

    def work(rank):
        try:
            torch.cuda.set_device(0)
            dec = SlabDecomposition(H, ghost, rank=rank, world=world, comm=comm)
            ops = ops_cls()
            bd = {k: dec.scatter(v) for k, v in state.items()}
            outs = []
            with torch.no_grad():
                for _ in range(steps):
                    simulate_distributed(mconf, bd, net, method, dec, ops=ops)
                    outs.append({k: dec.gather(bd[k]) for k in ("p", "U", "density")})
            results[rank] = outs
        except Exception as e:      # noqa: BLE001 - surface the failure in the main thread
            errors.append(e)
            comm.bar.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results[0]


def make_state(fluid, H, W, mconf, seed=11):
    bd = {"p": torch.zeros(1, 1, 1, H, W, device="cuda"), "U": torch.zeros(1, 2, 1, H, W, device="cuda"),
          "flags": torch.zeros(1, 1, 1, H, W, device="cuda"), "density": torch.zeros(1, 1, 1, H, W, device="cuda")}
    fluid.emptyDomain(bd["flags"])
    bd["flags"][..., H // 2 - 5:H // 2 + 6, W // 3:W // 3 + 9] = 2.0       # obstacle across a slab boundary
    fluid.createPlumeBCs(bd, mconf["injectionDensity"], mconf["injectionVelocity"], mconf["sourceRadius"])
    g = torch.Generator(device="cuda").manual_seed(seed)
    bd["U"] = torch.randn(bd["U"].shape, device="cuda", generator=g) * 0.5
    bd["density"] = torch.rand(bd["density"].shape, device="cuda", generator=g)
    return bd


@pytest.mark.parametrize("world,H,W,ghost,iters", [(2, 256, 192, 48, 100), (4, 256, 160, 24, 28)])
def test_virtual_ranks_jacobi_bit_exact(world, H, W, ghost, iters):
    from fluidnet_cxx_b200.lib import fluid
    sim = importlib.import_module("fluidnet_cxx_b200.lib.simulate")
    from test_gpu_parity import plume_mconf
    mconf = plume_mconf(simMethod="jacobi")
    mconf["jacobiIter"] = iters
    state = make_state(fluid, H, W, mconf)
    ref_bd = {k: v.clone() for k, v in state.items()}
    ref = []
    for _ in range(3):
        sim._simulate_fused(mconf, ref_bd, None, "jacobi", float(mconf["dt"]), False)
        ref.append({k: ref_bd[k].clone() for k in ("p", "U", "density")})
    got = run_virtual(world, ghost, mconf, state, None, "jacobi", 3)
    for i in range(3):
        for k in ("p", "U", "density"):
            assert torch.equal(got[i][k], ref[i][k]), (i, k, int((got[i][k] != ref[i][k]).sum()))


@pytest.mark.parametrize("world,ghost,stop_at", [(2, 24, 19), (4, 24, 16)])
def test_virtual_ranks_jacobi_residual_terminated(world, ghost, stop_at, monkeypatch):
    """pTol > 0 across slabs: the decomposed solve stops at the iteration the single-GPU solver stops at (chunks of
    8 iterations here: 19 = inside a chunk, 16 = a chunk's last iteration) and the step is bit-identical."""
    from fluidnet_cxx_b200.lib import fluid
    sim = importlib.import_module("fluidnet_cxx_b200.lib.simulate")
    from test_gpu_parity import plume_mconf
    H, W = 256, 160
    mconf = plume_mconf(simMethod="jacobi")
    mconf["jacobiIter"] = 60
    state = make_state(fluid, H, W, mconf)
    # the divergence the step solves for, then a tolerance between the residuals of iterations stop_at-1 and stop_at
    seen = {}
    real = fluid.solveLinearSystemJacobi

    def spy(flags, div, **kw):
        seen["div"] = div.clone()
        return real(flags=flags, div=div, **kw)
    monkeypatch.setattr(fluid, "solveLinearSystemJacobi", spy)
    probe = {k: v.clone() for k, v in state.items()}
    sim._simulate_ops(dict(mconf, pTol=0.0, jacobiIter=1), probe, None, "jacobi", float(mconf["dt"]), False)
    monkeypatch.setattr(fluid, "solveLinearSystemJacobi", real)
    r = [float(real(flags=state["flags"], div=seen["div"], is_3d=False, p_tol=0.0, max_iter=n)[1])
         for n in (stop_at - 1, stop_at)]
    assert r[1] < r[0]
    mconf["pTol"] = 0.5 * (r[0] + r[1])
    ref_bd = {k: v.clone() for k, v in state.items()}
    sim.simulate(mconf, ref_bd, None, "jacobi")
    p_fixed = real(flags=state["flags"], div=seen["div"], is_3d=False, p_tol=0.0, max_iter=stop_at)[0]
    assert torch.equal(ref_bd["p"], p_fixed)                 # the single-GPU solver stopped where intended
    got = run_virtual(world, ghost, mconf, state, None, "jacobi", 1)[0]
    for k in ("p", "U", "density"):
        assert torch.equal(got[k], ref_bd[k]), (k, int((got[k] != ref_bd[k]).sum()))


def test_virtual_ranks_convnet():
    from fluidnet_cxx_b200.lib import fluid
    from fluidnet_cxx_b200.lib.pretrained import load_scalenet
    sim = importlib.import_module("fluidnet_cxx_b200.lib.simulate")
    from test_gpu_parity import plume_mconf
    model, mconf_net = load_scalenet("cuda")
    mconf = dict(mconf_net)
    mconf.update(plume_mconf(simMethod="convnet"))
    model.mconf = mconf
    model.scale.mconf = mconf
    H, W = 256, 128
    state = make_state(fluid, H, W, mconf)
    ref_bd = {k: v.clone() for k, v in state.items()}
    ref = []
    with torch.no_grad():
        for _ in range(2):
            sim._simulate_fused(mconf, ref_bd, model, "convnet", float(mconf["dt"]), False)
            ref.append({k: ref_bd[k].clone() for k in ("p", "U", "density")})
    got = run_virtual(2, 64, mconf, state, model, "convnet", 2)
    for i in range(2):
        assert torch.equal(got[i]["density"], ref[i]["density"]) or i > 0      # advection is bit-exact
        for k in ("p", "U", "density"):
            err = float((got[i][k] - ref[i][k]).abs().max() / ref[i][k].abs().max())
            assert err < 2e-5 * (i + 1), (i, k, err)


@pytest.mark.parametrize("world,per_x", [(2, False), (4, False), (2, True)])
def test_virtual_ranks_convnet_periodic_seam(world, per_x):
    """Rayleigh-Taylor's periodic-y seam across slabs (*_saved.py:123-132, 228-237): row H-1 lives on the last rank,
    row 1 on the first.  Decomposed == single-GPU step within the CNN tolerance; a missing seam is an O(1) error in
    row 1 of Ux."""
    from fluidnet_cxx_b200.lib import fluid
    from fluidnet_cxx_b200.lib.pretrained import load_scalenet
    sim = importlib.import_module("fluidnet_cxx_b200.lib.simulate")
    from test_gpu_parity import plume_mconf
    model, mconf_net = load_scalenet("cuda")
    mconf = dict(mconf_net)
    mconf.update(plume_mconf(simMethod="convnet"))
    mconf["periodic-y"], mconf["periodic-x"] = True, per_x
    model.mconf = mconf
    model.scale.mconf = mconf
    H, W = 256, 128
    # no inflow masks (as the Rayleigh-Taylor driver): the plume's velocity mask pins rows 0..3 and would hide the seam
    state = {k: v for k, v in make_state(fluid, H, W, mconf).items() if k in ("p", "U", "flags", "density")}
    ref_bd = {k: v.clone() for k, v in state.items()}
    ref = []
    sim.clear_graph_cache()
    with torch.no_grad():
        for _ in range(2):
            sim.simulate(mconf, ref_bd, model, "convnet")
            ref.append({k: ref_bd[k].clone() for k in ("p", "U", "density")})
    # the seam is there in the single-GPU step: without it row 1 of Ux would be setWallBcs' value
    noseam = dict(mconf)
    noseam["periodic-y"] = False
    model.mconf = noseam
    nb = {k: v.clone() for k, v in state.items()}
    with torch.no_grad():
        sim.simulate(noseam, nb, model, "convnet")
    assert float((nb["U"][:, 0, :, 1] - ref[0]["U"][:, 0, :, 1]).abs().max()) > 1e-2
    model.mconf = mconf
    got = run_virtual(world, 64, mconf, state, model, "convnet", 2)
    for i in range(2):
        for k in ("p", "U", "density"):
            err = float((got[i][k] - ref[i][k]).abs().max() / ref[i][k].abs().max())
            assert err < 2e-5 * (i + 1), (i, k, err)
        seam = float((got[i]["U"][:, 0, :, 1] - ref[i]["U"][:, 0, :, 1]).abs().max())
        assert seam < 1e-4 * float(ref[i]["U"].abs().max()), (i, seam)


@pytest.mark.parametrize("name,world", [("plume512_scalenet", 2), ("plume512_scalenet", 4), ("rt1024_scalenet", 4)])
def test_virtual_ranks_convnet_step_vs_reference_cpu(name, world):
    """The SLAB-DECOMPOSED ScaleNet step of a bench workload at full size against the reference's own lib.simulate
    on CPU (oracle/_ref): the north star's 1e-5 relative (max norm) on p and U, density bit-exact -- the window-local
    activation scales of the split-fp16 convolutions and the all-reduced std stay inside the bar (rt1024: through
    the cross-slab periodic-y seam)."""
    from test_gpu_cnn import RTOL, _bench_state, n_bad, rel_err
    from fluidnet_cxx_b200.lib.pretrained import load_scalenet
    model, _ = load_scalenet("cuda")
    reflib, ref_net, mconf, bd = _bench_state(name)
    ref_net.mconf = mconf
    ref_net.scale.mconf = mconf
    state = {k: v.cuda() for k, v in bd.items()}
    with torch.no_grad():
        reflib.simulate(mconf, bd, ref_net, "convnet")
    model.mconf = mconf
    model.scale.mconf = mconf
    got = run_virtual(world, 64, mconf, state, model, "convnet", 1)[0]
    assert n_bad(got["density"].cpu().numpy(), bd["density"].numpy()) == 0, name
    assert rel_err(got["p"].cpu().numpy(), bd["p"].numpy()) < RTOL, name
    assert rel_err(got["U"].cpu().numpy(), bd["U"].numpy()) < RTOL, name


@pytest.mark.parametrize("method", ["jacobi", "convnet"])
def test_graphed_stepper_single_rank(method):
    """GraphedDistributedStep on one rank (no communication, one graph) == the fused single-GPU step"""
    from fluidnet_cxx_b200.lib import fluid
    from fluidnet_cxx_b200.lib.distributed import GraphedDistributedStep, SlabDecomposition
    from fluidnet_cxx_b200.lib.pretrained import load_scalenet
    sim = importlib.import_module("fluidnet_cxx_b200.lib.simulate")
    from test_gpu_parity import plume_mconf
    model, mconf_net = load_scalenet("cuda")
    mconf = dict(mconf_net)
    mconf.update(plume_mconf(simMethod=method))
    model.mconf = mconf
    model.scale.mconf = mconf
    H, W = 128, 160
    state = make_state(fluid, H, W, mconf)
    ref_bd = {k: v.clone() for k, v in state.items()}
    dec = SlabDecomposition(H, 0, rank=0, world=1)
    stepper = GraphedDistributedStep(mconf, state, model, method, dec)
    assert stepper.graphed, stepper.capture_error
    for i in range(3):
        with torch.no_grad():
            sim._simulate_fused(mconf, ref_bd, model, method, float(mconf["dt"]), False)
        stepper.step()
        for k in ("p", "U", "density"):
            if method == "jacobi":
                assert torch.equal(stepper.state[k], ref_bd[k]), (i, k)
            else:   # the std is formed from fp64 partial sums here, by the two-pass kernel there
                err = float((stepper.state[k] - ref_bd[k]).abs().max() / ref_bd[k].abs().max())
                assert err < 1e-5, (i, k, err)
    assert stepper.verify() == 0.0


def test_virtual_ranks_convnet_3d_slicewise():
    """BASELINE configs[4] as this package defines it (slice-wise CNN projection, parity unpinned): slabs along D,
    two virtual ranks with the real kernels == the single-GPU 3-D convnet step (CNN tolerance)."""
    from fluidnet_cxx_b200.lib import fluid
    from fluidnet_cxx_b200.lib.distributed import GHOST_CONVNET_3D, SlabDecomposition, simulate_distributed
    from fluidnet_cxx_b200.lib.pretrained import load_scalenet
    sim = importlib.import_module("fluidnet_cxx_b200.lib.simulate")
    from test_gpu_parity import plume_mconf, plume_state
    model, mconf_net = load_scalenet("cuda")
    mconf = dict(mconf_net)
    mconf.update(plume_mconf(simMethod="convnet"))
    model.mconf = mconf
    model.scale.mconf = mconf
    D, res, world = 48, 40, 2
    state = plume_state(fluid, res, mconf, depth=D)
    g = torch.Generator(device="cuda").manual_seed(21)
    state["U"] = torch.randn(state["U"].shape, device="cuda", generator=g) * 0.3
    state["density"] = torch.rand(state["density"].shape, device="cuda", generator=g)
    ref_bd = {k: v.clone() for k, v in state.items()}
    sim.clear_graph_cache()
    ref = []
    with torch.no_grad():
        for _ in range(2):
            sim._simulate_fused(mconf, ref_bd, model, "convnet", float(mconf["dt"]), False)
            ref.append({k: ref_bd[k].clone() for k in ("p", "U", "density")})
    comm = ThreadComm(world)
    ops_cls = locked_ops()
    results, errors = [None] * world, []

    def work(rank):
        try:
            torch.cuda.set_device(0)
            dec = SlabDecomposition(D, GHOST_CONVNET_3D, rank=rank, world=world, comm=comm, axis=2)
            ops = ops_cls()
            bd = {k: dec.scatter(v) for k, v in state.items()}
            outs = []
            with torch.no_grad():
                for _ in range(2):
                    simulate_distributed(mconf, bd, model, "convnet", dec, ops=ops)
                    outs.append({k: dec.gather(bd[k]) for k in ("p", "U", "density")})
            results[rank] = outs
        except Exception as e:      # noqa: BLE001
            errors.append(e)
            comm.bar.abort()
    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    got = results[0]
    for i in range(2):
        for k in ("p", "U", "density"):
            err = float((got[i][k] - ref[i][k]).abs().max() / ref[i][k].abs().max())
            assert err < 2e-5 * (i + 1), (i, k, err)

