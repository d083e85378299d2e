"""Slab decomposition + halo exchange (fluidnet_cxx_b200.lib.distributed) on CPU: world_size 2 and 4
over gloo, with the C oracle as the per-slab compute.  The decomposed run must reproduce the
single-domain oracle step BIT FOR BIT on the owned rows (same per-cell arithmetic, ghost rows wide
enough for every stage's dependency radius)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

MCONF = {"dt": 0.1, "maccormackStrength": 0.6, "sampleOutsideFluid": False, "buoyancyScale": 0.25,
         "gravityScale": 0, "viscosity": 0, "correctScalar": False, "operatingDensity": 0.0,
         "gravityVec": {"x": 0, "y": -1, "z": 0}, "pTol": 0.0, "jacobiIter": 20}
GRAV = np.array([-0.0, 0.25, -0.0], np.float32)


def jacobi_numpy(flags, div, p, iters):
    """fluids_init.cpp:809-1004 continued from p (2-D), same fp32 operation order as the oracle."""
    f = flags[0, 0, 0]
    dv = div[0, 0, 0]
    p = np.zeros_like(dv) if p is None else p[0, 0, 0].copy()
    H, W = f.shape
    obst = f == 2
    interior = np.zeros_like(obst)
    interior[1:-1, 1:-1] = True
    act = interior & ~obst
    for _ in range(iters):
        pc = p
        def nb(dy, dx):
            sh = np.zeros_like(pc)
            ob = np.zeros_like(obst)
            ys, yd = (slice(1, H), slice(0, H - 1)) if dy == 1 else ((slice(0, H - 1), slice(1, H)) if dy == -1 else (slice(None), slice(None)))
            xs, xd = (slice(1, W), slice(0, W - 1)) if dx == 1 else ((slice(0, W - 1), slice(1, W)) if dx == -1 else (slice(None), slice(None)))
            sh[yd, xd] = pc[ys, xs]
            ob[yd, xd] = obst[ys, xs]
            return np.where(ob, pc, sh)
        s = nb(0, -1) + nb(0, 1)
        s = s + nb(-1, 0)
        s = s + nb(1, 0)
        s = s + np.float32(0) + np.float32(0) + dv
        p = np.where(act, s / np.float32(4), np.float32(0)).astype(np.float32)
    return p[None, None, None]


class OracleOps:
    """per-slab compute = the oracle's op-by-op restatement of simulate.py:28-171"""

    def __init__(self):
        sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
        import oracle
        oracle.build()
        self.o = oracle

    @staticmethod
    def _np(t):
        return t.numpy()

    def _const(self, U, rho, bd):
        o = self.o
        return (torch.from_numpy(o.setConstVals(U, self._np(bd["UBCInvMask"]), self._np(bd["UBC"]))),
                torch.from_numpy(o.setConstVals(rho, self._np(bd["densityBCInvMask"]), self._np(bd["densityBC"]))))

    @staticmethod
    def _win(new, rows):
        """the kernels write the window rows only: poison everything else (NaN would be ideal but the
        real kernels keep outside rows finite -- zeros -- so that stale velocities cannot stretch a
        line trace; do the same)"""
        out = torch.zeros_like(new)
        out[:, :, :, rows[0]:rows[1]] = new[:, :, :, rows[0]:rows[1]]
        return out

    def p_buffer(self, like, current_p):
        return torch.zeros_like(like)

    def advect_forces_div(self, mconf, dt, bd, want_div, wall_bcs, rows):
        o = self.o
        f, U0, r0 = self._np(bd["flags"]), self._np(bd["U"]), self._np(bd["density"])
        rho = o.advectScalar(dt, r0, U0, f, "maccormackFluidNet", 1, False, mconf["maccormackStrength"])
        U = o.advectVelocity(dt, U0, U0, f, "maccormackFluidNet", 1, mconf["maccormackStrength"])
        U, rho = self._const(U, rho, bd)
        U = o.addBuoyancy(U.numpy(), f, rho.numpy(), GRAV, 0.0, dt)
        if wall_bcs:
            U = o.setWallBcs(U, f)
        U, rho = self._const(U, rho.numpy(), bd)
        div = torch.from_numpy(o.velocityDivergence(U.numpy(), f)) if want_div else None
        return self._win(rho, rows), self._win(U, rows), (self._win(div, rows) if want_div else None)

    def jacobi(self, flags, div, p_init, iters, rows):
        p = torch.from_numpy(jacobi_numpy(flags.numpy(), div.numpy(), None if p_init is None else p_init.numpy(), iters))
        return self._win(p, rows)

    def jacobi_resid(self, flags, div, p_init, iters, rows, own):
        ssq = torch.zeros((iters, 1), dtype=torch.float64)
        p = p_init
        for it in range(iters):
            new = self.jacobi(flags, div, p, 1, rows)
            d = (new - (p if p is not None else torch.zeros_like(new)))[:, :, :, own[0]:own[1]].double()
            ssq[it, 0] = float((d * d).sum())
            p = new
        return p, ssq

    def set_const(self, x, inv_mask, bc):
        x.mul_(inv_mask).add_(bc)

    def cnn(self, net, U, flags, scale, prewall=False):
        """a stand-in "network" with the wrapper's structure (*_saved.py:135-232): normalise, divergence, a 3-iteration
        Jacobi as the pressure model (dependency radius 3, translation invariant like the CNN), velocity update,
        un-normalise, setWallBcs; prewall also returns the field before setWallBcs (the periodic seam's source)"""
        o = self.o
        f = flags.numpy()
        u = (U / scale).numpy()
        pn = jacobi_numpy(f, o.velocityDivergence(u, f), None, 3)
        ut = torch.from_numpy(o.velocityUpdate(pn, u, f)) * scale
        uo = torch.from_numpy(o.setWallBcs(ut.numpy(), f))
        p = torch.from_numpy(pn) * scale
        return (p, uo, ut) if prewall else (p, uo)

    def project(self, p, U, bd, rows):
        o = self.o
        f = self._np(bd["flags"])
        Un = o.setWallBcs(o.velocityUpdate(p.numpy(), U.numpy(), f), f)
        return self._win(torch.from_numpy(o.setConstVals(Un, self._np(bd["UBCInvMask"]), self._np(bd["UBC"]))), rows)


def global_state(H, W, seed):
    rng = np.random.RandomState(seed)
    flags = np.ones((1, 1, 1, H, W), np.float32)
    flags[..., 0, :] = flags[..., -1, :] = flags[..., :, 0] = flags[..., :, -1] = 2
    flags[..., H // 2 - 3:H // 2 + 4, W // 3:W // 3 + 6] = 2      # an obstacle box straddling the slab boundary
    flags[..., 5:9, W // 2:W // 2 + 3] = 2
    U = (rng.standard_normal((1, 2, 1, H, W)) * 0.5).astype(np.float32)
    rho = rng.random_sample((1, 1, 1, H, W)).astype(np.float32)
    UBC = np.zeros_like(U); UBCInv = np.ones_like(U)
    rBC = np.zeros_like(rho); rBCInv = np.ones_like(rho)
    UBC[:, 1, :, 0:4, W // 4:W // 2] = 2.0; UBCInv[:, :, :, 0:4, W // 4:W // 2] = 0.0
    rBC[:, :, :, 0:4, W // 4:W // 2] = 0.1; rBCInv[:, :, :, 0:4, W // 4:W // 2] = 0.0
    return {"p": np.zeros_like(rho), "U": U, "flags": flags, "density": rho, "UBC": UBC, "UBCInvMask": UBCInv,
            "densityBC": rBC, "densityBCInvMask": rBCInv}


def single_domain_steps(H, W, seed, steps):
    ops = OracleOps()
    bd = {k: torch.from_numpy(v.copy()) for k, v in global_state(H, W, seed).items()}
    out = []
    for _ in range(steps):
        rho, U, div = ops.advect_forces_div(MCONF, MCONF["dt"], bd, True, True, (0, H))
        p, _ = ops.o.solveLinearSystemJacobi(bd["flags"].numpy(), div.numpy(), False, 0.0, MCONF["jacobiIter"])
        U = ops.project(torch.from_numpy(p), U, bd, (0, H))
        bd["U"], bd["density"], bd["p"] = U, rho, torch.from_numpy(p)
        out.append({k: bd[k].numpy().copy() for k in ("p", "U", "density")})
    return out


def _worker(rank, world, port, H, W, ghost, steps, seed, result_path, mconf=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
        from fluidnet_cxx_b200.lib.distributed import SlabDecomposition, simulate_distributed
        mconf = mconf or MCONF
        dec = SlabDecomposition(H, ghost)
        ops = OracleOps()
        bd = {k: dec.scatter(torch.from_numpy(v)) for k, v in global_state(H, W, seed).items()}
        results = []
        for _ in range(steps):
            simulate_distributed(mconf, bd, None, "jacobi", dec, ops=ops)
            results.append({k: dec.gather(bd[k]).numpy() for k in ("p", "U", "density")})
            if "jacobi_iterations" in bd:
                results[-1]["iters"] = np.array([bd["jacobi_iterations"]])
        if rank == 0:
            np.savez(result_path, **{f"{i}/{k}": v for i, r in enumerate(results) for k, v in r.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,H,ghost", [(2, 64, 24), (4, 128, 32)])
def test_slab_decomposition_bit_exact(tmp_path, world, H, ghost):
    W, steps, seed = 48, 2, 7
    ref = single_domain_steps(H, W, seed, steps)
    # the numpy Jacobi continuation used by the slab ops == the oracle's solver from p = 0
    ops = OracleOps()
    st = global_state(H, W, seed)
    div = ops.o.velocityDivergence(st["U"], st["flags"])
    p_ref, _ = ops.o.solveLinearSystemJacobi(st["flags"], div, False, 0.0, 9)
    assert np.array_equal(jacobi_numpy(st["flags"], div, jacobi_numpy(st["flags"], div, None, 4), 5), p_ref)
    out = str(tmp_path / "res.npz")
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, H, W, ghost, steps, seed, out), nprocs=world, join=True)
    z = np.load(out)
    for i in range(steps):
        for k in ("p", "U", "density"):
            got, want = z[f"{i}/{k}"], ref[i][k]
            bad = int(np.sum(~((got == want) | (np.isnan(got) & np.isnan(want)))))
            assert bad == 0, f"step {i} field {k}: {bad} cells differ from the single-domain step"


def _oracle_iterations(flags, div, p_tol, max_iter):
    """the iteration the residual-terminated solver stops at: the smallest n whose fixed-count solve the
    tolerance solve reproduces (the oracle does not report its count)"""
    import oracle
    want, _ = oracle.solveLinearSystemJacobi(flags, div, False, p_tol, max_iter)
    for n in range(1, max_iter + 1):
        got, _ = oracle.solveLinearSystemJacobi(flags, div, False, 0.0, n)
        if np.array_equal(got, want):
            return n, want
    raise AssertionError("no fixed count reproduces the tolerance solve")


@pytest.mark.parametrize("p_tol,world,H", [(2.7, 2, 64), (1.4, 2, 64), (2.2, 4, 128)])
def test_slab_residual_terminated_jacobi(tmp_path, p_tol, world, H):
    """pTol > 0 across slabs (fluids_init.cpp:958-990): one all-reduce(sum) of the per-iteration squared residuals
    per chunk of iterations; the decomposed solve stops at the SAME iteration as the single-domain oracle and gives
    the same p, U bit for bit (stopping inside a chunk, not on its last iteration, included)."""
    W, ghost, seed = 48, 20, 7                         # ghost 20 -> chunks of 7 iterations
    mconf = dict(MCONF, pTol=p_tol, jacobiIter=60)
    ops = OracleOps()
    bd = {k: torch.from_numpy(v.copy()) for k, v in global_state(H, W, seed).items()}
    rho, U, div = ops.advect_forces_div(mconf, mconf["dt"], bd, True, True, (0, H))
    n_ref, p_ref = _oracle_iterations(bd["flags"].numpy(), div.numpy(), p_tol, mconf["jacobiIter"])
    assert 1 < n_ref < mconf["jacobiIter"], n_ref        # the tolerance, not the cap, ends the solve
    U_ref = ops.project(torch.from_numpy(p_ref), U, bd, (0, H)).numpy()
    out = str(tmp_path / "res.npz")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, H, W, ghost, 1, seed, out, mconf), nprocs=world, join=True)
    z = np.load(out)
    assert int(z["0/iters"][0]) == n_ref
    assert np.array_equal(z["0/p"], p_ref) and np.array_equal(z["0/U"], U_ref)


def _seam_worker(rank, world, port, H, W, ghost, seed, mconf, netconf, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
        import types
        from fluidnet_cxx_b200.lib.distributed import SlabDecomposition, simulate_distributed
        dec = SlabDecomposition(H, ghost)
        bd = {k: dec.scatter(torch.from_numpy(v)) for k, v in global_state(H, W, seed).items()}
        simulate_distributed(mconf, bd, types.SimpleNamespace(mconf=netconf), "convnet", dec, ops=OracleOps())
        got = {k: dec.gather(bd[k]).numpy() for k in ("p", "U", "density")}
        if rank == 0:
            np.savez(result_path, **got)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("per_x", [False, True])
def test_slab_periodic_seam_over_gloo(tmp_path, per_x):
    """The periodic-y seam of the ScaleNet step (*_saved.py:123-132, 228-237) across two slabs over gloo: row H-1 of
    Ux lives on the last rank, row 1 on the first; the line travels as int32 bits through an all-reduce.  A stand-in
    pressure model (OracleOps.cnn) keeps the step's structure; the decomposed step equals the single-domain one (to
    the rounding of the all-reduced std), and the seam line itself is checked against a run without it."""
    import types
    from fluidnet_cxx_b200.lib import distributed as D
    world, H, W, ghost, seed = 2, 128, 40, 64, 5
    netconf = {"normalizeInputThreshold": 1e-5, "periodic-y": True, "periodic-x": per_x}
    ops = OracleOps()

    def single(conf):
        bd = {k: torch.from_numpy(v.copy()) for k, v in global_state(H, W, seed).items()}
        rho, U, _ = ops.advect_forces_div(MCONF, MCONF["dt"], bd, False, False, (0, H))
        if conf["periodic-x"]:
            U[:, 1, :, :, 1] = U[:, 1, :, :, W - 1]
        if conf["periodic-y"]:
            U[:, 0, :, 1] = U[:, 0, :, H - 1].clone()
        one = types.SimpleNamespace(world=1, owned=lambda t: t)
        scale = D._std_finish(one, D._std_partial(one, U), U.numel(), conf["normalizeInputThreshold"])
        p, Uo, Ut = ops.cnn(None, U, bd["flags"], scale, prewall=True)
        if conf["periodic-x"]:
            Uo[:, 1, :, :, 1] = Ut[:, 1, :, :, W - 1]
        if conf["periodic-y"]:
            Uo[:, 0, :, 1] = Ut[:, 0, :, H - 1]
        ops.set_const(Uo, bd["UBCInvMask"], bd["UBC"])
        ops.set_const(rho, bd["densityBCInvMask"], bd["densityBC"])
        return {"p": p.numpy(), "U": Uo.numpy(), "density": rho.numpy()}

    want = single(netconf)
    noseam = single(dict(netconf, **{"periodic-y": False}))
    # (the inflow mask of this state pins rows 0..3 on a band of columns only: the seam shows on the rest of row 1)
    assert np.abs(want["U"][:, 0, :, 1] - noseam["U"][:, 0, :, 1]).max() > 1e-3
    out = str(tmp_path / "seam.npz")
    port = 33500 + (os.getpid() % 2000) + int(per_x)
    mp.spawn(_seam_worker, args=(world, port, H, W, ghost, seed, MCONF, netconf, out), nprocs=world, join=True)
    z = np.load(out)
    assert np.array_equal(z["density"], want["density"])
    for k in ("p", "U"):
        err = np.abs(z[k] - want[k]).max() / np.abs(want[k]).max()
        assert err < 1e-6, (k, err)
    assert np.abs(z["U"][:, 0, :, 1] - want["U"][:, 0, :, 1]).max() < 1e-6


# ---------------------------------------------------------------------------------------------
# The same decomposition with the ranks as THREADS of this process (barrier-based transport on CPU tensors, the
# oracle under a lock): no process spawn, so many geometries can be replayed -- seeded random ones below.
# ---------------------------------------------------------------------------------------------
class _ThreadComm:
    def __init__(self, world):
        import threading
        self.world, self.bar, self.box = world, threading.Barrier(world), {}

    def exchange_rows(self, dec, sends):
        for peer, buf in sends.items():
            self.box[(dec.rank, peer)] = buf.clone()
        self.bar.wait()
        out = {peer: self.box[(peer, dec.rank)].clone() for peer in sends}
        self.bar.wait()
        return out

    def all_reduce(self, dec, t, op):
        self.box[("r", dec.rank)] = t.clone()
        self.bar.wait()
        parts = torch.stack([self.box[("r", r)] for r in range(self.world)])
        t.copy_(parts.max(0).values if op == dist.ReduceOp.MAX else parts.sum(0))
        self.bar.wait()

    def all_gather(self, dec, mine):
        self.box[("g", dec.rank)] = mine
        self.bar.wait()
        parts = [self.box[("g", r)].clone() for r in range(self.world)]
        self.bar.wait()
        return parts


def run_threads(world, H, W, ghost, mconf, steps, seed, method="jacobi", net=None):
    """`steps` decomposed steps with `world` thread-ranks; returns the gathered fields of every step (+ the executed
    Jacobi iterations when the solve is residual-terminated)"""
    import threading
    from fluidnet_cxx_b200.lib.distributed import SlabDecomposition, simulate_distributed
    comm, lock = _ThreadComm(world), threading.Lock()
    base = OracleOps()

    class Locked:       # the C oracle keeps global counters: one call at a time
        def __getattr__(self, name):
            fn = getattr(base, name)

            def call(*a, **kw):
                with lock:
                    return fn(*a, **kw)
            return call
    results, errors = [None] * world, []

    def work(rank):
        try:
            dec = SlabDecomposition(H, ghost, rank=rank, world=world, comm=comm)
            bd = {k: dec.scatter(torch.from_numpy(v)) for k, v in global_state(H, W, seed).items()}
            outs = []
            for _ in range(steps):
                simulate_distributed(mconf, bd, net, method, dec, ops=Locked())
                rec = {k: dec.gather(bd[k]).numpy() for k in ("p", "U", "density")}
                if "jacobi_iterations" in bd:
                    rec["iters"] = int(bd["jacobi_iterations"])
                outs.append(rec)
            results[rank] = outs
        except Exception as e:      # noqa: BLE001 - surfaced in the main thread
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


def _random_decompositions(n, seed):
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        world = int(rng.randint(2, 6))
        ghost = int(rng.choice([16, 20, 24, 32]))                 # RA = 12: 3 / 7 / 8 / 16 iterations per exchange
        Hs = ghost + 4 * int(rng.randint(0, 4))
        iters = int(rng.choice([1, 3, 7, 8, 9, 16, 21, 30]))
        out.append((world, world * Hs, ghost, iters))
    return out


@pytest.mark.parametrize("world,H,ghost,iters", _random_decompositions(12, seed=4711))
def test_slab_decomposition_random_geometries(world, H, ghost, iters):
    """fixed-count Jacobi step, seeded random geometries (slab height from the ghost width upwards, iteration counts
    around the exchange interval): decomposed == single-domain, bit for bit, two steps"""
    W, steps, seed = 36, 2, 3
    mconf = dict(MCONF, jacobiIter=iters)
    ops = OracleOps()
    bd = {k: torch.from_numpy(v.copy()) for k, v in global_state(H, W, seed).items()}
    got = run_threads(world, H, W, ghost, mconf, steps, seed)
    for i in range(steps):
        rho, U, div = ops.advect_forces_div(mconf, mconf["dt"], bd, True, True, (0, H))
        p, _ = ops.o.solveLinearSystemJacobi(bd["flags"].numpy(), div.numpy(), False, 0.0, iters)
        U = ops.project(torch.from_numpy(p), U, bd, (0, H))
        bd["U"], bd["density"], bd["p"] = U, rho, torch.from_numpy(p)
        for k in ("p", "U", "density"):
            want = bd[k].numpy()
            bad = int(np.sum(~((got[i][k] == want) | (np.isnan(got[i][k]) & np.isnan(want)))))
            assert bad == 0, f"step {i} field {k}: {bad} cells differ"


@pytest.mark.parametrize("world,H,ghost,frac", [(3, 96, 20, 0.35), (5, 160, 24, 0.6), (2, 48, 16, 0.2), (4, 144, 32, 0.8)])
def test_slab_residual_terminated_random_stops(world, H, ghost, frac):
    """residual-terminated solve: the tolerance is placed between the residuals of two consecutive iterations chosen
    by `frac` of a 40-iteration run, so that the stop falls at various positions of a chunk"""
    W, seed, cap = 36, 9, 40
    ops = OracleOps()
    bd = {k: torch.from_numpy(v.copy()) for k, v in global_state(H, W, seed).items()}
    rho, U, div = ops.advect_forces_div(MCONF, MCONF["dt"], bd, True, True, (0, H))
    f, dv = bd["flags"].numpy(), div.numpy()
    res, p = [], None
    for it in range(cap):
        new = jacobi_numpy(f, dv, p, 1)
        res.append(float(np.sqrt(((new - (p if p is not None else 0)).astype(np.float64) ** 2).sum())))
        p = new
    stop = max(2, int(frac * cap))                  # 1-based iteration the solver should stop at
    assert res[stop - 1] < res[stop - 2]
    p_tol = 0.5 * (res[stop - 1] + res[stop - 2])
    mconf = dict(MCONF, pTol=p_tol, jacobiIter=cap)
    got = run_threads(world, H, W, ghost, mconf, 1, seed)[0]
    assert got["iters"] == stop
    assert np.array_equal(got["p"], jacobi_numpy(f, dv, None, stop))


@pytest.mark.parametrize("world,per_x", [(3, False), (4, True), (5, False)])
def test_slab_periodic_seam_thread_ranks(world, per_x):
    """the cross-slab periodic seam (see test_slab_periodic_seam_over_gloo) with 3-5 thread-ranks: the source row
    H-1 and the target row 1 are several ranks apart, and interior ranks neither send nor receive it"""
    import types
    from fluidnet_cxx_b200.lib import distributed as D
    H, W, ghost, seed = world * 64, 40, 64, 5
    netconf = {"normalizeInputThreshold": 1e-5, "periodic-y": True, "periodic-x": per_x}
    ops = OracleOps()
    bd = {k: torch.from_numpy(v.copy()) for k, v in global_state(H, W, seed).items()}
    rho, U, _ = ops.advect_forces_div(MCONF, MCONF["dt"], bd, False, False, (0, H))
    if per_x:
        U[:, 1, :, :, 1] = U[:, 1, :, :, W - 1]
    U[:, 0, :, 1] = U[:, 0, :, H - 1].clone()
    one = types.SimpleNamespace(world=1, owned=lambda t: t)
    scale = D._std_finish(one, D._std_partial(one, U), U.numel(), netconf["normalizeInputThreshold"])
    p, Uo, Ut = ops.cnn(None, U, bd["flags"], scale, prewall=True)
    if per_x:
        Uo[:, 1, :, :, 1] = Ut[:, 1, :, :, W - 1]
    Uo[:, 0, :, 1] = Ut[:, 0, :, H - 1]
    ops.set_const(Uo, bd["UBCInvMask"], bd["UBC"])
    ops.set_const(rho, bd["densityBCInvMask"], bd["densityBC"])
    want = {"p": p.numpy(), "U": Uo.numpy(), "density": rho.numpy()}
    got = run_threads(world, H, W, ghost, MCONF, 1, seed, method="convnet", net=types.SimpleNamespace(mconf=netconf))[0]
    assert np.array_equal(got["density"], want["density"])
    for k in ("p", "U"):
        err = np.abs(got[k] - want[k]).max() / np.abs(want[k]).max()
        assert err < 1e-6, (k, err)
    assert np.abs(got["U"][:, 0, :, 1] - want["U"][:, 0, :, 1]).max() < 1e-6


def test_decomposition_geometry():
    from fluidnet_cxx_b200.lib.distributed import SlabDecomposition, jacobi_chunk, RA
    d = SlabDecomposition(256, 48, rank=1, world=4)
    assert (d.lo, d.hi, d.g_top, d.g_bot, d.local_rows, d.rows()) == (64, 128, 48, 48, 160, (16, 176))
    d0 = SlabDecomposition(256, 48, rank=0, world=4)
    assert (d0.g_top, d0.g_bot, d0.local_rows, d0.rows()) == (0, 48, 112, (0, 112))
    assert SlabDecomposition(100, 48, rank=0, world=1).local_rows == 100      # no ghosts on one rank
    with pytest.raises(ValueError):
        SlabDecomposition(130, 48, rank=0, world=4)
    with pytest.raises(ValueError):
        SlabDecomposition(256, 96, rank=0, world=4)                            # ghost wider than a slab
    assert jacobi_chunk(48) == 32 and jacobi_chunk(RA + 4) == 3
    x = torch.arange(256.).view(1, 1, 1, 256, 1).expand(1, 2, 1, 256, 3)
    loc = d.scatter(x)
    assert loc.shape[3] == 256 and float(d.window(loc)[0, 0, 0, 0, 0]) == 16.0 and float(d.owned(loc)[0, 0, 0, 0, 0]) == 64.0
    assert d.window(loc).shape[3] == 160


def test_pack_transfer_unpack_roundtrip_in_process():
    """SlabDecomposition.pack / transfer / unpack with an in-process transport (three virtual ranks, no
    process group): after an exchange every ghost row holds its owner's value, owned rows are untouched,
    and the packed message is one contiguous buffer per neighbour and direction."""
    from fluidnet_cxx_b200.lib.distributed import SlabDecomposition

    class Mailbox:
        def __init__(self):
            self.box = {}

        def exchange_rows(self, dec, sends):          # called rank by rank: two passes (post, then collect)
            for peer, buf in sends.items():
                self.box[(dec.rank, peer)] = buf.clone()
            return {peer: self.box.get((peer, dec.rank), torch.zeros_like(buf)) for peer, buf in sends.items()}

    H, W, g, world = 48, 5, 4, 3
    comm = Mailbox()
    decs = [SlabDecomposition(H, g, rank=r, world=world, comm=comm) for r in range(world)]
    full = torch.arange(H, dtype=torch.float32).view(1, 1, 1, H, 1).expand(1, 2, 1, H, W).contiguous()
    locs = []
    for d in decs:
        t = torch.full_like(full, -1.0)
        t[:, :, :, d.lo:d.hi] = full[:, :, :, d.lo:d.hi] + 1000 * d.rank       # owned rows tagged with the owner
        locs.append(t)
    bufs = [{} for _ in decs]
    for _ in range(2):                                   # pass 1 posts every message, pass 2 delivers them
        for d, t, b in zip(decs, locs, bufs):
            d.pack([t], b)
            d.transfer(b)
            d.unpack([t], b)
    for d, t in zip(decs, locs):
        rows = t[0, 0, 0, :, 0]
        assert torch.equal(rows[d.lo:d.hi], full[0, 0, 0, d.lo:d.hi, 0] + 1000 * d.rank)
        if d.rank > 0:
            assert torch.equal(rows[d.lo - g:d.lo], full[0, 0, 0, d.lo - g:d.lo, 0] + 1000 * (d.rank - 1))
        if d.rank < world - 1:
            assert torch.equal(rows[d.hi:d.hi + g], full[0, 0, 0, d.hi:d.hi + g, 0] + 1000 * (d.rank + 1))
        outside = torch.cat([rows[:max(d.r0, 0)], rows[d.r1:]])
        assert bool((outside == -1).all())               # nothing outside the window is ever written
        n_neigh = (d.rank > 0) + (d.rank < world - 1)
        assert sum(1 for k in bufs[d.rank] if k[0] == "send") == n_neigh
        assert all(v.numel() == 2 * g * W for k, v in bufs[d.rank].items() if k[0] == "send")
