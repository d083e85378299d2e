"""The schedule of the window-based slab step (fluidnet_cxx_b200.lib.slab.schedule: ghost widths, row
windows of every stage, which rows every exchange moves) replayed on CPU with the C oracle as compute.

Every virtual rank works on arrays whose rows OUTSIDE its held rows are poisoned (the real rank does not
even allocate them) and only the rows the schedule says a stage writes are taken from that stage's
result.  If a ghost width or a window were one row too small, poison would reach an owned row: the
gathered owned rows must equal the single-domain oracle step BIT FOR BIT."""
import numpy as np
import pytest
import torch

from test_distributed_cpu import MCONF, OracleOps, global_state, jacobi_numpy


def single_domain(st, steps, iters=None):
    """the single-domain oracle step (simulate.py:28-171, jacobi branch) from state `st`"""
    iters = MCONF["jacobiIter"] if iters is None else iters
    ops = OracleOps()
    H = st["flags"].shape[3]
    bd = {k: torch.from_numpy(v.copy()) for k, v in st.items()}
    out = []
    for _ in range(steps):
        rho, U, div = ops.advect_forces_div(MCONF, MCONF["dt"], bd, True, True, (0, H))
        p, _ = ops.o.solveLinearSystemJacobi(bd["flags"].numpy(), div.numpy(), False, 0.0, iters)
        U = ops.project(torch.from_numpy(p), U, bd, (0, H))
        bd["U"], bd["density"], bd["p"] = U, rho, torch.from_numpy(p)
        out.append({k: bd[k].numpy().copy() for k in ("p", "U", "density")})
    return out


def poison(shape, rng, scale):
    return (rng.standard_normal(shape) * scale + 3.0 * scale).astype(np.float32)


# fast = True: the velocity is scaled so that max |u| dt = 0.95 cells, the limit the ghost widths are sized
# for (one step only: the projected velocity of the next step is not bounded by construction).
# order: how the virtual ranks are interleaved.  On the GPUs nothing keeps two ranks in lockstep between two
# exchanges: a rank pushes its rows into the neighbour as soon as IT is ready, and only its own progress
# past an exchange depends on the neighbours.  "lockstep" runs every operation on all ranks in turn;
# "ahead" / "behind" let the lowest / highest runnable rank run as far as its exchanges allow before anyone
# else moves; "random" picks the next rank at random.  A push that lands in rows the receiver still reads
# (a write-after-read hazard of the schedule) corrupts owned rows under the skewed orders.
# iters: 20 (three 8-iteration launches: one chunk at K = 3) unless given; the longer solves run SEVERAL chunks per
# step -- 52 iterations = 7 launches = chunks of 3, 3, 1 at K = 3 (the production shape: 100 iterations = 13 launches
# = 5 chunks), up to the 8 ranks of a full node.
@pytest.mark.parametrize("world,H,K,fast,order,iters", [
    (2, 64, 1, False, "lockstep", 20), (3, 96, 1, True, "ahead", 20), (4, 128, 1, False, "random", 20),
    (2, 96, 2, True, "behind", 20), (3, 144, 2, False, "ahead", 20), (2, 64, 1, True, "behind", 20),
    (3, 144, 3, False, "ahead", 20), (4, 160, 3, False, "behind", 20), (3, 120, 3, True, "random", 20),
    (4, 128, 2, False, "random", 20),
    (4, 192, 3, False, "random", 52), (8, 384, 3, False, "behind", 52), (8, 384, 2, False, "ahead", 44),
    (2, 96, 3, False, "ahead", 100), (8, 384, 3, True, "random", 100)])
def test_slab_schedule_bit_exact(world, H, K, fast, order, iters):
    from fluidnet_cxx_b200.lib import slab
    W, steps, seed = 40, (1 if fast else 2), 11
    ops = OracleOps()
    st = global_state(H, W, seed)
    if fast:
        st["U"] = (st["U"] * np.float32(0.95 / (np.abs(st["U"]).max() * MCONF["dt"]))).astype(np.float32)
    ref = single_domain(st, steps, iters)
    rng = np.random.RandomState(99)
    geo = [slab.geometry(H, world, r, K) for r in range(world)]
    scheds = [slab.schedule(H, world, r, iters, K)[1] for r in range(world)]
    assert len({len(s) for s in scheds}) == 1 and all([o[0] for o in s] == [o[0] for o in scheds[0]] for s in scheds)
    nops = len(scheds[0])

    def held_only(name, g):
        full = st[name]
        out = poison(full.shape, rng, 0.5 if name == "U" else 1.0)
        out[:, :, :, g["ya0"]:g["ya1"]] = full[:, :, :, g["ya0"]:g["ya1"]]
        return out
    R = []
    for g in geo:
        R.append({"U": [held_only("U", g), poison(st["U"].shape, rng, 0.5)],
                  "rho": [held_only("density", g), poison(st["density"].shape, rng, 1.0)],
                  "P": [poison(st["p"].shape, rng, 1.0) for _ in range(4)],
                  "div": poison(st["p"].shape, rng, 1.0),
                  "bd": {k: torch.from_numpy(st[k].copy()) for k in ("flags", "UBC", "UBCInvMask", "densityBC", "densityBCInvMask")}})
        # a rank holds its rows of the static fields only: outside, walls (harmless, never owned, never read by the kernels)
        f = R[-1]["bd"]["flags"].numpy()
        f[:, :, :, :g["ya0"]] = 2.0
        f[:, :, :, g["ya1"]:] = 2.0

    def run_op(r, step, i):
        """operation i of time step `step` on rank r (an exchange: the PUSH half only)"""
        g, op, me, par = geo[r], scheds[r][i], R[r], step % 2
        kind = op[0]
        if kind == "X":
            _, what, rows = op
            for q, first in ((r - 1, g["lo"]), (r + 1, g["hi"] - rows)):
                if q < 0 or q >= world:
                    continue
                sl = slice(first, first + rows)
                assert geo[q]["ya0"] <= sl.start and sl.stop <= geo[q]["ya1"], "push outside the neighbour's held rows"
                if what == "state":
                    R[q]["U"][par][:, :, :, sl] = me["U"][par][:, :, :, sl]
                    R[q]["rho"][par][:, :, :, sl] = me["rho"][par][:, :, :, sl]
                else:
                    R[q]["P"][what][:, :, :, sl] = me["P"][what][:, :, :, sl]
        elif kind == "advect":
            w0, w1 = op[1], op[2]
            bd = dict(me["bd"])
            bd["U"], bd["density"] = torch.from_numpy(me["U"][par]), torch.from_numpy(me["rho"][par])
            rho, U, div = ops.advect_forces_div(MCONF, MCONF["dt"], bd, True, True, (0, H))
            me["rho"][1 - par][:, :, :, w0:w1] = rho.numpy()[:, :, :, w0:w1]
            me["U"][1 - par][:, :, :, w0:w1] = U.numpy()[:, :, :, w0:w1]
            me["div"][:, :, :, w0:w1] = div.numpy()[:, :, :, w0:w1]
        elif kind == "jacobi":
            _, src, dst, it, r0, r1 = op
            p = jacobi_numpy(me["bd"]["flags"].numpy(), me["div"], None if src is None else me["P"][src], it)
            me["P"][dst][:, :, :, r0:r1] = p[:, :, :, r0:r1]
        else:
            _, pbuf, lo, hi = op
            U = ops.project(torch.from_numpy(me["P"][pbuf]), torch.from_numpy(me["U"][1 - par]), me["bd"], (0, H))
            me["U"][1 - par][:, :, :, lo:hi] = U.numpy()[:, :, :, lo:hi]
            me["p_final"] = pbuf

    # discrete-event replay: pc[r] = next operation (step * nops + i); pushed[r] = exchanges whose push rank r has done
    total = steps * nops
    pc = [0] * world
    pushed = [set() for _ in range(world)]
    snapshots = {}
    pick = np.random.RandomState(5)

    def runnable(r):
        if pc[r] >= total:
            return False
        step, i = divmod(pc[r], nops)
        if scheds[r][i][0] == "X" and pc[r] in pushed[r]:
            # waiting half of the exchange: every neighbour must have pushed at this site
            return all(pc[r] in pushed[q] for q in (r - 1, r + 1) if 0 <= q < world)
        return True

    def advance(r):
        step, i = divmod(pc[r], nops)
        if scheds[r][i][0] == "X":
            if pc[r] not in pushed[r]:
                run_op(r, step, i)
                pushed[r].add(pc[r])
                return                      # the wait half is a separate event
        else:
            run_op(r, step, i)
        pc[r] += 1
        if pc[r] % nops == 0:               # end of a time step on this rank: keep its owned rows for the comparison
            par = (step + 1) % 2
            me, g = R[r], geo[r]
            snapshots[(step, r)] = {"U": me["U"][par][:, :, :, g["lo"]:g["hi"]].copy(),
                                    "density": me["rho"][par][:, :, :, g["lo"]:g["hi"]].copy(),
                                    "p": me["P"][me["p_final"]][:, :, :, g["lo"]:g["hi"]].copy()}

    while any(p < total for p in pc):
        ready = [r for r in range(world) if runnable(r)]
        assert ready, "deadlock in the schedule"
        if order == "lockstep":
            r = min(ready, key=lambda q: pc[q])
            advance(r)
        elif order == "random":
            advance(int(pick.choice(ready)))
        else:
            r = ready[0] if order == "ahead" else ready[-1]
            while runnable(r):
                advance(r)
    for step in range(steps):
        for k in ("p", "U", "density"):
            got = np.concatenate([snapshots[(step, r)][k] for r in range(world)], axis=3)
            want = ref[step][k]
            bad = int(np.sum(~((got == want) | (np.isnan(got) & np.isnan(want)))))
            assert bad == 0, f"world {world} K {K} {order} step {step} field {k}: {bad} owned cells differ"


def _random_slab_configs(n, seed):
    """seeded random (world, H, K, fast, order, iters): slab heights from the minimum the ghost width allows (Hs = G)
    upwards, iteration counts that end inside / on / right after an 8-iteration launch and a K-launch chunk"""
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        K = int(rng.randint(1, 5))
        G = 8 * K + 8
        world = int(rng.randint(2, 7))
        Hs = G + int(rng.choice([0, 1, 3, 8, 17]))
        iters = int(rng.choice([1, 5, 8, 9, 16, 8 * K, 8 * K + 1, 8 * K + 7, 16 * K, 16 * K + 3, 40]))
        out.append((world, world * Hs, K, bool(rng.randint(0, 2)), str(rng.choice(["lockstep", "ahead", "behind", "random"])),
                    iters))
    return out


@pytest.mark.parametrize("world,H,K,fast,order,iters", _random_slab_configs(24, seed=2024))
def test_slab_schedule_random_configs(world, H, K, fast, order, iters):
    """the same replay on seeded random geometries: minimal slab heights (Hs = G), K up to 4, 2-6 ranks, iteration
    counts around the launch / chunk boundaries"""
    test_slab_schedule_bit_exact(world, H, K, fast, order, iters)


def test_slab_geometry():
    from fluidnet_cxx_b200.lib import slab
    g = slab.geometry(4096, 8, 3)
    assert (g["lo"], g["hi"], g["G"], g["dg"], g["ya0"], g["ya1"], g["rows_held"]) == (1536, 2048, 16, 8, 1520, 2064, 544)
    assert slab.geometry(4096, 8, 0)["ya0"] == 0 and slab.geometry(4096, 8, 7)["ya1"] == 4096
    assert slab.geometry(100, 1, 0)["rows_held"] == 100
    with pytest.raises(ValueError):
        slab.geometry(100, 3, 0)
    with pytest.raises(ValueError):
        slab.geometry(64, 8, 0)           # 8 rows per slab < ghost width
    # every rank's schedule has the same shape (the exchanges pair up)
    shapes = {tuple(o[0] for o in slab.schedule(1024, 4, r, 100)[1]) for r in range(4)}
    assert len(shapes) == 1
    _, ops = slab.schedule(1024, 4, 1, 100)
    # K = 1: the state exchange + one pressure exchange between consecutive launches (none after the last:
    # every launch also computes the row below the slab that the pressure gradient reads)
    assert sum(1 for o in ops if o[0] == "X") == 1 + 12 and sum(o[3] for o in ops if o[0] == "jacobi") == 100
    _, ops3 = slab.schedule(1024, 4, 1, 100, 3)
    assert sum(1 for o in ops3 if o[0] == "X") == 1 + 4
    _, one = slab.schedule(1024, 4, 1, 20, 3)           # a single chunk keeps one hand-shake after the stencil stage
    assert [o[0] for o in one].count("X") == 2


def test_bench_held_state_equals_global_init():
    """bench.plume_held_state builds a rank's rows of the initial state directly; they must be the rows of the
    state bench.init_state builds on the whole grid (here with the reference's own emptyDomain / createPlumeBCs)."""
    import bench
    import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref not built")
    reflib = ref_loader.load()
    H, W = 96, 72
    wl = dict(bench.WORKLOADS["plume4096_jacobi100"], res=(1, H, W))
    mconf = bench.workload_mconf(wl)
    U_np, rho_np = bench.synthetic_state_numpy(1, H, W, 0)
    bd = {"p": torch.zeros(1, 1, 1, H, W), "U": torch.zeros(1, 2, 1, H, W), "flags": torch.zeros(1, 1, 1, H, W),
          "density": torch.zeros(1, 1, 1, H, W)}
    bench.init_state(reflib.fluid, wl, mconf, bd, U_np, rho_np, torch.from_numpy)
    held = bench.plume_held_state(wl, mconf, H, W)
    for r0, r1 in ((0, 40), (2, 50), (30, 96), (0, 96), (5, 9)):
        for k in ("U", "density", "flags", "UBC", "UBCInvMask", "densityBC", "densityBCInvMask"):
            assert torch.equal(held(k, r0, r1), bd[k][:, :, :, r0:r1]), (k, r0, r1)
