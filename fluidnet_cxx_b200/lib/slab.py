"""Slab-decomposed 2-D Jacobi step on PER-RANK ROW WINDOWS with halo exchange over NVLink peer memory
(the reference has no multi-device path).

Rank r of N owns rows [lo, hi) of the global H x W grid and HOLDS only rows [lo - G, hi + G) (clipped to
the grid): device memory per rank is proportional to 1/N.  The kernels index with global coordinates
through held-row arrays (fnx_step_params.held_row_*, fnx_jacobi_iterate_held, fnx_step_project_bcs_held),
so every owned row is bit-identical to the single-GPU step (same fp32 binade for every back-traced
position, same operation order).

Halos travel by ONE kernel per exchange (csrc/halo.cu, fnx_halo_exchange): boundary rows are stored
straight into the neighbours' ghost rows through the peer mapping of a symmetric-memory arena, then a
per-site flag is raised in the neighbour and awaited from it.  No pack / unpack, no NCCL call, no host
synchronisation inside a step, so a whole time step -- stencil kernels and exchanges -- is ONE CUDA
graph (two are captured, for the two parities of the state ping-pong).  Waits are bounded on the device.

Schedule of a step (K = Jacobi launches per exchange, 8 iterations per launch; dg = 8 K; G = dg + 8):
  X(density, U)                     G ghost rows of the state: valid inputs on [lo-G, hi+G)
  advect + forces + divergence      on [lo-dg-1, hi+dg+1)  (MacCormack reads rows j-2 .. j+4 for |u| dt <= 1)
                                    -> divergence valid on [lo-dg, hi+dg)
  Jacobi chunk 0                    K launches from p = 0, writing [lo-8(K-1-l), hi+8(K-1-l)) in launch l
  repeat: X(p) dg+1 ghost rows, Jacobi chunk (K launches, the last one writes rows [lo-1, hi))
                                    (every launch also computes row lo-1: the pressure gradient of the owned rows reads
                                     it, and one redundant row is cheaper than one more exchange per step)
  project + BCs                     owned rows
A rank only pushes into ghost rows of the buffer its neighbour will read AFTER the neighbour's matching
wait, and never touches a ghost row of a buffer between its own signal and the next push into it: every
write-after-read pair is ordered by the flag chain (halo.cu).
"""
import ctypes

import torch
import torch.distributed as dist

from .. import _native as N

JACOBI_LAUNCH_ITERS = 8


def geometry(H, world, rank, K=1):
    """Row bookkeeping of rank `rank`: owned [lo, hi), held [ya0, ya1), ghost widths."""
    if H % world:
        raise ValueError(f"{H} rows do not split evenly over {world} ranks")
    Hs = H // world
    dg = JACOBI_LAUNCH_ITERS * K
    G = dg + 8          # state ghost rows: divergence on [lo-dg-1, hi+dg) + MacCormack reach (2 rows for |u| dt < 1) + slack
    if world > 1 and Hs < G:
        raise ValueError(f"slab height {Hs} is smaller than the ghost width {G}")
    lo, hi = rank * Hs, (rank + 1) * Hs
    g = dict(H=H, Hs=Hs, lo=lo, hi=hi, dg=dg, G=G, K=K, rank=rank, world=world,
             ya0=max(0, lo - G) if world > 1 else 0, ya1=min(H, hi + G) if world > 1 else H)
    g["rows_held"] = g["ya1"] - g["ya0"]
    return g


def schedule(H, world, rank, iters, K=1):
    """The step of one rank as a list of operations on row intervals (pure bookkeeping, shared by the
    CUDA stepper below and by the CPU test that replays it with the oracle on NaN-poisoned arrays):
      ("X", what, rows)            exchange `rows` boundary rows; what = "state" or a pressure buffer index (0..3)
      ("advect", w0, w1)           advect + forces + divergence on rows [w0, w1)
      ("jacobi", src, dst, it, r0, r1)   `it` iterations, pressure buffer src (None = from zero) -> dst, rows written
      ("project", p, lo, hi)       velocity update + BCs on the owned rows, pressure buffer p
    """
    g = geometry(H, world, rank, K)
    multi = world > 1
    lo, hi, dg = g["lo"], g["hi"], g["dg"]
    ops = []
    if multi:
        ops.append(("X", "state", g["G"]))
    ops.append(("advect", max(0, lo - dg - 2), min(H, hi + dg + 1)) if multi else ("advect", 0, H))
    # Pressure buffers: 0 / 1 receive the result of a chunk (the only buffers ever pushed into, alternating
    # from chunk to chunk), 2 / 3 hold the intermediate launches of a chunk (their ghost rows are computed
    # locally).  A neighbour may still be inside chunk c -- reading the ghost rows of buffer (c-1) % 2 and
    # of the intermediates -- when this rank pushes the result of chunk c: it goes into buffer c % 2, which
    # nobody reads before the wait of that exchange.
    # The pressure gradient of the owned rows reads p one row BELOW the slab: every Jacobi launch also computes
    # row lo-1 (one redundant row instead of one more exchange per step), so the exchanges carry dg + 1 rows.
    lo_p = max(0, lo - 1) if multi else 0
    n_launch = (iters + JACOBI_LAUNCH_ITERS - 1) // JACOBI_LAUNCH_ITERS
    src = None
    for l in range(n_launch):
        it = min(JACOBI_LAUNCH_ITERS, iters - l * JACOBI_LAUNCH_ITERS)
        chunk, in_chunk = divmod(l, K)
        n_in_chunk = min(K, n_launch - chunk * K)
        if multi and l > 0 and in_chunk == 0:
            ops.append(("X", src, dg + 1))
        margin = JACOBI_LAUNCH_ITERS * (n_in_chunk - 1 - in_chunk) if multi else 0
        r0, r1 = (max(0, lo_p - margin), min(H, hi + margin)) if multi else (0, H)
        dst = chunk % 2 if in_chunk == n_in_chunk - 1 else 2 + in_chunk % 2
        ops.append(("jacobi", src, dst, it, r0, r1))
        src = dst
    if multi and n_launch <= K:
        # a step without any pressure exchange (a single chunk) still needs ONE hand-shake after the stencil
        # stage: the neighbours' next state push lands in ghost rows of the new state that this rank's
        # advect / forces kernels have just written, and only a flag received after those kernels orders the two
        ops.append(("X", src, 4))
    ops.append(("project", src, lo, hi))
    return g, ops


class ProcessTopology:
    """One rank per process: the arena is torch symmetric memory, peers are mapped over NVLink."""

    def __init__(self, device, group=None):
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.device = device
        self.stream = None        # the caller's current stream

    def arena(self, nbytes):
        import torch.distributed._symmetric_memory as symm
        buf = symm.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        hdl = symm.rendezvous(buf, self.group)
        buf.zero_()
        peers = {r: (buf if r == self.rank else hdl.get_buffer(r, (int(nbytes),), torch.uint8)) for r in range(self.world)}
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)   # every arena is zeroed before anybody raises a flag in it
        self._keep = (hdl, peers)
        return buf, {r: t.data_ptr() for r, t in peers.items()}

    def barrier(self):
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)


class VirtualWorld:
    """N virtual ranks on ONE GPU (tests): every rank has its own arena and stream in this process; a
    peer pointer is simply the other rank's device pointer."""

    def __init__(self, world, device):
        self.world, self.device = world, device
        self.arenas = {}
        self.topos = [self._Topo(self, r) for r in range(world)]

    class _Topo:
        def __init__(self, vw, rank):
            self.vw, self.rank, self.world, self.device = vw, rank, vw.world, vw.device
            self.stream = torch.cuda.Stream(device=vw.device)
            self.n_arena = 0

        def arena(self, nbytes):
            key = self.n_arena
            self.n_arena += 1
            if key not in self.vw.arenas:
                self.vw.arenas[key] = [torch.zeros(int(nbytes), dtype=torch.uint8, device=self.device)
                                       for _ in range(self.world)]
            bufs = self.vw.arenas[key]
            return bufs[self.rank], {r: bufs[r].data_ptr() for r in range(self.world)}

        def barrier(self):
            torch.cuda.synchronize(self.device)


def _align(x, a=256):
    return (x + a - 1) // a * a


class _Arena:
    def __init__(self, topo, nbytes):
        self.buf, self.peer_base = topo.arena(nbytes)
        self.base = self.buf.data_ptr()
        self.off = 0

    def take(self, shape, dtype=torch.float32):
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        t = self.buf[self.off:self.off + nbytes].view(dtype).view(*shape)
        self.off = _align(self.off + nbytes)
        assert self.off <= self.buf.numel(), "arena too small"
        return t

    def offset_of(self, t):
        return t.data_ptr() - self.base


class SlabJacobiStep:
    """The slab-decomposed fused Jacobi step of one rank (see the module docstring).

        step = SlabJacobiStep(topo, mconf, H, W, held_state, K=1)
        step.step(); ...; step.owned('U')

    `held_state(name, ya0, ya1)` returns this rank's held rows of the GLOBAL field `name` in
    {'U', 'density', 'flags', 'UBC', 'UBCInvMask', 'densityBC', 'densityBCInvMask'} as a CPU or CUDA tensor
    (1, C, 1, ya1-ya0, W), or None for an absent mask -- the global grid never has to exist on a GPU.
    """

    def __init__(self, topo, mconf, H, W, held_state, K=1, use_graph=True, check_reach_every=64):
        # the ghost rows leave room for MAX_REACH cells of back-trace per step: verified every `check_reach_every`
        # steps by step() itself (a host read, amortised; 0 = the caller calls check_reach())
        self.check_reach_every, self._steps = int(check_reach_every), 0
        import importlib
        self.sim = importlib.import_module(__package__ + ".simulate")
        self.lib = N.load()
        self.topo, self.mconf, self.H, self.W = topo, mconf, int(H), int(W)
        self.g = g = geometry(H, topo.world, topo.rank, K)
        dev = topo.device
        if W % 4:
            raise ValueError("the slab path needs W to be a multiple of 4 (16-byte rows)")
        if mconf['viscosity'] != 0 or mconf.get('correctScalar', False) or mconf['pTol'] > 0:
            raise NotImplementedError("SlabJacobiStep covers the fused inviscid step with a fixed Jacobi count")
        if ('periodic-x' in mconf and mconf['periodic-x']) or ('periodic-y' in mconf and mconf['periodic-y']):
            raise NotImplementedError("periodic seams across slabs are not implemented")
        self.iters = int(mconf['jacobiIter'])
        if self.iters < 1:
            raise ValueError("At least 1 iteration of the solver is needed.")
        Rh, Rmax = g["rows_held"], g["Hs"] + 2 * g["G"]
        plane_max = Rmax * W * 4
        n_launch = (self.iters + JACOBI_LAUNCH_ITERS - 1) // JACOBI_LAUNCH_ITERS
        self.n_chunks = (n_launch + K - 1) // K
        n_sites = 2 * (2 + self.n_chunks)      # per parity: state + (n_chunks - 1) pressure exchanges
        # ---- symmetric arena: state ping-pong, pressure ping-pong, flags of the exchange sites.  Every rank
        # reserves room for the tallest slab (interior ranks hold Hs + 2G rows, edge ranks fewer) so that a
        # field starts at the same offset everywhere; a rank's tensor covers its own held rows only.
        self.arena = ar = _Arena(topo, 10 * _align(plane_max) + _align(n_sites * 2 * 4) + 4096)

        def field(channels):
            start = ar.off
            t = ar.take((1, channels, 1, Rh, W))
            ar.off = _align(start + channels * plane_max)
            return t
        self.U = [field(2), field(2)]
        self.rho = [field(1), field(1)]
        self.P = [field(1), field(1), field(1), field(1)]   # see schedule(): 0/1 exchanged, 2/3 intermediates
        self.site_flags = ar.take((n_sites, 2), torch.int32)
        # ---- local (never pushed into) ----
        z = lambda c: torch.zeros((1, c, 1, Rh, W), dtype=torch.float32, device=dev)  # noqa: E731
        self.flags, self.div = z(1), z(1)
        self.counters = torch.zeros((n_sites, 2), dtype=torch.int32, device=dev)   # [site] = (epoch, done)
        self.masks = {}
        ya0, ya1 = g["ya0"], g["ya1"]
        for name in ("UBC", "UBCInvMask", "densityBC", "densityBCInvMask"):
            t = held_state(name, ya0, ya1)
            self.masks[name] = None if t is None else t.to(dev, torch.float32).contiguous()
        self.flags.copy_(held_state("flags", ya0, ya1))
        self.U[0].copy_(held_state("U", ya0, ya1))
        self.rho[0].copy_(held_state("density", ya0, ya1))
        self.parity = 0
        self._bd = {k: v for k, v in self.masks.items() if v is not None}
        self.mrows = self.sim._mask_rows(self.lib, self._bd, self.flags, 0)
        nbytes = self.lib.fnx_step_workspace(1, 1, Rh, W, 0)
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self._descs, self._params = [], []      # ctypes objects the captured launches point at: kept alive here
        self._n_sites = 0
        self._mask_cache = {}
        self._plans = [self._build_plan(par) for par in (0, 1)]
        self._graphs = [None, None]
        self.use_graph = bool(use_graph) and isinstance(topo, ProcessTopology)
        topo.barrier()

    # ---- peer addressing ---------------------------------------------------------------------
    def _peer_row_ptr(self, q, t, channel, row):
        """device address, as mapped on THIS rank, of (channel, global row) of arena tensor `t` on rank q"""
        gq = geometry(self.H, self.topo.world, q, self.g["K"])
        off = self.arena.offset_of(t) + (channel * gq["rows_held"] + (row - gq["ya0"])) * self.W * 4
        return self.arena.peer_base[q] + off

    def _row_ptr(self, t, channel, row):
        return t.data_ptr() + (channel * self.g["rows_held"] + (row - self.g["ya0"])) * self.W * 4

    def _exchange_desc(self, fields, rows):
        """descriptor of one exchange site: `rows` boundary rows of every (tensor, channel) in `fields`
        to the lower / upper neighbour's ghost rows"""
        g, W = self.g, self.W
        site = self._n_sites
        self._n_sites += 1
        d = N.HaloDesc()
        d.n_fields = len(fields)
        k = 0
        for q, first_row in ((g["rank"] - 1, g["lo"]), (g["rank"] + 1, g["hi"] - rows)):
            if q < 0 or q >= g["world"]:
                continue
            for f, (t, ch) in enumerate(fields):
                d.src[k][f] = self._row_ptr(t, ch, first_row)
                d.dst[k][f] = self._peer_row_ptr(q, t, ch, first_row)
            d.count[k] = rows * W
            # my slot in q's flag pair: I am q's upper neighbour (slot 1) when q is below me, else slot 0
            slot_in_q = 1 if q < g["rank"] else 0
            my_slot = 0 if q < g["rank"] else 1
            d.flag_out[k] = self.arena.peer_base[q] + self.arena.offset_of(self.site_flags) + (site * 2 + slot_in_q) * 4
            d.flag_in[k] = self.site_flags.data_ptr() + (site * 2 + my_slot) * 4
            k += 1
        d.n_peers = k
        d.epoch = self.counters.data_ptr() + (site * 2) * 4
        d.done = self.counters.data_ptr() + (site * 2 + 1) * 4
        self._descs.append(d)
        return d

    # ---- the step as a list of phases (each enqueues this rank's work of that phase) ------------------
    def _build_plan(self, par):
        W, H, lib = self.W, self.H, self.lib
        g, ops = schedule(H, self.topo.world, self.topo.rank, self.iters, self.g["K"])
        multi = g["world"] > 1
        U_in, U_out, rho_in, rho_out = self.U[par], self.U[1 - par], self.rho[par], self.rho[1 - par]
        ya0, ya1 = (g["ya0"], g["ya1"]) if multi else (0, H)
        m = self.masks
        mrows = self.mrows.data_ptr() if self.mrows is not None else None
        st = lambda: N.stream_of(self.flags)  # noqa: E731
        phases, p_final = [], None
        for op in ops:
            if op[0] == "X":
                fields = [(rho_in, 0), (U_in, 0), (U_in, 1)] if op[1] == "state" else [(self.P[op[1]], 0)]
                d = self._exchange_desc(fields, op[2])
                phases.append(lambda d=d: N.check(lib.fnx_halo_exchange(ctypes.byref(d), st()), "slab exchange"))
            elif op[0] == "advect":
                prm = self.sim._step_params(self.mconf, float(self.mconf['dt']), 0)
                prm.apply_wall_bcs = 1
                prm.density_const_passes = 2
                prm.row_begin, prm.row_end = op[1], op[2]
                prm.held_row_begin, prm.held_row_end = (ya0, ya1) if multi else (0, 0)
                self._params.append(prm)
                phases.append(lambda prm=prm: N.check(lib.fnx_step_advect_forces_div(
                    ctypes.byref(prm), N.ptr(rho_in), N.ptr(U_in), N.ptr(self.flags), N.ptr(m["UBC"]),
                    N.ptr(m["UBCInvMask"]), N.ptr(m["densityBC"]), N.ptr(m["densityBCInvMask"]), mrows, N.ptr(rho_out),
                    N.ptr(U_out), N.ptr(self.div), 1, 1, H, W, 0, self.ws.data_ptr(), self.ws.numel(), st()), "slab step"))
            elif op[0] == "jacobi":
                _, src, dst, it, r0, r1 = op
                src_t, dst_t = (None if src is None else self.P[src]), self.P[dst]
                mk = self._tile_masks(r0, r1, ya0, ya1)
                phases.append(lambda src_t=src_t, dst_t=dst_t, it=it, r0=r0, r1=r1, mk=mk: N.check(
                    lib.fnx_jacobi_iterate_held_masked(N.ptr(self.flags), N.ptr(self.div), N.ptr(src_t), N.ptr(dst_t), 1, H, W,
                                                       it, r0, r1, ya0, ya1, mk.data_ptr(), mk.numel(), st()), "slab jacobi"))
            else:
                _, pbuf, lo, hi = op
                p_final = self.P[pbuf]
                phases.append(lambda p_final=p_final, lo=lo, hi=hi: N.check(lib.fnx_step_project_bcs_held(
                    N.ptr(p_final), N.ptr(U_out), N.ptr(self.flags), N.ptr(m["UBC"]), N.ptr(m["UBCInvMask"]), mrows, 1, 1, 1,
                    H, W, 0, lo, hi, ya0 if multi else 0, ya1 if multi else 0, st()), "slab project"))
        return dict(phases=phases, p=p_final)

    def _tile_masks(self, r0, r1, ya0, ya1):
        """per-thread Neumann / fixed-cell masks of the Jacobi launch that writes rows [r0, r1): the flags are static
        during a simulation, so they are decoded once per launch shape (refresh_flags() after changing them)"""
        key = (r0, r1, ya0, ya1)
        mk = self._mask_cache.get(key)
        if mk is None:
            nbytes = self.lib.fnx_jacobi_tilemask_bytes(1, r1 - r0, self.W)
            mk = torch.empty(nbytes, dtype=torch.uint8, device=self.flags.device)
            self._mask_cache[key] = mk
            N.check(self.lib.fnx_jacobi_tilemask_held(N.ptr(self.flags), 1, self.H, self.W, r0, r1, ya0, ya1, mk.data_ptr(),
                                                      mk.numel(), N.stream_of(self.flags)), "slab tile masks")
        return mk

    def refresh_flags(self):
        """call after writing new cell types into `self.flags` (obstacles moved): re-derives what depends on them"""
        for (r0, r1, ya0, ya1), mk in self._mask_cache.items():
            N.check(self.lib.fnx_jacobi_tilemask_held(N.ptr(self.flags), 1, self.H, self.W, r0, r1, ya0, ya1, mk.data_ptr(),
                                                      mk.numel(), N.stream_of(self.flags)), "slab tile masks")

    def phases(self):
        return self._plans[self.parity]["phases"]

    def advance(self):
        self.parity ^= 1

    def _run_direct(self):
        for ph in self.phases():
            ph()

    def step(self):
        """one time step of this rank (every rank of the group must call it): a replay of the captured
        graph of this parity, or direct launches before capture()"""
        if self.check_reach_every and self._steps % self.check_reach_every == 0 and self.g["world"] > 1:
            self.check_reach()
        self._steps += 1
        gph = self._graphs[self.parity]
        if gph is not None:
            gph.replay()
        else:
            self._run_direct()
        self.advance()

    def capture(self):
        """capture one CUDA graph per parity.  Call it on every rank, after at least one direct step of
        each parity (lazy initialisation must not happen under capture)."""
        if not self.use_graph:
            return False
        torch.cuda.synchronize(self.topo.device)
        for par in (0, 1):
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                for ph in self._plans[par]["phases"]:
                    ph()
            self._graphs[par] = gph
        torch.cuda.synchronize(self.topo.device)
        return True

    # ---- state access -------------------------------------------------------------------------
    def _own(self, t):
        g = self.g
        return t[:, :, :, g["lo"] - g["ya0"]:g["hi"] - g["ya0"], :]

    def owned(self, name):
        """owned rows of the CURRENT state ('U', 'density') or of the pressure of the last step ('p')"""
        if name == "p":
            return self._own(self._plans[self.parity ^ 1]["p"])
        return self._own((self.U if name == "U" else self.rho)[self.parity])

    def held(self, name):
        return (self.U if name == "U" else self.rho)[self.parity]

    MAX_REACH = 2.0     # cells per step the ghost width G = 8K + 8 leaves room for (the MacCormack stencil of a cell
                        # that moves d cells reads about 2 ceil(d) rows beyond the window; tests/test_slab_cpu.py)

    def max_reach(self):
        """max |U| dt over the owned rows of this rank, in cells per step (a host synchronisation)"""
        return float(self.owned("U").abs().max()) * abs(float(self.mconf["dt"]))

    def check_reach(self):
        """raise if the velocity outruns the ghost rows (call it every few steps, not every step: it synchronises)"""
        r = self.max_reach()
        if not r <= self.MAX_REACH:
            raise RuntimeError(f"slab step: max|U|*dt = {r:.3f} cells per step exceeds the {self.MAX_REACH} the ghost "
                               f"rows are sized for; use a smaller dt")
        return r


def run_virtual(steppers, steps):
    """drive the virtual ranks of one VirtualWorld through `steps` time steps: phase by phase, every rank's
    work of a phase is enqueued on that rank's stream before the next phase (the exchange kernels of all
    ranks must be in flight together)"""
    for _ in range(steps):
        plans = [s.phases() for s in steppers]
        for i in range(len(plans[0])):
            for s, ph in zip(steppers, plans):
                with torch.cuda.stream(s.topo.stream):
                    ph[i]()
        for s in steppers:
            s.advance()
    torch.cuda.synchronize()
