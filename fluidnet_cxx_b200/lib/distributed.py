"""Slab-decomposed fluid step: one process per GPU, halo exchange over torch.distributed
(NCCL over NVLink on the GPU box, gloo in the CPU tests) -- SURVEY.md section 8e.

The reference has no multi-device path; this is the B200 scaling layer around the same kernels.

Decomposition: 1-D slabs along the outermost spatial axis (H in 2-D, D in 3-D) so a halo is a
contiguous block of rows per channel.  Rank r owns rows [r*Hs, (r+1)*Hs) of the global grid and
computes the WINDOW [r*Hs - ghost, (r+1)*Hs + ghost).  Every rank keeps GLOBAL-SIZED arrays and
global coordinates (back-traced positions are rounded in the same fp32 binade as on one GPU; a
translated local frame would not be bit-compatible) and the kernels are launched on the window's
rows only (fnx_step_params.row_begin/row_end, fnx_jacobi_iterate, fnx_step_project_bcs_rows).
Rows outside the window are never written: a stencil that reaches them sees stale data, and that
error travels inward one row per stencil radius.  Every stage's dependency radius is known, so with
enough ghost rows the OWNED rows are exactly (bit for bit) what the single-GPU step computes; ghost
rows are refreshed from their owners before the error can reach owned rows:

  exchange(density, U)                                ghost rows valid
  advect + forces + BCs + divergence on the window    valid >= RA rows inside the window edge
  Jacobi: chunks of k <= ghost - RA - 1 iterations    each iteration costs one more row;
          exchange(p) between chunks                  after a chunk, valid >= RA + k - 1 < ghost rows in
  velocityUpdate + BCs                                radius 1
  (convnet: global std by all-reduce, then the CNN on a compact copy of the window -- convolutions
   and x2 / x4 resizes are translation invariant --; receptive field <= 54 rows)

Collectives: neighbour send/recv of halo rows (batched), one all-reduce of (sum, sum of squares)
for the ScaleNet normalisation, one all-reduce (max) of |U| for the advection reach check.
"""
import math

import torch
import torch.distributed as dist

# rows an interior edge can contaminate in advect(MacCormack, |u| dt <= 1) + forces + BCs + divergence
RA = 12
# full-resolution rows the CNN stage adds on top of RA (three-scale receptive field)
CNN_REACH = 54 - 9
GHOST_JACOBI = 48      # k = 32 iterations between pressure exchanges
GHOST_CONVNET = 64
# 3-D slabs along D with the slice-wise CNN projection (FluidNet.forward_fields_3d): the network itself does not couple
# planes; only the 3-D divergence (planes k, k+1) and the z pressure gradient (k, k-1) do, so the CNN stage needs the
# owned planes +- CNN3D_PLANES and the ghost width is RA + CNN3D_PLANES rounded up
CNN3D_PLANES = 2
GHOST_CONVNET_3D = 16


def jacobi_chunk(ghost):
    """Jacobi iterations between two pressure halo exchanges for a ghost width: after n iterations
    rows closer than RA + n - 1 to an interior edge are stale, and velocityUpdate reads one ghost
    row, so n <= ghost - RA - 1; rounded down to the blocked kernel's 8 iterations per launch."""
    k = ghost - RA - 1
    if k < 1:
        raise ValueError(f"ghost width {ghost} leaves no room for a Jacobi iteration (needs > {RA + 1})")
    return (k // 8) * 8 if k >= 8 else k


class PeerMemoryTransport:
    """Halo transfer over NVLink peer memory instead of NCCL send/recv.

    Every rank owns a symmetric "inbox" per exchange tag (torch symmetric memory: the allocation and
    the address exchange are torch.distributed plumbing); a sender copies its packed boundary rows
    STRAIGHT INTO THE NEIGHBOUR'S INBOX with an ordinary device copy through the peer mapping (stores
    over NVLink / NVSwitch), then raises a signal in the neighbour's signal pad; the receiver waits for
    the signal, unpacks, and raises an ack that lets the sender overwrite the inbox next time (credit
    flow control, pre-armed at creation).  Everything is stream-ordered device work: the whole
    exchange can live inside a CUDA graph, which removes the host round trips of the NCCL path (the
    dominant cost of a strong-scaled step).  `SlabDecomposition(..., transport="peer")` selects it.
    """

    def __init__(self, decomp):
        import torch.distributed._symmetric_memory as symm
        self.symm, self.decomp = symm, decomp
        self.group = decomp.group if decomp.group is not None else dist.group.WORLD
        self.regions = {}

    def _slot_in_peer(self, peer):          # which inbox slot of `peer` is mine
        return 1 if peer < self.decomp.rank else 0

    def _slot_of_peer(self, peer):          # which of my inbox slots `peer` writes
        return 0 if peer < self.decomp.rank else 1

    def region(self, tag, nelem, dtype, device):
        """symmetric inbox [2 senders][nelem] for `tag` (collective on first use / growth)"""
        r = self.regions.get(tag)
        if r is not None and r["nelem"] >= nelem:
            return r
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("peer transport: inbox allocation during graph capture (warm up first)")
        t = self.symm.empty((2, int(nelem)), dtype=dtype, device=device)
        hdl = self.symm.rendezvous(t, self.group)
        # channels come from a counter that only grows: a regrown tag must never land on the pair of another
        # tag (channel 0/1 are left to hdl.barrier)
        self._next_channel = getattr(self, "_next_channel", 2)
        base = self._next_channel
        self._next_channel += 2
        r = {"t": t, "hdl": hdl, "nelem": int(nelem), "data": base, "ack": base + 1}
        self.regions[tag] = r
        # credits: every neighbour may write my inbox once before my first ack
        for peer, _, _ in self.decomp._halo_slices():
            hdl.put_signal(peer, r["ack"])
        hdl.barrier(0)
        return r

    def transfer(self, tag, bufs):
        d = self.decomp
        neigh = [peer for peer, _, _ in d._halo_slices()]
        if not neigh:
            return
        send0 = bufs[("send", neigh[0])]
        nelem = max(bufs[("send", p)].numel() for p in neigh)
        r = self.region(tag, nelem, send0.dtype, send0.device)
        hdl = r["hdl"]
        for peer in neigh:
            send = bufs[("send", peer)]
            hdl.wait_signal(peer, r["ack"])          # the neighbour has consumed my previous message
            inbox = hdl.get_buffer(peer, (2, r["nelem"]), send.dtype)
            inbox[self._slot_in_peer(peer), :send.numel()].copy_(send)   # NVLink peer stores
            hdl.put_signal(peer, r["data"])
        for peer in neigh:
            hdl.wait_signal(peer, r["data"])
            bufs[("recv", peer)] = r["t"][self._slot_of_peer(peer), :bufs[("send", peer)].numel()]

    def release(self, tag):
        """after unpack: hand the inbox back to the neighbours"""
        r = self.regions.get(tag)
        if r is None:
            return
        for peer, _, _ in self.decomp._halo_slices():
            r["hdl"].put_signal(peer, r["ack"])


class SlabDecomposition:
    """Row slabs of a global (B, C, D, H, W) grid over the ranks of a process group.  Tensors handled
    by this class are GLOBAL-SIZED on every rank; only the window rows are meaningful."""

    def __init__(self, global_rows, ghost, group=None, rank=None, world=None, axis=3, align=4, comm=None,
                 transport="nccl"):
        self.group = group
        # transport: "nccl" = batched send/recv; "peer" = stores into the neighbour's inbox over NVLink
        # peer memory (PeerMemoryTransport), capturable in CUDA graphs
        self.transport = transport
        self.peer = None
        # comm: optional transport with exchange_rows(decomp, sends) / all_reduce(t, op) / all_gather(t)
        # replacing torch.distributed (the single-GPU tests run several virtual ranks in one process)
        self.comm = comm
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.axis = axis
        self.H = int(global_rows)
        if self.H % self.world:
            raise ValueError(f"{self.H} rows do not split evenly over {self.world} ranks")
        self.Hs = self.H // self.world
        if self.world > 1 and (self.Hs % align or ghost % align):
            raise ValueError(f"slab height {self.Hs} and ghost width {ghost} must be multiples of {align} "
                             "(pyramid alignment of the multi-scale CNN)")
        if self.world > 1 and ghost > self.Hs:
            raise ValueError(f"ghost width {ghost} exceeds the slab height {self.Hs}")
        self.ghost = int(ghost) if self.world > 1 else 0
        self.lo = self.rank * self.Hs
        self.hi = self.lo + self.Hs
        self.g_top = self.ghost if self.rank > 0 else 0
        self.g_bot = self.ghost if self.rank < self.world - 1 else 0
        self.r0, self.r1 = self.lo - self.g_top, self.hi + self.g_bot      # the window
        self.local_rows = self.r1 - self.r0

    # -- slicing ----------------------------------------------------------------------------
    def _sl(self, a, b):
        idx = [slice(None)] * 5
        idx[self.axis] = slice(a, b)
        return tuple(idx)

    def rows(self, plane_rows=1):
        """window as a (row_begin, row_end) pair of the kernels' flattened (D*H) row space"""
        return (self.r0 * plane_rows, self.r1 * plane_rows)

    def scatter(self, full):
        """this rank's working copy of a replicated global tensor (global-sized)"""
        return full.clone()

    def owned(self, t):
        return t[self._sl(self.lo, self.hi)]

    def window(self, t):
        """compact contiguous copy of the window rows"""
        return t[self._sl(self.r0, self.r1)].contiguous()

    def put_window(self, full, compact):
        full[self._sl(self.r0, self.r1)] = compact
        return full

    def gather(self, t):
        """global tensor (on every rank) assembled from every rank's owned rows"""
        mine = self.owned(t).contiguous()
        if self.world == 1:
            return mine
        if self.comm is not None:
            parts = self.comm.all_gather(self, mine)
        else:
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine, group=self.group)
        return torch.cat(parts, dim=self.axis)

    # -- halo exchange ------------------------------------------------------------------------
    # pack (device copies) -> transfer (the only communication) -> unpack (device copies); the three
    # are separate so the copies can live inside CUDA graphs while the NCCL calls stay outside.
    def _halo_slices(self):
        g, out = self.ghost, []
        if self.rank > 0:                   # (peer, rows I send, rows I receive)
            out.append((self.rank - 1, (self.lo, self.lo + g), (self.lo - g, self.lo)))
        if self.rank < self.world - 1:
            out.append((self.rank + 1, (self.hi - g, self.hi), (self.hi, self.hi + g)))
        return out

    def pack(self, tensors, bufs):
        """copy my boundary owned rows of every tensor into one send buffer per neighbour
        (bufs: dict reused across steps so the buffers are static)"""
        for peer, (a, b), _ in self._halo_slices():
            n = sum(t[self._sl(a, b)].numel() for t in tensors)
            key = ("send", peer)
            if key not in bufs or bufs[key].numel() != n:
                bufs[key] = torch.empty(n, dtype=tensors[0].dtype, device=tensors[0].device)
                bufs[("recv", peer)] = torch.empty_like(bufs[key])
            o = 0
            for t in tensors:
                v = t[self._sl(a, b)]
                bufs[key][o:o + v.numel()].view(v.shape).copy_(v)
                o += v.numel()

    @property
    def in_graph_transfers(self):
        """True when halo transfers are plain device work (no host-side communication call)"""
        return self.world > 1 and self.transport == "peer" and self.comm is None

    def transfer(self, bufs, tag="halo"):
        """send buffers -> the neighbours' receive buffers (batched NCCL / gloo send+recv, or peer
        memory stores)"""
        if self.world == 1:
            return
        if self.in_graph_transfers:
            if self.peer is None:
                self.peer = PeerMemoryTransport(self)
            self.peer.transfer(tag, bufs)
            return
        if self.comm is not None:
            recvs = self.comm.exchange_rows(self, {peer: bufs[("send", peer)] for peer, _, _ in self._halo_slices()})
            for peer, buf in recvs.items():
                bufs[("recv", peer)].copy_(buf)
            return
        ops = []
        for peer, _, _ in self._halo_slices():
            ops += [dist.P2POp(dist.isend, bufs[("send", peer)], self._peer(peer), self.group),
                    dist.P2POp(dist.irecv, bufs[("recv", peer)], self._peer(peer), self.group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    def unpack(self, tensors, bufs, tag="halo"):
        for peer, _, (a, b) in self._halo_slices():
            o = 0
            for t in tensors:
                v = t[self._sl(a, b)]
                v.copy_(bufs[("recv", peer)][o:o + v.numel()].view(v.shape))
                o += v.numel()
        if self.in_graph_transfers and self.peer is not None:
            self.peer.release(tag)

    def exchange(self, tensors, bufs=None, tag="halo"):
        """Refresh the ghost rows of every tensor in `tensors` (in place) from the neighbours' owned
        rows: one packed message per neighbour and direction."""
        if self.world == 1:
            return
        bufs = {} if bufs is None else bufs
        self.pack(tensors, bufs)
        self.transfer(bufs, tag)
        self.unpack(tensors, bufs, tag)

    def _peer(self, r):
        return r if self.group is None else dist.get_global_rank(self.group, r)

    def all_reduce(self, t, op=dist.ReduceOp.SUM):
        if self.world > 1:
            if self.comm is not None:
                self.comm.all_reduce(self, t, op)
            else:
                dist.all_reduce(t, op=op, group=self.group)
        return t


# ---------------------------------------------------------------------------------------------
class CudaLocalOps:
    """The per-window compute: the single-GPU kernels behind the C-ABI (no fallback)."""

    def __init__(self):
        import importlib
        from .. import _native as N
        # (the package attribute `lib.simulate` is the function; the helpers live in the module)
        sim = importlib.import_module(__package__ + ".simulate")
        self.N, self.sim, self.lib = N, sim, N.load()
        self._keep = {}     # per-row mask maps handed to kernels (a captured graph keeps pointing at them)
        self._pool = {}     # persistent zero-initialised output buffers (see _out)

    def _out(self, name, like, avoid=()):
        """A global-sized output buffer for `name`.  The kernels write the window rows only and the
        rows outside must stay finite (they are read as stale halo), so outputs are taken from a small
        pool of persistent buffers zeroed ONCE, instead of a fresh memset of the whole global array per
        step (at 8 ranks that memset cost more than the halo exchange).  Never returns a buffer that
        aliases one of `avoid` (the step's inputs / the previous Jacobi chunk)."""
        key = (name, tuple(like.shape), like.device)
        bufs = self._pool.setdefault(key, [])
        busy = {t.data_ptr() for t in avoid if t is not None}
        for b in bufs:
            if b.data_ptr() not in busy:
                return b
        b = torch.zeros_like(like)
        bufs.append(b)
        return b

    def advect_forces_div(self, mconf, dt, bd, want_div, wall_bcs, rows):
        N, sim, lib = self.N, self.sim, self.lib
        import ctypes
        flags, U_in, rho_in = bd['flags'], bd['U'], bd['density']
        # rows outside the window are never written: they stay at the pool's initial zeros
        U, density = self._out("U", U_in, (U_in,)), self._out("density", rho_in, (rho_in,))
        div = self._out("div", flags) if want_div else None
        B, D, H, W = N.grid_of(flags)
        is3d = int(U.size(1) == 3)
        UBC, UBCInv, rBC, rBCInv = sim._masks(bd)
        mrows = sim._mask_rows(lib, bd, flags, is3d)
        if mrows is not None:
            self._keep[mrows.data_ptr()] = mrows
        prm = sim._step_params(mconf, dt, 0)
        prm.apply_wall_bcs = int(wall_bcs)
        prm.density_const_passes = 2 if wall_bcs else 1
        prm.row_begin, prm.row_end = rows
        ws = N.workspaces.get(U.device, "step", lib.fnx_step_workspace(B, D, H, W, is3d))
        N.check(lib.fnx_step_advect_forces_div(ctypes.byref(prm), N.ptr(rho_in), N.ptr(U_in), N.ptr(flags),
                                               N.ptr(UBC), N.ptr(UBCInv), N.ptr(rBC), N.ptr(rBCInv),
                                               mrows.data_ptr() if mrows is not None else None, N.ptr(density),
                                               N.ptr(U), N.ptr(div), B, D, H, W, is3d, ws.data_ptr(), ws.numel(),
                                               N.stream_of(U)), "simulate_distributed")
        return density, U, div

    def jacobi(self, flags, div, p_init, iters, rows):
        N, lib = self.N, self.lib
        B, D, H, W = N.grid_of(flags)
        p = self._out("p", flags, (p_init,))
        ws = N.workspaces.get(flags.device, "jacobi", lib.fnx_jacobi_workspace(B, D, H, W, iters))
        N.check(lib.fnx_jacobi_iterate(N.ptr(flags), N.ptr(div), N.ptr(p_init), N.ptr(p), B, D, H, W,
                                       int(D > 1), int(iters), rows[0], rows[1], ws.data_ptr(), ws.numel(),
                                       N.stream_of(flags)), "simulate_distributed")
        return p

    def jacobi_resid(self, flags, div, p_init, iters, rows, own):
        """`iters` iterations, one per launch; returns p and this rank's (iters, B) sums over its OWNED rows of
        (p_it - p_{it-1})^2 (float64, device)."""
        N, lib = self.N, self.lib
        B, D, H, W = N.grid_of(flags)
        p = self._out("p", flags, (p_init,))
        ssq = torch.empty((iters, B), dtype=torch.float64, device=flags.device)
        ws = N.workspaces.get(flags.device, "jacobi", lib.fnx_jacobi_workspace(B, D, H, W, iters))
        N.check(lib.fnx_jacobi_iterate_resid(N.ptr(flags), N.ptr(div), N.ptr(p_init), N.ptr(p), B, D, H, W, int(D > 1),
                                             int(iters), rows[0], rows[1], own[0], own[1], ssq.data_ptr(),
                                             ws.data_ptr(), ws.numel(), N.stream_of(flags)), "simulate_distributed")
        return p, ssq

    def project(self, p, U, bd, rows):
        N, sim, lib = self.N, self.sim, self.lib
        flags = bd['flags']
        B, D, H, W = N.grid_of(flags)
        is3d = int(U.size(1) == 3)
        UBC, UBCInv, _, _ = sim._masks(bd)
        mrows = sim._mask_rows(lib, bd, flags, is3d)
        if mrows is not None:
            self._keep[mrows.data_ptr()] = mrows
        N.check(lib.fnx_step_project_bcs_rows(N.ptr(p), N.ptr(U), N.ptr(flags), N.ptr(UBC), N.ptr(UBCInv),
                                              mrows.data_ptr() if mrows is not None else None, 1, B, D, H, W, is3d,
                                              rows[0], rows[1], N.stream_of(U)), "simulate_distributed")
        return U

    def p_buffer(self, like, current_p):
        return self._out("p", like, (current_p,))

    def set_const(self, x, inv_mask, bc):
        from . import fluid
        fluid.setConstVals(x, inv_mask, bc)

    def cnn(self, net, U, flags, scale, prewall=False):
        if U.size(1) == 3:
            return net.forward_fields_3d(U, flags, scale=scale)
        return net.forward_fields(U, flags, scale=scale, prewall=prewall)


def _std_partial(decomp, U):
    """(sum, sum of squares) of this rank's owned rows of U, fp64 (to be all-reduced)"""
    own = decomp.owned(U).double()
    return torch.stack([own.sum(), (own * own).sum()])


def _std_finish(decomp, part, owned_count, threshold):
    """max(unbiased std of the GLOBAL U, threshold) (model.py:8-23) from the reduced partial sums"""
    n = float(owned_count * decomp.world)
    var = (part[1] - part[0] * part[0] / n) / (n - 1.0)
    s = torch.sqrt(torch.clamp(var, min=0.0)).float().clamp(min=float(threshold))
    return s.view(1, 1, 1, 1, 1)


def _seam_y(decomp, U, src_row, multi):
    """U[:, 0, :, 1] = <row H-1 of component 0 of the seam's source field> (the periodic-y seam of the saved model;
    component 0 as the reference writes it) across slabs: row H-1 is owned by the last rank -- `src_row` is that
    (B, D, W) line there and ignored elsewhere -- and row 1 is held by every rank whose window starts at row 0 or 1.
    The line travels as its int32 bit pattern through an all-reduce(sum) in which every other rank contributes
    zeros, so the copy is exact (-0.0 included).  A generator like step_phases: yields the communication."""
    H = decomp.H
    if not multi:
        U[:, 0, :, 1] = src_row
        return
    row = torch.zeros(U[:, 0, :, 0].shape, dtype=torch.int32, device=U.device)
    if decomp.lo <= H - 1 < decomp.hi:
        row.copy_(src_row.contiguous().view(torch.int32))
    yield lambda: decomp.all_reduce(row)
    if decomp.r0 <= 1 < decomp.r1:
        U[:, 0, :, 1] = row.view(torch.float32)


def check_reach(decomp, U_local, dt):
    """The ghost width assumes the back-trace moves at most one cell per step (RA): verify it.
    One all-reduce(max) and a host read -- call it every few steps, not every step."""
    m = decomp.owned(U_local).abs().max().reshape(1).clone()
    decomp.all_reduce(m, op=dist.ReduceOp.MAX)
    reach = float(m.item()) * abs(float(dt))
    if reach > 1.0:
        raise RuntimeError(f"slab decomposition: max|U|*dt = {reach:.3f} cells exceeds the advection reach the "
                           f"ghost width was sized for (1 cell); use a wider ghost or a smaller dt")
    return reach


def step_phases(mconf, bd, net, sim_method, decomp, ops, bufs):
    """The slab-decomposed step as a generator: pure device work between two `yield`s, and every yield
    hands back a zero-argument callable that performs one communication (halo transfer or the
    all-reduce).  simulate_distributed runs it straight through; GraphedDistributedStep captures each
    compute stretch into a CUDA graph and keeps the communication calls between the replays."""
    assert sim_method in ('jacobi', 'convnet')
    dt = float(mconf['dt'])
    g = decomp.ghost
    multi = decomp.world > 1
    if bd['U'].size(0) != 1:
        raise NotImplementedError("the slab-decomposed step handles one simulation per call (B = 1): the ScaleNet "
                                  "normalisation is reduced over the whole batch tensor")
    # periodic seams (Rayleigh-Taylor: periodic-y).  The ScaleNet model copies one line of U before the network and
    # one line of the pre-setWallBcs result after it (*_saved.py:123-132, 228-237).  periodic-x copies a column: local
    # to every row.  periodic-y copies row H-1 into row 1: across slabs the source row lives on the LAST rank and
    # the target on every rank whose window holds row 1 -- `_seam_y` moves it.  The Jacobi step's seam
    # (simulate.py:120-128, twice per step around setWallBcs inside the fused kernels) stays single-GPU.
    net_conf = getattr(net, 'mconf', None) or {}
    per = sim_method == 'convnet' and 'periodic-x' in net_conf and 'periodic-y' in net_conf
    per_x, per_y = bool(per and net_conf['periodic-x']), bool(per and net_conf['periodic-y'])
    if sim_method == 'jacobi' and multi and (mconf.get('periodic-x') or mconf.get('periodic-y')):
        raise NotImplementedError("the Jacobi step's periodic seam across slabs is not implemented (simulate.py:120-128 "
                                  "copies it around setWallBcs, which the fused kernels apply in place); run it on one "
                                  "GPU, or use the ScaleNet projection")
    if (per_x or per_y) and decomp.axis != 3:
        raise NotImplementedError("periodic seams are defined for the 2-D model only (model.py:93)")
    if multi:
        is3d_cnn = sim_method == 'convnet' and decomp.axis == 2
        need = RA + (CNN3D_PLANES if is3d_cnn else (CNN_REACH if sim_method == 'convnet' else 1))
        if g < need:
            raise ValueError(f"ghost width {g} < {need} rows needed by the {sim_method} step")
    plane = bd['flags'].size(3) if decomp.axis == 2 else 1     # 3-D slabs along D: H rows per plane
    rows = decomp.rows(plane)
    sb = bufs.setdefault('state', {})
    inline = decomp.in_graph_transfers      # peer-memory transfers are device work: no yield, one graph
    if multi:
        decomp.pack([bd['density'], bd['U']], sb)
        if inline:
            decomp.transfer(sb, "state")
        else:
            yield lambda: decomp.transfer(sb, "state")
        decomp.unpack([bd['density'], bd['U']], sb, "state")
    if sim_method == 'jacobi':
        density, U, div = ops.advect_forces_div(mconf, dt, bd, True, True, rows)
        iters = int(mconf['jacobiIter'])
        k = jacobi_chunk(g) if multi else iters
        p, done = None, 0
        pb = bufs.setdefault('p', {})
        p_tol = float(mconf.get('pTol', 0.0))
        own = (decomp.lo * plane, decomp.hi * plane)
        while done < iters:
            n = min(k, iters - done)
            if done > 0:
                decomp.pack([p], pb)
                if inline:
                    decomp.transfer(pb, "p")
                else:
                    yield lambda: decomp.transfer(pb, "p")
                decomp.unpack([p], pb, "p")
            if p_tol > 0:
                # residual-terminated (fluids_init.cpp:958-990: max_b ||p - p_prev||_2 < p_tol, tested after every
                # iteration): every rank sums its owned rows per iteration, one all-reduce(sum) per chunk of n
                # iterations, and the chunk is redone up to the stopping iteration when that falls inside it --
                # the iteration count, hence p, is the single-GPU solver's.  Reads the result on the host (as the
                # reference does every iteration): this path is not graph-captured.
                p_start = p
                p, ssq = ops.jacobi_resid(bd['flags'], div, p_start, n, rows, own)
                if multi:
                    yield lambda: decomp.all_reduce(ssq)
                r = ssq.sqrt().to(torch.float32).max(dim=1).values
                below = (r < torch.tensor(p_tol, dtype=torch.float32)).nonzero()
                if below.numel():
                    hit = int(below[0])
                    if hit < n - 1:
                        p, _ = ops.jacobi_resid(bd['flags'], div, p_start, hit + 1, rows, own)
                    bd['jacobi_iterations'], bd['residual'] = done + hit + 1, r[hit]
                    break
                bd['jacobi_iterations'], bd['residual'] = done + n, r[-1]
            else:
                p = ops.jacobi(bd['flags'], div, p, n, rows)
            done += n
        U = ops.project(p, U, bd, rows)      # radius 1: p is valid from row ghost-1 inwards after any chunk
    else:
        density, U, _ = ops.advect_forces_div(mconf, dt, bd, False, False, rows)
        if per_x:       # U[:,1,:,:,1] = U[:,1,:,:,W-1] (*_saved.py:123-132), every window row
            U[:, 1, :, decomp.r0:decomp.r1, 1] = U[:, 1, :, decomp.r0:decomp.r1, U.size(4) - 1]
        if per_y:       # U[:,0,:,1] = U[:,0,:,H-1]
            yield from _seam_y(decomp, U, U[:, 0, :, decomp.H - 1].clone(), multi)
        part = _std_partial(decomp, U)
        if multi:
            yield lambda: decomp.all_reduce(part)
        scale = _std_finish(decomp, part, decomp.owned(U).numel(), net.mconf['normalizeInputThreshold'])
        if decomp.axis == 2:
            # 3-D, slice-wise projection: a compact copy of the owned planes +- CNN3D_PLANES (the first / last plane of
            # the copy is treated as a border plane by the 3-D stencils: that only touches ghost planes)
            c0 = max(decomp.lo - CNN3D_PLANES, decomp.r0)
            c1 = min(decomp.hi + CNN3D_PLANES, decomp.r1)
            sl = decomp._sl(c0, c1)
            p_w, U_w = ops.cnn(net, U[sl].contiguous(), bd['flags'][sl].contiguous(), scale)
            p = ops.p_buffer(bd['flags'], bd.get('p'))
            p[sl] = p_w
            U[sl] = U_w
        else:
            # the CNN runs on a compact copy of the window (translation invariant), results go back in place
            if per_x or per_y:
                p_w, U_w, Ut_w = ops.cnn(net, decomp.window(U), decomp.window(bd['flags']), scale, prewall=True)
                if per_x:
                    U_w[:, 1, :, :, 1] = Ut_w[:, 1, :, :, U_w.size(4) - 1]
            else:
                p_w, U_w = ops.cnn(net, decomp.window(U), decomp.window(bd['flags']), scale)
            p = decomp.put_window(ops.p_buffer(bd['flags'], bd.get('p')), p_w)
            U = decomp.put_window(U, U_w)
            if per_y:   # the seam source is the corrected velocity BEFORE setWallBcs (*_saved.py:228-237)
                yield from _seam_y(decomp, U, Ut_w[:, 0, :, min(decomp.H - 1 - decomp.r0, Ut_w.size(3) - 1)], multi)
        if 'UBC' in bd and 'UBCInvMask' in bd:
            ops.set_const(U, bd['UBCInvMask'], bd['UBC'])
        if 'densityBC' in bd and 'densityBCInvMask' in bd:
            ops.set_const(density, bd['densityBCInvMask'], bd['densityBC'])
    bd['U'], bd['density'], bd['p'] = U, density, p


def simulate_distributed(mconf, bd, net, sim_method, decomp, ops=None, bufs=None):
    """One solver step on this rank's window (bd holds GLOBAL-SIZED tensors; masks included).
    Same state transitions as lib.simulate for the fused configuration (inviscid, density-carrying,
    fixed Jacobi count or the ScaleNet model).  Only the owned rows of the returned state are
    meaningful (decomp.owned / decomp.gather); ghost rows are refreshed by the next call.
    The ghost width assumes max|U| dt <= 1 cell per step: call check_reach() every few steps (GraphedDistributedStep
    does it by itself) -- a faster flow contaminates owned rows without any other symptom.
    Returns nothing; bd['p'], bd['U'], bd['density'] are rebound -- to buffers of `ops`' output pool,
    which are recycled two steps later: clone what must outlive the next calls (pass the same `ops`
    every step; a fresh one allocates a fresh pool)."""
    ops = ops or CudaLocalOps()
    for comm in step_phases(mconf, bd, net, sim_method, decomp, ops, {} if bufs is None else bufs):
        comm()


class GraphedDistributedStep:
    """The slab-decomposed step with every stretch of device work between two communications captured
    ONCE into a CUDA graph and replayed per step; the NCCL calls (one batched halo send/recv per
    exchange, one all-reduce for the ScaleNet std) are issued between the replays.  Below a few
    million cells per GPU the step is launch-bound: a replay removes ~80 launches and the Python
    between them.

        stepper = GraphedDistributedStep(mconf, bd, net, 'convnet', decomp)
        stepper.step(); ...; state = stepper.state      # global-sized tensors, owned rows valid

    Falls back to direct launches (same results) when capture is not possible (`graphed` says which).
    """

    def __init__(self, mconf, bd, net, sim_method, decomp, ops=None, use_graph=True, warmup=2, check_reach_every=64):
        # The ghost width is sized for a back-trace of at most one cell per step (RA): a faster flow would silently
        # contaminate owned rows.  The stepper therefore verifies max|U| dt itself every `check_reach_every` steps
        # (one all-reduce(max) and a host read, amortised; 0 = the caller takes care of it with check_reach()).
        self.check_reach_every, self._steps = int(check_reach_every), 0
        self.mconf, self.net, self.method, self.decomp = mconf, net, sim_method, decomp
        self.ops = ops or CudaLocalOps()
        self.state = {k: (v.clone() if k in ('p', 'U', 'density') else v) for k, v in bd.items()}
        self.bufs = {}
        self.schedule, self.graphed, self.capture_error = None, False, None
        self._win = decomp._sl(decomp.r0, decomp.r1)
        if sim_method == 'jacobi' and float(mconf.get('pTol', 0.0)) > 0:
            use_graph = False       # the residual test reads device results on the host every chunk
        if not (use_graph and decomp.comm is None and self.state['flags'].is_cuda):
            return
        # warm-up on a scratch copy: NCCL communicators, workspaces, halo buffers, the CNN plan
        scratch = dict(self.state)
        for k in ('p', 'U', 'density'):
            scratch[k] = self.state[k].clone()
        with torch.no_grad():
            for _ in range(max(1, warmup)):
                simulate_distributed(mconf, scratch, net, sim_method, decomp, ops=self.ops, bufs=self.bufs)
        torch.cuda.synchronize()
        try:
            self._capture()
            self.graphed = True
        except Exception as e:      # noqa: BLE001 - capture is an optimisation, never a requirement
            self.capture_error = repr(e)
            self.schedule = None
            torch.cuda.synchronize()
        if decomp.world > 1:        # all ranks take the same path
            ok = torch.tensor([1.0 if self.graphed else 0.0], device=self.state['flags'].device)
            decomp.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() < 1.0:
                self.schedule, self.graphed = None, False

    def _capture(self):
        work = dict(self.state)
        gen = step_phases(self.mconf, work, self.net, self.method, self.decomp, self.ops, self.bufs)
        schedule, pool, done = [], None, False
        with torch.no_grad():
            while not done:
                graph = torch.cuda.CUDAGraph()
                comm = None
                with torch.cuda.graph(graph, pool=pool):
                    try:
                        comm = next(gen)
                    except StopIteration:
                        done = True
                pool = graph.pool()
                schedule.append((graph, comm))
        self.schedule = schedule
        self.outs = {k: work[k] for k in ('p', 'U', 'density')}
        # A CUDA graph holds raw pointers, not references: the tensors it READS must stay alive and must stay
        # the ones `state` names.  (Round 1 lost both: bench.py's per-stage pass re-bound the entries of
        # `stepper.state`, the original tensors were freed and re-used by the allocator, and every later replay
        # advected whatever bytes landed there -- at 4 ranks an infinite velocity, i.e. a line trace that never
        # ends: the "4-GPU hang".)
        self._inputs = {k: self.state[k] for k in ('p', 'U', 'density', 'flags')}

    def verify(self):
        """one step through the graphs and one by direct launches from the same state: max |difference|
        over this rank's owned rows (0.0 = bit-identical).  Leaves the state advanced by one step."""
        start = {k: self.state[k].clone() for k in ('p', 'U', 'density')}
        self.step()
        got = {k: self.decomp.owned(self.state[k]).clone() for k in ('p', 'U', 'density')}
        ref = dict(self.state)
        ref.update(start)
        with torch.no_grad():
            simulate_distributed(self.mconf, ref, self.net, self.method, self.decomp, ops=self.ops, bufs={})
        return max(float((self.decomp.owned(ref[k]) - got[k]).abs().max()) for k in got)

    def step(self):
        if self.check_reach_every and self._steps % self.check_reach_every == 0 and self.decomp.world > 1:
            check_reach(self.decomp, self.state['U'], self.mconf['dt'])
        self._steps += 1
        if self.graphed:
            for k in ('p', 'U', 'density'):
                if self.state[k] is not self._inputs[k]:       # someone re-bound an entry: take its rows, keep our buffer
                    self._inputs[k][self._win].copy_(self.state[k][self._win])
                    self.state[k] = self._inputs[k]
            for graph, comm in self.schedule:
                graph.replay()
                if comm is not None:
                    comm()
            for k in ('p', 'U', 'density'):
                self.state[k][self._win].copy_(self.outs[k][self._win])
        else:
            with torch.no_grad():
                simulate_distributed(self.mconf, self.state, self.net, self.method, self.decomp, ops=self.ops,
                                     bufs=self.bufs)
