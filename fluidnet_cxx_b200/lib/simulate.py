"""Per-timestep orchestrator: same call, same state transitions as the reference
`lib.simulate` (pytorch/lib/simulate.py:28-171), with the standard sequence routed through the
fused sm_100a entry points.

    simulate(mconf, batch_dict, net, sim_method, output_div=False)

batch_dict {'p','U','flags','density', [UBC, UBCInvMask, densityBC, densityBCInvMask]} is
advanced in place (the dict entries are rebound to the new tensors, as in the reference).
Sequence: advectScalar -> advectVelocity -> setConstVals -> addBuoyancy / addGravity ->
[setWallBcs] -> setConstVals -> (Jacobi | net) -> [velocityUpdate -> setWallBcs] -> setConstVals.

Also accepts the legacy 5-argument form `simulate(conf, mconf, batch_dict, net, sim_method)` that
pytorch/rayleighTaylor.py:240 still uses (SURVEY.md hazard H3).
"""
import ctypes
import os

import torch

from .. import _native as N
from . import fluid


_stage_hook = None


def set_stage_hook(fn):
    """fn(stage_name, 'begin'|'end') is called around the stages of the fused step ('advect_forces',
    'pressure', 'project'; bench.py records CUDA events there); None restores the single-call path."""
    global _stage_hook
    _stage_hook = fn


def setConstVals(batch_dict, p, U, flags, density):
    """Apply the imposed-value masks (simulate.py:4-26): x = x*InvMask + BC, and publish clones."""
    if ('UBCInvMask' in batch_dict) and ('UBC' in batch_dict):
        fluid.setConstVals(U, batch_dict['UBCInvMask'], batch_dict['UBC'])
        batch_dict['U'] = U.clone()
    if ('densityBCInvMask' in batch_dict) and ('densityBC' in batch_dict):
        fluid.setConstVals(density, batch_dict['densityBCInvMask'], batch_dict['densityBC'])
        batch_dict['density'] = density.clone()


def _gravity(mconf, scale):
    g = mconf['gravityVec']
    # the reference builds an fp32 tensor and multiplies it by -scale in fp32 (simulate.py:101-105)
    t = torch.tensor([g['x'], g['y'], g['z']], dtype=torch.float32)
    t.mul_(-scale)
    return t


def _periodic(mconf):
    return 'periodic-x' in mconf and 'periodic-y' in mconf


def _wall_bcs_with_seam(U, flags, mconf):
    """setWallBcs plus the periodic seam copy of simulate.py:120-128."""
    per = _periodic(mconf)
    if per:
        U_temp = U.clone()
    U = fluid.setWallBcs(U, flags)
    if per:
        if mconf['periodic-x']:
            U[:, 1, :, :, 1] = U_temp[:, 1, :, :, U.size(4) - 1]
        if mconf['periodic-y']:
            U[:, 0, :, 1] = U_temp[:, 0, :, U.size(3) - 1]
    return U


def _step_params(mconf, dt, jacobi_iters=0):
    prm = N.StepParams()
    prm.dt = dt
    prm.maccormack_strength = float(mconf['maccormackStrength'])
    prm.sample_outside_fluid = int(bool(mconf['sampleOutsideFluid']))
    bs, gs = mconf['buoyancyScale'], mconf['gravityScale']
    prm.use_buoyancy = int(bs > 0)
    prm.use_gravity = int(gs > 0)
    if bs > 0:
        prm.buoyancy3 = (ctypes.c_float * 3)(*_gravity(mconf, bs).tolist())
    if gs > 0:
        prm.gravity3 = (ctypes.c_float * 3)(*_gravity(mconf, gs).tolist())
    prm.rho_star = float(mconf['operatingDensity']) if bs > 0 else 0.0
    prm.jacobi_iters = int(jacobi_iters)
    return prm


def _fusable(mconf, batch_dict, sim_method, output_div=False):
    """The fused entry points cover the inviscid density-carrying step of both drivers."""
    if 'density' not in batch_dict or 'flags_stick' in batch_dict:
        return False
    if mconf['viscosity'] != 0 or mconf.get('correctScalar', False):
        return False
    has_u = ('UBCInvMask' in batch_dict) and ('UBC' in batch_dict)
    has_r = ('densityBCInvMask' in batch_dict) and ('densityBC' in batch_dict)
    if has_u != has_r and (('UBC' in batch_dict) != ('UBCInvMask' in batch_dict)):
        return False
    if output_div:
        return False
    if sim_method == 'jacobi':
        if _periodic(mconf) and (mconf['periodic-x'] or mconf['periodic-y']):
            return False
        if mconf['pTol'] > 0 or mconf['jacobiIter'] < 1:
            return False
    return all(batch_dict[k].is_cuda and batch_dict[k].is_contiguous() and batch_dict[k].dtype == torch.float32
               for k in ('p', 'U', 'flags', 'density'))


def _masks(batch_dict):
    has_u = ('UBCInvMask' in batch_dict) and ('UBC' in batch_dict)
    has_r = ('densityBCInvMask' in batch_dict) and ('densityBC' in batch_dict)
    return (batch_dict['UBC'] if has_u else None, batch_dict['UBCInvMask'] if has_u else None,
            batch_dict['densityBC'] if has_r else None, batch_dict['densityBCInvMask'] if has_r else None)


def simulate(*args, **kwargs):
    if len(args) >= 3 and isinstance(args[2], dict) and isinstance(args[1], dict):
        args = args[1:]           # legacy (conf, mconf, batch_dict, net, sim_method)
    return _simulate(*args, **kwargs)


def _simulate(mconf, batch_dict, net, sim_method, output_div=False):
    assert sim_method == 'convnet' or sim_method == 'jacobi', \
        'Simulation method not supported. Choose either convnet or jacobi.'
    dt = float(mconf['dt'])
    assert mconf['viscosity'] >= 0, 'Viscosity must be positive'
    if _fusable(mconf, batch_dict, sim_method, output_div):
        if _graphable(batch_dict, net, sim_method):
            return _simulate_graphed(mconf, batch_dict, net, sim_method, dt)
        return _simulate_fused(mconf, batch_dict, net, sim_method, dt, output_div)
    return _simulate_ops(mconf, batch_dict, net, sim_method, dt, output_div)


# ---------------------------------------------------------------------------------------------
# CUDA-graph replay of the fused step.  Below a few million cells the
# step is launch-bound (4 stencil kernels + ~50 launches of the CNN forward): the whole fixed
# sequence is captured once per (grid, configuration, masks, weights) into a CUDA graph that works
# on graph-owned state buffers; a step is then: copy the caller's state in, one graph launch, hand
# out clones (the reference returns fresh tensors every step and never mutates the old ones).
GRAPH_MAX_CELLS = 1 << 22
GRAPH_MAX_CELLS_3D_CNN = 1 << 25
_graphs = {}
_graph_seen = {}


def graphs_enabled():
    return os.environ.get("FLUIDNET_B200_GRAPHS", "1") != "0"


def clear_graph_cache():
    _graphs.clear()
    _graph_seen.clear()


def _native_net(net):
    """This package's FluidNet for `net`, or None.

    The reference drivers do not use lib.FluidNet: they execute the model class saved next to the weights
    (plume.py:87-96,131: trained_models/.../*_saved.py), whose forward is eager torch code (std, cat, masks,
    lib.fluid calls) around `lib.MultiScaleNet`.  Under compat.install() that MultiScaleNet is this package's
    tcgen05 network; when the instance has the saved model's structure and the shipped configuration, its
    forward is routed through FluidNet.forward (same arithmetic: tests/test_gpu_cnn.py pins it against the
    saved model's own CPU outputs) sharing the instance's weights, so the drivers take the fused, graph-replayed
    path.  Anything else (other model variants, foreign modules) keeps running its own forward."""
    from .model import FluidNet
    from .multi_scale_net import MultiScaleNet
    if isinstance(net, FluidNet):
        return net
    if net is None or type(net).__name__ != 'FluidNet' or not hasattr(net, 'mconf'):
        return None
    ms = getattr(net, 'multiScale', None)
    if not isinstance(ms, MultiScaleNet):
        return None
    cached = net.__dict__.get('_fnx_native')
    if cached is not None and cached.mconf is net.mconf and cached.multiScale is ms:
        return cached
    try:
        native = FluidNet(net.mconf, dropout=False)
        native._check_config()
    except (NotImplementedError, AssertionError, KeyError):
        return None
    native.multiScale = ms            # the instance's own (already loaded, already on the device) weights
    native.eval()
    net.__dict__['_fnx_native'] = native
    return native


def _graphable(batch_dict, net, sim_method):
    if not graphs_enabled() or _stage_hook is not None:
        return False
    f = batch_dict['flags']
    # (the slice-wise 3-D CNN step is ~10^4 small launches: it is launch-bound at any size)
    limit = GRAPH_MAX_CELLS_3D_CNN if (sim_method == 'convnet' and f.size(2) > 1) else GRAPH_MAX_CELLS
    if f.numel() > limit or torch.cuda.is_current_stream_capturing():
        return False
    if sim_method == 'convnet' and _native_net(net) is None:
        return False                          # a foreign nn.Module: its launches may not be capturable
    return True


def _freeze(v):
    if isinstance(v, dict):
        return tuple(sorted((k, _freeze(x)) for k, x in v.items()))
    if isinstance(v, (list, tuple)):
        return tuple(_freeze(x) for x in v)
    if isinstance(v, torch.Tensor):
        return (v.data_ptr(), v._version)
    return v


# every mconf entry the fused step / the model wrapper reads
_STEP_KEYS = ('dt', 'maccormackStrength', 'sampleOutsideFluid', 'buoyancyScale', 'gravityScale', 'gravityVec',
              'operatingDensity', 'jacobiIter', 'pTol', 'viscosity', 'correctScalar', 'periodic-x', 'periodic-y')
_NET_KEYS = ('normalizeInputThreshold', 'normalizeInput', 'normalizeInputChan', 'inputChannels', 'model', 'is3D',
             'periodic-x', 'periodic-y')


def _graph_key(mconf, batch_dict, net, sim_method):
    state = tuple((k, tuple(batch_dict[k].shape)) for k in ('p', 'U', 'flags', 'density'))
    masks = tuple((t.data_ptr(), t._version, tuple(t.shape)) if t is not None else None for t in _masks(batch_dict))
    conf = tuple(_freeze(mconf.get(k)) for k in _STEP_KEYS)
    netk = None
    if sim_method == 'convnet':
        nn_ = _native_net(net)
        netk = (id(net), nn_.multiScale._plan_key(batch_dict['U'].device),
                tuple(_freeze(nn_.mconf.get(k)) for k in _NET_KEYS))
    return (str(batch_dict['U'].device), sim_method, state, masks, conf, netk)


def _simulate_graphed(mconf, batch_dict, net, sim_method, dt):
    key = _graph_key(mconf, batch_dict, net, sim_method)
    entry = _graphs.get(key)
    if entry is None:
        # first sighting: run eagerly (this also performs every lazy initialisation: workspaces, the
        # CNN plan, kernel attributes); capture on the second call with the same key
        if _graph_seen.get(key, 0) < 1:
            if len(_graph_seen) > 64:
                _graph_seen.clear()
            _graph_seen[key] = _graph_seen.get(key, 0) + 1
            return _simulate_fused(mconf, batch_dict, net, sim_method, dt, False)
        if len(_graphs) >= 8:
            _graphs.clear()
        static = {k: batch_dict[k].clone() for k in ('p', 'U', 'flags', 'density')}
        work = dict(batch_dict)
        work.update(static)
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        try:
            with torch.cuda.graph(graph):
                _simulate_fused(mconf, work, net, sim_method, dt, False)
        except Exception:      # noqa: BLE001 - capture is an optimisation: fall back to direct launches
            torch.cuda.synchronize()
            _graph_seen[key] = -(1 << 30)          # do not try again for this configuration
            return _simulate_fused(mconf, batch_dict, net, sim_method, dt, False)
        is3d = int(batch_dict['U'].size(1) == 3)
        entry = (graph, static, {k: work[k] for k in ('p', 'U', 'density')},
                 (_masks(batch_dict), _mask_rows(N.load(), batch_dict, batch_dict['flags'], is3d)))
        _graphs[key] = entry
    graph, static, outs, _keepalive = entry
    for k in ('p', 'U', 'flags', 'density'):
        static[k].copy_(batch_dict[k])
    graph.replay()
    for k in ('p', 'U', 'density'):
        batch_dict[k] = outs[k].clone()


# ---------------------------------------------------------------------------------------------
_mask_rows_cache = {}


def _mask_rows(lib, batch_dict, flags, is3d):
    """Per-row "mask differs from identity" bytes (fnx_mask_rows), cached on the mask tensors'
    storage + version counters so it is recomputed only when a mask is modified."""
    UBC, UBCInv, rBC, rBCInv = _masks(batch_dict)
    if UBC is None and rBC is None:
        return None
    key = tuple((t.data_ptr(), t._version, tuple(t.shape)) if t is not None else None
                for t in (UBC, UBCInv, rBC, rBCInv))
    dev_cache = _mask_rows_cache.setdefault(flags.device, {})
    hit = dev_cache.get(key)
    if hit is not None:
        return hit[0]
    B, D, H, W = N.grid_of(flags)
    rows = torch.empty(B * D * H, dtype=torch.uint8, device=flags.device)
    N.check(lib.fnx_mask_rows(N.ptr(UBC), N.ptr(UBCInv), N.ptr(rBC), N.ptr(rBCInv), rows.data_ptr(), B, D, H, W,
                              is3d, N.stream_of(flags)), "simulate")
    if len(dev_cache) >= 16:
        dev_cache.pop(next(iter(dev_cache)))
    dev_cache[key] = (rows, (UBC, UBCInv, rBC, rBCInv))   # keep the mask tensors alive
    return rows


def _simulate_fused(mconf, batch_dict, net, sim_method, dt, output_div):
    lib = N.load()
    flags = batch_dict['flags']
    U_in, rho_in = batch_dict['U'], batch_dict['density']
    # the new state goes to new tensors (the reference never mutates the caller's old U / density)
    U = torch.empty_like(U_in)
    density = torch.empty_like(rho_in)
    B, D, H, W = N.grid_of(flags)
    is3d = int(U.size(1) == 3)
    UBC, UBCInv, rBC, rBCInv = _masks(batch_dict)
    rows = _mask_rows(lib, batch_dict, flags, is3d)
    rows_ptr = rows.data_ptr() if rows is not None else None
    ws = N.workspaces.get(U.device, "step", lib.fnx_step_workspace(B, D, H, W, is3d))
    st = N.stream_of(U)
    if sim_method == 'convnet':
        # advection + BCs + forces + BCs in the fused kernels (no setWallBcs ahead of the CNN,
        # simulate.py:120-133), then the model, then the last setConstVals (simulate.py:168)
        prm = _step_params(mconf, dt, 0)
        prm.apply_wall_bcs = 0
        prm.density_const_passes = 1
        if _stage_hook is not None:
            _stage_hook("advect_forces", "begin")
        N.check(lib.fnx_step_advect_forces_div(ctypes.byref(prm), N.ptr(rho_in), N.ptr(U_in), N.ptr(flags),
                                               N.ptr(UBC), N.ptr(UBCInv), N.ptr(rBC), N.ptr(rBCInv), rows_ptr,
                                               N.ptr(density), N.ptr(U), None, B, D, H, W, is3d,
                                               ws.data_ptr(), ws.numel(), st), "simulate")
        if _stage_hook is not None:
            _stage_hook("advect_forces", "end")
        net.eval()
        if _stage_hook is not None:
            _stage_hook("pressure", "begin")
        data = torch.cat((batch_dict['p'], U, flags, density), 1)
        p, U = (_native_net(net) or net)(data)
        if _stage_hook is not None:
            _stage_hook("pressure", "end")
        if UBC is not None:
            fluid.setConstVals(U, UBCInv, UBC)
        if rBC is not None:
            fluid.setConstVals(density, rBCInv, rBC)
        batch_dict['U'], batch_dict['density'], batch_dict['p'] = U, density, p
        return
    prm = _step_params(mconf, dt, mconf['jacobiIter'])
    p = torch.empty_like(flags)
    # (the reference's simulate drops the solver's residual, simulate.py:150-153: not computed here)
    if _stage_hook is None:
        # one C-ABI call for the whole step
        N.check(lib.fnx_step_jacobi(ctypes.byref(prm), N.ptr(rho_in), N.ptr(U_in), N.ptr(flags), N.ptr(UBC),
                                    N.ptr(UBCInv), N.ptr(rBC), N.ptr(rBCInv), rows_ptr, N.ptr(density), N.ptr(U),
                                    N.ptr(p), None, B, D, H, W, is3d, ws.data_ptr(), ws.numel(), st),
                "simulate")
    else:
        # same kernels, issued stage by stage so a profiler hook can bracket the pressure solve
        prm.apply_wall_bcs = 1
        prm.density_const_passes = 2
        div = torch.empty_like(flags)
        _stage_hook("advect_forces", "begin")
        N.check(lib.fnx_step_advect_forces_div(ctypes.byref(prm), N.ptr(rho_in), N.ptr(U_in), N.ptr(flags),
                                               N.ptr(UBC), N.ptr(UBCInv), N.ptr(rBC), N.ptr(rBCInv), rows_ptr,
                                               N.ptr(density), N.ptr(U), N.ptr(div), B, D, H, W, is3d,
                                               ws.data_ptr(), ws.numel(), st), "simulate")
        _stage_hook("advect_forces", "end")
        wj = N.workspaces.get(U.device, "jacobi", lib.fnx_jacobi_workspace(B, D, H, W, prm.jacobi_iters))
        _stage_hook("pressure", "begin")
        N.check(lib.fnx_solve_linear_system_jacobi(N.ptr(flags), N.ptr(div), N.ptr(p), None, B, D, H,
                                                   W, is3d, 0.0, prm.jacobi_iters, None, wj.data_ptr(), wj.numel(),
                                                   st), "simulate")
        _stage_hook("pressure", "end")
        _stage_hook("project", "begin")
        N.check(lib.fnx_step_project_bcs(N.ptr(p), N.ptr(U), N.ptr(flags), N.ptr(UBC), N.ptr(UBCInv), rows_ptr, 1,
                                         B, D, H, W, is3d, st), "simulate")
        _stage_hook("project", "end")
    batch_dict['U'], batch_dict['density'], batch_dict['p'] = U, density, p


def _simulate_ops(mconf, batch_dict, net, sim_method, dt, output_div):
    """Operator-by-operator sequence, line for line the order of simulate.py:28-171."""
    maccormackStrength = mconf['maccormackStrength']
    sampleOutsideFluid = mconf['sampleOutsideFluid']
    buoyancyScale = mconf['buoyancyScale']
    gravityScale = mconf['gravityScale']
    viscosity = mconf['viscosity']

    p = batch_dict['p']
    U = batch_dict['U']
    flags = batch_dict['flags']
    stick = 'flags_stick' in batch_dict
    if stick:
        flags_stick = batch_dict['flags_stick']

    if viscosity > 0:
        orig = U.clone()
        fluid.addViscosity(dt, orig, flags, viscosity)

    if 'density' in batch_dict:
        density = batch_dict['density']
        density = fluid.advectScalar(dt, density, U, flags, method="maccormackFluidNet", boundary_width=1,
                                     sample_outside_fluid=sampleOutsideFluid,
                                     maccormack_strength=maccormackStrength)
        if mconf.get('correctScalar', False):
            div = fluid.velocityDivergence(U, flags)
            fluid.correctScalar(dt, density, div, flags)
    else:
        density = torch.zeros_like(flags)

    if viscosity == 0:
        U = fluid.advectVelocity(dt=dt, orig=U, U=U, flags=flags, method="maccormackFluidNet",
                                 boundary_width=1, maccormack_strength=maccormackStrength)
    else:
        U = fluid.advectVelocity(dt=dt, orig=orig, U=U, flags=flags, method="maccormackFluidNet",
                                 boundary_width=1, maccormack_strength=maccormackStrength)

    setConstVals(batch_dict, p, U, flags, density)

    if 'density' in batch_dict:
        if buoyancyScale > 0:
            U = fluid.addBuoyancy(U, flags, density, _gravity(mconf, buoyancyScale), mconf['operatingDensity'], dt)
        if gravityScale > 0:
            U = fluid.addGravity(U, flags, _gravity(mconf, gravityScale), dt)

    if output_div:
        return

    if sim_method != 'convnet':
        U = _wall_bcs_with_seam(U, flags, mconf)
    elif stick:
        fluid.setWallBcsStick(U, flags, flags_stick)

    setConstVals(batch_dict, p, U, flags, density)

    if sim_method == 'convnet':
        net.eval()
        data = torch.cat((p, U, flags, density), 1)
        p, U = net(data)
    else:
        div = fluid.velocityDivergence(U, flags)
        is3D = (U.size(2) > 1)
        p, residual = fluid.solveLinearSystemJacobi(flags=flags, div=div, is_3d=is3D, p_tol=mconf['pTol'],
                                                    max_iter=mconf['jacobiIter'])
        fluid.velocityUpdate(pressure=p, U=U, flags=flags)

    if sim_method != 'convnet':
        U = _wall_bcs_with_seam(U, flags, mconf)
    elif stick:
        fluid.setWallBcsStick(U, flags, flags_stick)

    setConstVals(batch_dict, p, U, flags, density)
    batch_dict['U'] = U
    batch_dict['density'] = density
    batch_dict['p'] = p
