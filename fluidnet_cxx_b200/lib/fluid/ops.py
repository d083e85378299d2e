"""The `lib.fluid` operators, each a thin validated call into the sm_100a C-ABI.

Signatures, argument meaning, in-place/return behaviour and assertion messages follow the
reference Python surface (pytorch/lib/fluid/__init__.py:1-14):

  velocityDivergence  velocity_divergence.py:4-74     new tensor
  velocityUpdate      velocity_update.py:6-162        in place, returns None
  setWallBcs          set_wall_bcs.py:4-86            in place, returns U
  addBuoyancy         source_terms.py:6-116           in place, returns U
  addGravity          source_terms.py:122-219         in place, returns U
  flagsToOccupancy    flags_to_occupancy.py:6-19      new tensor
  emptyDomain         util.py:5-49                    in place
  getDx, getCentered  grid.py:3-30
  advectScalar / advectVelocity / correctScalar       cpp/advection.py:9-118
  solveLinearSystemJacobi                             cpp/solve_linear_sys.py:4-40

Unlike the reference (which asserts 3-D off, advection.py:58,108) 3-D fields are accepted: the
kernels implement the intended per-cell semantics with D=1 as the pinned 2-D special case.
"""
import ctypes

import torch

from ... import _native as N

_METHODS = {"eulerFluidNet": 0, "maccormackFluidNet": 1}


def _check_vel_flags(U, flags, *others):
    assert U.dim() == 5 and flags.dim() == 5 and all(o.dim() == 5 for o in others), "Dimension mismatch"
    assert flags.size(1) == 1, "flags is not scalar"
    bsz, d, h, w = flags.size(0), flags.size(2), flags.size(3), flags.size(4)
    is3d = U.size(1) == 3
    if not is3d:
        assert d == 1, "2D velocity field but zdepth > 1"
        assert U.size(1) == 2, "2D velocity field must have only 2 channels"
    assert U.size(0) == bsz and U.size(2) == d and U.size(3) == h and U.size(4) == w, "Size mismatch"
    assert U.is_contiguous() and flags.is_contiguous() and all(o.is_contiguous() for o in others), \
        "Input is not contiguous"
    for t in (U, flags) + others:
        N.ptr(t)    # CUDA + fp32 or raise: there is no CPU fallback
    return int(bsz), int(d), int(h), int(w), int(is3d)


def _gravity3(gravity):
    g = [float(x) for x in (gravity.tolist() if isinstance(gravity, torch.Tensor) else gravity)]
    assert len(g) == 3, "gravity must be a 3D vector (even in 2D)"
    # the reference holds gravity as an fp32 tensor: round each component to fp32 first
    return (ctypes.c_float * 3)(*g)


# ------------------------------------------------------------------------------------------
def getDx(self):
    grid_size_max = max(max(self.size(2), self.size(3)), self.size(4))
    return 1.0 / grid_size_max


def getCentered(self):
    """Cell-centred velocity (B,3,D,H,W); last row/column/plane stay 0 (grid.py:7-30)."""
    assert self.dim() == 5, "Dimension mismatch"
    B, C, D, H, W = (int(s) for s in self.shape)
    is3d = int(D > 1)
    assert C == (3 if is3d else 2), "Size mismatch"
    U = self.contiguous()
    out = torch.empty((B, 3, D, H, W), dtype=U.dtype, device=U.device)
    N.check(N.load().fnx_get_centered(N.ptr(U), N.ptr(out), B, D, H, W, is3d, N.stream_of(U)), "getCentered")
    return out


OUTPUT_PLANES = ("div", "velx", "vely", "velz", "vel_norm", "gradRhox", "gradRhoy", "gradPx", "gradPy", "pressure")


def outputFields(batch_dict, mask_obstacles=True, out=None):
    """The drivers' output step in one kernel (plume.py:238-263, 330-423): divergence, centred velocity and its
    norm, centred density / pressure gradients (2-D) and pressure, Obstacle cells of the velocity / norm /
    pressure planes NaN-filled.  Returns (out, views): out (B, 10, D, H, W) on the device and a dict
    name -> (B, D, H, W) view of it (names: OUTPUT_PLANES).  Not part of the reference's lib.fluid (which forms
    these with ~60 tensor ops and 9 host copies); `outputFieldsToHost` adds the single device-to-host copy."""
    U, flags = batch_dict['U'], batch_dict['flags']
    _check_vel_flags(U, flags)
    B, D, H, W = N.grid_of(flags)
    is3d = int(U.size(1) == 3)
    rho, p = batch_dict.get('density'), batch_dict.get('p')
    if out is None:
        out = torch.empty((B, len(OUTPUT_PLANES), D, H, W), dtype=torch.float32, device=U.device)
    assert out.is_contiguous() and tuple(out.shape) == (B, len(OUTPUT_PLANES), D, H, W) and out.device == U.device
    N.check(N.load().fnx_output_fields(N.ptr(U.contiguous()), N.ptr(flags.contiguous()),
                                       N.ptr(rho.contiguous()) if rho is not None else None,
                                       N.ptr(p.contiguous()) if p is not None else None, N.ptr(out), B, D, H, W, is3d,
                                       int(bool(mask_obstacles)), N.stream_of(U)), "outputFields")
    return out, {name: out[:, i] for i, name in enumerate(OUTPUT_PLANES)}


def outputFieldsToHost(batch_dict, mask_obstacles=True, pinned=None, device_out=None):
    """outputFields + ONE asynchronous copy into a pinned host buffer (allocated on first use: pass the returned
    `pinned` / `device_out` back in on the next output step).  Returns (arrays, pinned, device_out); `arrays` maps
    the plane names to numpy views of the pinned buffer, valid after the stream has been synchronised (done here)."""
    out, _ = outputFields(batch_dict, mask_obstacles, device_out)
    if pinned is None:
        pinned = torch.empty(out.shape, dtype=out.dtype).pin_memory()
    pinned.copy_(out, non_blocking=True)
    torch.cuda.current_stream(out.device).synchronize()
    host = pinned.numpy()
    return {name: host[:, i] for i, name in enumerate(OUTPUT_PLANES)}, pinned, out


def emptyDomain(flags, boundary_width=1):
    assert boundary_width > 0, 'Boundary width must be greater than zero!'
    assert flags.dim() == 5, 'Flags tensor should be 5D'
    assert flags.size(1) == 1, 'Flags should have only one channels (scalar field)'
    is3d = flags.size(2) > 1
    bw = boundary_width
    assert (((not is3d) or (flags.size(2) > bw * 2)) and (flags.size(3) > bw * 2) or (flags.size(4) > bw * 2)), \
        'Simulation domain is not big enough'
    assert flags.is_contiguous(), "Input is not contiguous"
    B, D, H, W = N.grid_of(flags)
    N.check(N.load().fnx_empty_domain(N.ptr(flags), B, D, H, W, int(is3d), int(bw), N.stream_of(flags)),
            "emptyDomain")


def flagsToOccupancy(flags):
    f = flags.contiguous()
    occ = torch.empty_like(f)
    N.check(N.load().fnx_flags_to_occupancy(N.ptr(f), N.ptr(occ), f.numel(), N.stream_of(f)), "flagsToOccupancy")
    return occ


def velocityDivergence(U, flags):
    B, D, H, W, is3d = _check_vel_flags(U, flags)
    div = torch.empty_like(flags)
    N.check(N.load().fnx_velocity_divergence(N.ptr(U), N.ptr(flags), N.ptr(div), B, D, H, W, is3d,
                                             N.stream_of(U)), "velocityDivergence")
    return div


def velocityUpdate(pressure, U, flags):
    B, D, H, W, is3d = _check_vel_flags(U, flags, pressure)
    assert pressure.is_same_size(flags), "Size mismatch"
    N.check(N.load().fnx_velocity_update(N.ptr(pressure), N.ptr(U), N.ptr(flags), B, D, H, W, is3d,
                                         N.stream_of(U)), "velocityUpdate")


def setWallBcs(U, flags):
    B, D, H, W, is3d = _check_vel_flags(U, flags)
    N.check(N.load().fnx_set_wall_bcs(N.ptr(U), N.ptr(flags), B, D, H, W, is3d, N.stream_of(U)), "setWallBcs")
    return U


def addBuoyancy(U, flags, density, gravity, rho_star, dt):
    B, D, H, W, is3d = _check_vel_flags(U, flags, density)
    assert density.is_same_size(flags), "Size mismatch"
    g = _gravity3(gravity)
    N.check(N.load().fnx_add_buoyancy(N.ptr(U), N.ptr(flags), N.ptr(density), g, float(rho_star), float(dt),
                                      B, D, H, W, is3d, N.stream_of(U)), "addBuoyancy")
    return U


def addGravity(U, flags, gravity, dt):
    B, D, H, W, is3d = _check_vel_flags(U, flags)
    g = _gravity3(gravity)
    N.check(N.load().fnx_add_gravity(N.ptr(U), N.ptr(flags), g, float(dt), B, D, H, W, is3d, N.stream_of(U)),
            "addGravity")
    return U


def setConstVals(x, inv_mask, bc):
    """x = x*inv_mask + bc in place (simulate.py:16-26)."""
    assert x.is_same_size(inv_mask) and x.is_same_size(bc), "Size mismatch"
    N.check(N.load().fnx_set_const_vals(N.ptr(x), N.ptr(inv_mask), N.ptr(bc), x.numel(), N.stream_of(x)),
            "setConstVals")
    return x


# ---- native-extension wrappers (cpp/advection.py, cpp/solve_linear_sys.py) ---------------------
def _check_advection_method(method):
    assert method == 'eulerFluidNet' or method == 'maccormackFluidNet', \
        'Error: Advection method not supported. Options are: maccormackFluidNet, eulerFluidNet'


def correctScalar(dt, src, div, flags):
    """src += dt*0.5*src*div on fluid cells, in place (advection.py:9-12; mconf['correctScalar'])."""
    assert src.is_same_size(div) and src.is_same_size(flags), "Size mismatch"
    N.check(N.load().fnx_correct_scalar(N.ptr(src), N.ptr(div), N.ptr(flags), float(dt), src.numel(), N.stream_of(src)),
            "correctScalar")


def addViscosity(dt, U, flags, viscosity):
    """Explicit viscous update of the velocity, in place (viscosity.py:7-70; 2-D, as the reference)."""
    assert U.dim() == 5 and flags.dim() == 5, "Dimension mismatch"
    assert flags.size(1) == 1, "flags is not scalar"
    b, d, h, w = (int(flags.size(i)) for i in (0, 2, 3, 4))
    is3D = (U.size(1) == 3)
    if not is3D:
        assert d == 1, "d > 1 for a 2D domain"
        assert U.size(4) == w, "2D velocity field must have only 2 channels"
    assert U.size(0) == b and U.size(2) == d and U.size(3) == h and U.size(4) == w, "size mismatch"
    assert U.is_contiguous() and flags.is_contiguous(), "Input is not contiguous"
    if is3D:
        raise NameError("name 'mask_fluid_k' is not defined")      # viscosity.py:50: the 3-D branch never ran
    lib = N.load()
    ws = N.workspaces.get(U.device, "viscosity", U.numel() * 4)
    N.check(lib.fnx_add_viscosity(N.ptr(U), N.ptr(flags), float(dt), float(viscosity), b, h, w, ws.data_ptr(),
                                  ws.numel(), N.stream_of(U)), "addViscosity")


def advectScalar(dt, src, U, flags, method='maccormackFluidNet', boundary_width=1,
                 sample_outside_fluid=False, maccormack_strength=0.75):
    _check_advection_method(method)
    assert src.dim() == 5 and U.dim() == 5 and flags.dim() == 5, "Dimension mismatch"
    B, D, H, W, is3d = _check_vel_flags(U, flags, src)
    assert src.is_same_size(flags), "Size mismatch"
    lib = N.load()
    dst = torch.empty_like(src)
    nbytes = lib.fnx_advect_scalar_workspace(B, D, H, W)
    ws = N.workspaces.get(src.device, "advect_scalar", nbytes)
    N.check(lib.fnx_advect_scalar(float(dt), N.ptr(src), N.ptr(U), N.ptr(flags), N.ptr(dst), B, D, H, W, is3d,
                                  _METHODS[method], int(boundary_width), int(bool(sample_outside_fluid)),
                                  float(maccormack_strength), ws.data_ptr(), ws.numel(), N.stream_of(src)),
            "advectScalar")
    return dst


def advectVelocity(dt, orig, U, flags, method='maccormackFluidNet', boundary_width=1,
                   maccormack_strength=0.75):
    _check_advection_method(method)
    B, D, H, W, is3d = _check_vel_flags(U, flags, orig)
    assert orig.is_same_size(U), "Size mismatch"
    lib = N.load()
    dst = torch.empty_like(U)
    nbytes = lib.fnx_advect_vel_workspace(B, D, H, W, is3d)
    ws = N.workspaces.get(U.device, "advect_vel", nbytes)
    N.check(lib.fnx_advect_vel(float(dt), N.ptr(orig), N.ptr(U), N.ptr(flags), N.ptr(dst), B, D, H, W, is3d,
                               _METHODS[method], int(boundary_width), float(maccormack_strength),
                               ws.data_ptr(), ws.numel(), N.stream_of(U)), "advectVelocity")
    return dst


def solveLinearSystemJacobi(flags, div, is_3d=False, p_tol=1e-5, max_iter=1000, verbose=False):
    assert div.dim() == 5 and flags.dim() == 5, "Dimension mismatch"
    assert flags.size(1) == 1, "flags is not scalar"
    assert div.is_same_size(flags), "Size mismatch"
    assert flags.is_contiguous() and div.is_contiguous(), "Input is not contiguous"
    B, D, H, W = N.grid_of(flags)
    lib = N.load()
    p = torch.empty_like(flags)
    residual = torch.empty((), dtype=torch.float32, device=flags.device)
    nbytes = lib.fnx_jacobi_workspace(B, D, H, W, int(max_iter))
    ws = N.workspaces.get(flags.device, "jacobi", nbytes)
    iters = ctypes.c_int(0)
    N.check(lib.fnx_solve_linear_system_jacobi(N.ptr(flags), N.ptr(div), N.ptr(p), residual.data_ptr(), B, D, H, W,
                                               int(bool(is_3d)), float(p_tol), int(max_iter), ctypes.byref(iters),
                                               ws.data_ptr(), ws.numel(), N.stream_of(flags)),
            "solveLinearSystemJacobi")
    if verbose:
        print(f"Jacobi: {iters.value} iterations, residual {residual.item():.3e}")
    return p, residual
