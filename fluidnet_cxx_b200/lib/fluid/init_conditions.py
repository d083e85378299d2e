"""Initial / boundary-condition masks of the two reference drivers.  Setup code, runs once per
simulation, so it stays plain torch on the state tensors' own device.

Behaviour follows pytorch/lib/fluid/init_conditions.py:4-83 (plume inlet: rows 0:4 of the
domain, inside |x - W//2| <= floor(W*rad) velocity (0,u_scale) and density forced, rest of those
rows velocity forced to 0) and :88-130 (Rayleigh-Taylor tanh density interface)."""
import math

import torch


def createPlumeBCs(batch_dict, density_val, u_scale, rad):
    assert len(batch_dict) == 4, "Batch must contain 4 tensors (p, UDiv, flags, density)"
    U = batch_dict['U']
    density = batch_dict['density']
    assert U.dim() == 5, 'UBC must have 5 dimensions'
    assert U.size(0) == 1, 'Only single batches allowed (inference)'
    xdim, zdim = U.size(4), U.size(2)
    is3d = U.size(1) == 3
    if not is3d:
        assert zdim == 1, 'For 2D, zdim must be 1'
    dev = U.device
    UBC = torch.zeros_like(U)
    UBCInvMask = torch.ones_like(U)
    densityBC = torch.zeros_like(density)
    densityBCInvMask = torch.ones_like(density)

    centre = xdim // 2
    plume_rad = math.floor(xdim * rad)
    ix = torch.arange(xdim, device=dev) - centre
    inside = (ix.pow(2) <= plume_rad * plume_rad)            # (W,)
    inside_f = inside.to(U.dtype)
    rows = slice(0, 4)                                        # inlet = first 4 rows (y)
    # velocity: (0, u_scale[, 0]) inside the inlet, 0 outside; both overwrite the solver value
    UBC[:, 1, :, rows, :] = inside_f * float(u_scale)
    UBCInvMask[:, :, :, rows, :] = 0
    # density: forced inside the inlet, left alone outside
    densityBC[:, :, :, rows, :] = inside_f * float(density_val)
    densityBCInvMask[:, :, :, rows, :] = 1 - inside_f

    batch_dict['UBC'] = UBC
    batch_dict['UBCInvMask'] = UBCInvMask
    batch_dict['densityBC'] = densityBC
    batch_dict['densityBCInvMask'] = densityBCInvMask


def createRayleighTaylorBCs(batch_dict, mconf, rho1, rho2):
    assert len(batch_dict) == 4, "Batch must contain 4 tensors (p, UDiv, flags, density)"
    U = batch_dict['U']
    flags = batch_dict['flags']
    resX, resY = U.size(4), U.size(3)
    dev = U.device
    X = torch.arange(0, resX, device=dev).view(1, resX).expand(resY, resX)
    Y = torch.arange(0, resY, device=dev).view(resY, 1).expand(resY, resX)
    thick = mconf['perturbThickness']
    ampl = mconf['perturbAmplitude']
    h = mconf['height']
    # integer tensors divided by python ints: true division in fp32, as in the reference
    density = 0.5 * (rho2 + rho1 + (rho2 - rho1) * torch.tanh(
        thick * (Y / resY - (h + ampl * torch.cos(2 * math.pi * (X / resX))))))
    batch_dict['density'] = density.to(U.dtype).view(1, 1, 1, resY, resX).contiguous()
    batch_dict['flags'] = flags
