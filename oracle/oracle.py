"""numpy front-end of the C oracle (oracle/fluid_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg, never by the product package.

Every function takes/returns float32 numpy arrays in the reference layout
(B, C, D, H, W) and mirrors the reference's `lib.fluid` call of the same name.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "_build", "libfluid_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)


def build(force=False):
    src = os.path.join(HERE, "fluid_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"], env={**os.environ, "CC": "gcc"})
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def trace_stats(reset=True):
    """Coverage counters of the line trace (see fluid_oracle.c: trace_stats)."""
    out = (ctypes.c_long * 8)()
    lib().orc_trace_stats(out, int(reset))
    names = ("traces", "border_exits", "blocked_hits", "box_misses", "box_inside", "unit_steps", "interp_fallbacks")
    return dict(zip(names, list(out)))


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def _dims(flags):
    B, C, D, H, W = flags.shape
    assert C == 1
    return B, D, H, W


_METHOD = {"eulerFluidNet": 0, "maccormackFluidNet": 1}


def advectScalar(dt, src, U, flags, method="maccormackFluidNet", boundary_width=1,
                 sample_outside_fluid=False, maccormack_strength=0.75):
    assert boundary_width == 1
    B, D, H, W = _dims(flags)
    is3d = int(U.shape[1] == 3)
    src, ps = _c(src); U, pu = _c(U); flags, pf = _c(flags)
    dst = np.empty_like(src)
    err = lib().orc_advect_scalar(ctypes.c_float(dt), ps, pu, pf, B, D, H, W, is3d, _METHOD[method],
                                  int(bool(sample_outside_fluid)), ctypes.c_float(maccormack_strength),
                                  dst.ctypes.data_as(_f32p))
    advectScalar.last_errors = err
    return dst


def advectVelocity(dt, orig, U, flags, method="maccormackFluidNet", boundary_width=1,
                   maccormack_strength=0.75):
    assert boundary_width == 1
    B, D, H, W = _dims(flags)
    is3d = int(U.shape[1] == 3)
    orig, po = _c(orig); U, pu = _c(U); flags, pf = _c(flags)
    dst = np.empty_like(U)
    lib().orc_advect_vel(ctypes.c_float(dt), po, pu, pf, B, D, H, W, is3d, _METHOD[method],
                         ctypes.c_float(maccormack_strength), dst.ctypes.data_as(_f32p))
    return dst


def addBuoyancy(U, flags, density, gravity, rho_star, dt):
    """Returns a new array (the reference mutates U in place and returns it)."""
    B, D, H, W = _dims(flags)
    is3d = int(U.shape[1] == 3)
    U = np.array(U, dtype=np.float32, order="C", copy=True)
    flags, pf = _c(flags); density, pd = _c(density)
    g = (ctypes.c_float * 3)(*[float(np.float32(x)) for x in gravity])
    lib().orc_add_buoyancy(U.ctypes.data_as(_f32p), pf, pd, g, ctypes.c_float(rho_star),
                           ctypes.c_float(dt), B, D, H, W, is3d)
    return U


def addGravity(U, flags, gravity, dt):
    B, D, H, W = _dims(flags)
    is3d = int(U.shape[1] == 3)
    U = np.array(U, dtype=np.float32, order="C", copy=True)
    flags, pf = _c(flags)
    g = (ctypes.c_float * 3)(*[float(np.float32(x)) for x in gravity])
    lib().orc_add_gravity(U.ctypes.data_as(_f32p), pf, g, ctypes.c_float(dt), B, D, H, W, is3d)
    return U


def addViscosity(dt, U, flags, viscosity):
    """viscosity.py:7-70 (2-D); returns the updated copy (the reference works in place)."""
    B, D, H, W = _dims(flags)
    assert D == 1 and U.shape[1] == 2, "addViscosity: the reference only works in 2-D"
    U = np.array(U, dtype=np.float32, order="C", copy=True)
    flags, pf = _c(flags)
    lib().orc_add_viscosity(U.ctypes.data_as(_f32p), pf, ctypes.c_double(dt), ctypes.c_double(viscosity), B, H, W)
    return U


def correctScalar(dt, src, div, flags):
    """advection.py:9-12; returns the updated copy (the reference works in place)."""
    src = np.array(src, dtype=np.float32, order="C", copy=True)
    div, pd = _c(div)
    flags, pf = _c(flags)
    lib().orc_correct_scalar(src.ctypes.data_as(_f32p), pd, pf, ctypes.c_double(dt), ctypes.c_size_t(src.size))
    return src


def setWallBcs(U, flags):
    B, D, H, W = _dims(flags)
    is3d = int(U.shape[1] == 3)
    U = np.array(U, dtype=np.float32, order="C", copy=True)
    flags, pf = _c(flags)
    lib().orc_set_wall_bcs(U.ctypes.data_as(_f32p), pf, B, D, H, W, is3d)
    return U


def velocityDivergence(U, flags):
    B, D, H, W = _dims(flags)
    is3d = int(U.shape[1] == 3)
    U, pu = _c(U); flags, pf = _c(flags)
    div = np.empty_like(flags)
    lib().orc_velocity_divergence(pu, pf, div.ctypes.data_as(_f32p), B, D, H, W, is3d)
    return div


def velocityUpdate(pressure, U, flags):
    B, D, H, W = _dims(flags)
    is3d = int(U.shape[1] == 3)
    U = np.array(U, dtype=np.float32, order="C", copy=True)
    pressure, pp = _c(pressure); flags, pf = _c(flags)
    lib().orc_velocity_update(pp, U.ctypes.data_as(_f32p), pf, B, D, H, W, is3d)
    return U


def solveLinearSystemJacobi(flags, div, is_3d=False, p_tol=1e-5, max_iter=1000):
    B, D, H, W = _dims(flags)
    flags, pf = _c(flags); div, pd = _c(div)
    p = np.zeros_like(flags)
    res = ctypes.c_float(0)
    lib().orc_jacobi(pf, pd, B, D, H, W, int(bool(is_3d)), ctypes.c_float(p_tol), int(max_iter),
                     p.ctypes.data_as(_f32p), ctypes.byref(res))
    return p, np.float32(res.value)


def getCentered(U):
    """grid.py:7-30: cell-centred velocity (B, 3, D, H, W); the last column / row / plane stay 0."""
    U = np.asarray(U, np.float32)
    B, C, D, H, W = U.shape
    out = np.zeros((B, 3, D, H, W), np.float32)
    h = np.float32(0.5)
    out[:, 0, :, :, :-1] = h * (U[:, 0, :, :, :-1] + U[:, 0, :, :, 1:])
    out[:, 1, :, :-1, :] = h * (U[:, 1, :, :-1, :] + U[:, 1, :, 1:, :])
    if D > 1:
        out[:, 2, :-1] = h * (U[:, 2, :-1] + U[:, 2, 1:])
    return out


OUTPUT_PLANES = ("div", "velx", "vely", "velz", "vel_norm", "gradRhox", "gradRhoy", "gradPx", "gradPy", "pressure")


def outputFields(U, flags, density, pressure, mask_obstacles=True):
    """The drivers' output block (pytorch/plume.py:238-263 and :330-423) as one function -> (B, 10, D, H, W):
    divergence, getCentered(U) x / y / z and its norm, the centred density and pressure gradients of the VTK block
    (2-D: getCentered of the face differences of the (H-2) x (W-2) interior, placed at [1, H-1) x [1, W-1); 0 in
    3-D), pressure; Obstacle cells of the velocity / norm / pressure planes NaN-filled when `mask_obstacles`
    (numpy masked_array .filled(nan) in the drivers).  Restates what fnx_output_fields computes per cell."""
    U = np.asarray(U, np.float32); flags = np.asarray(flags, np.float32)
    B, _, D, H, W = flags.shape
    out = np.zeros((B, len(OUTPUT_PLANES), D, H, W), np.float32)
    out[:, 0] = velocityDivergence(U, flags)[:, 0]
    c = getCentered(U)
    out[:, 1:4] = c
    out[:, 4] = np.sqrt((c[:, 0] * c[:, 0] + c[:, 1] * c[:, 1]) + c[:, 2] * c[:, 2], dtype=np.float32)
    h = np.float32(0.5)
    if D == 1:
        for first, fld in ((5, density), (7, pressure)):
            if fld is None:
                continue
            f = np.asarray(fld, np.float32)[:, 0]
            gx = np.zeros_like(f); gy = np.zeros_like(f)
            # x: cells i in [1, W-3], rows j in [1, H-2]; y: rows j in [1, H-3], cells i in [1, W-2]
            gx[:, :, 1:H - 1, 1:W - 2] = h * ((f[:, :, 1:H - 1, 1:W - 2] - f[:, :, 1:H - 1, 0:W - 3]) +
                                              (f[:, :, 1:H - 1, 2:W - 1] - f[:, :, 1:H - 1, 1:W - 2]))
            gy[:, :, 1:H - 2, 1:W - 1] = h * ((f[:, :, 1:H - 2, 1:W - 1] - f[:, :, 0:H - 3, 1:W - 1]) +
                                              (f[:, :, 2:H - 1, 1:W - 1] - f[:, :, 1:H - 2, 1:W - 1]))
            out[:, first], out[:, first + 1] = gx, gy
    if pressure is not None:
        out[:, 9] = np.asarray(pressure, np.float32)[:, 0]
    if mask_obstacles:
        ob = flags[:, 0] == 2
        for pl in (1, 2, 3, 4, 9):
            out[:, pl][ob] = np.nan
    return out


def setConstVals(x, inv_mask, bc):
    x = np.array(x, dtype=np.float32, order="C", copy=True)
    inv_mask, pm = _c(inv_mask); bc, pb = _c(bc)
    lib().orc_set_const_vals(x.ctypes.data_as(_f32p), pm, pb, ctypes.c_size_t(x.size))
    return x


def flagsToOccupancy(flags):
    flags, pf = _c(flags)
    occ = np.empty_like(flags)
    lib().orc_flags_to_occupancy(pf, occ.ctypes.data_as(_f32p), ctypes.c_size_t(flags.size))
    return occ


def conv2d(x, w, b, relu=False):
    N, Cin, H, W = x.shape
    Cout, Cin2, K, K2 = w.shape
    assert Cin == Cin2 and K == K2
    x, px = _c(x); w, pw = _c(w); b, pb = _c(b)
    y = np.empty((N, Cout, H, W), dtype=np.float32)
    lib().orc_conv2d(px, pw, pb, y.ctypes.data_as(_f32p), N, Cin, H, W, Cout, K, int(relu))
    return y


def resize_bilinear(x, Ho, Wo):
    N, C, H, W = x.shape
    x, px = _c(x)
    y = np.empty((N, C, Ho, Wo), dtype=np.float32)
    lib().orc_resize_bilinear(px, y.ctypes.data_as(_f32p), N * C, H, W, Ho, Wo)
    return y


def std_unbiased(x):
    B = x.shape[0]
    x, px = _c(x.reshape(B, -1))
    out = np.empty((B,), dtype=np.float32)
    lib().orc_std_unbiased(px, B, ctypes.c_size_t(x.shape[1]), out.ctypes.data_as(_f32p))
    return out


# ---------------------------------------------------------------------------
# Composite restatements (multi_scale_net.py:101-127, model *_saved.py:78-238,
# simulate.py:28-171) built from the per-op oracles above.
# ---------------------------------------------------------------------------
# (Sequential index of the conv inside `.encode`, relu-after?) per block
_BLOCKS = {
    "convN_4": [(0, True), (2, True), (4, False), (6, False)],
    "convN_2": [(0, True), (2, True), (4, True), (6, True), (8, False), (10, False)],
    "convN_1": [(0, True), (2, True), (4, True), (6, True), (8, False), (10, False)],
}


def multi_scale_net(x, weights):
    """x: (N, 2, H, W); weights: dict 'convN_4.encode.0.weight' -> array (state-dict names
    under `multiScale.`)."""
    N, C, H, W = x.shape

    def block(name, t):
        for idx, relu in _BLOCKS[name]:
            t = conv2d(t, weights[f"{name}.encode.{idx}.weight"], weights[f"{name}.encode.{idx}.bias"], relu)
        return t
    q = (int(H * 0.25), int(W * 0.25))
    h = (int(H * 0.5), int(W * 0.5))
    o4 = block("convN_4", resize_bilinear(x, *q))
    o2 = block("convN_2", np.concatenate((resize_bilinear(x, *h), resize_bilinear(o4, *h)), axis=1))
    o1 = block("convN_1", np.concatenate((x, resize_bilinear(o2, H, W)), axis=1))
    return conv2d(o1, weights["final.weight"], weights["final.bias"], False)


def fluidnet_forward(p, U, flags, weights, threshold=1e-5):
    """The shipped ScaleNet wrapper (inputChannels.div, normalizeInput on UDiv): returns (p, U)."""
    B = U.shape[0]
    div = velocityDivergence(U, flags)
    s = np.maximum(std_unbiased(U), np.float32(threshold)).reshape(B, 1, 1, 1, 1).astype(np.float32)
    Un = (U / s).astype(np.float32)
    x = np.concatenate(((div / s).astype(np.float32), flagsToOccupancy(flags)), axis=1)[:, :, 0]
    pn = multi_scale_net(x, weights)[:, :, None]
    Un = velocityUpdate(pn, Un, flags)
    return (pn * s).astype(np.float32), setWallBcs((Un * s).astype(np.float32), flags)
