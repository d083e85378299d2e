"""Mirror of the reference's pybind module `fluidnet_cpp`
(pytorch/lib/fluid/cpp/fluids_init.cpp:1009-1014; signatures fluids_init.h:73-142):

    advect_scalar(dt, src, U, flags, method, bnd, sample_outside_fluid, maccormack_strength) -> Tensor
    advect_vel(dt, orig, U, flags, method, bnd, maccormack_strength) -> Tensor
    solve_linear_system(flags, div, is_3d, p_tol, max_iter, verbose) -> [p, residual]
"""
from .lib.fluid import ops as _ops


def advect_scalar(dt, src, U, flags, method, bnd, sample_outside_fluid, maccormack_strength):
    return _ops.advectScalar(dt, src, U, flags, method, bnd, sample_outside_fluid, maccormack_strength)


def advect_vel(dt, orig, U, flags, method, bnd, maccormack_strength):
    return _ops.advectVelocity(dt, orig, U, flags, method, bnd, maccormack_strength)


def solve_linear_system(flags, div, is_3d, p_tol, max_iter, verbose):
    p, residual = _ops.solveLinearSystemJacobi(flags, div, is_3d, p_tol, max_iter, verbose)
    return [p, residual]
