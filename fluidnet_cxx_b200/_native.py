"""ctypes binding of libfluidstep_b200.so (the C-ABI declared in include/fluidstep.h).

There is no CPU or PyTorch fallback: if the library is missing or a tensor is
not a contiguous fp32 CUDA tensor the call raises.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libfluidstep_b200.so")

_c_void_p = ctypes.c_void_p
_c_int = ctypes.c_int
_c_float = ctypes.c_float
_c_size_t = ctypes.c_size_t


class StepParams(ctypes.Structure):
    """struct fnx_step_params (include/fluidstep.h)."""
    _fields_ = [("dt", _c_float), ("maccormack_strength", _c_float),
                ("sample_outside_fluid", _c_int),
                ("use_buoyancy", _c_int), ("use_gravity", _c_int),
                ("buoyancy3", _c_float * 3), ("gravity3", _c_float * 3),
                ("rho_star", _c_float), ("jacobi_iters", _c_int),
                ("apply_wall_bcs", _c_int), ("density_const_passes", _c_int),
                ("row_begin", _c_int), ("row_end", _c_int),
                ("held_row_begin", _c_int), ("held_row_end", _c_int)]


class ActMeta(ctypes.Structure):
    """struct fnx_act_meta (include/fluidstep.h)."""
    _fields_ = [("amax_bits", ctypes.c_uint), ("scale", _c_float)]


class ConvLayer(ctypes.Structure):
    """struct fnx_conv_layer (include/fluidstep.h)."""
    _fields_ = [("weight", _c_void_p), ("bias", _c_void_p), ("w_tc", _c_void_p),
                ("cin", _c_int), ("cout", _c_int), ("ksize", _c_int), ("relu", _c_int),
                ("w_scale", _c_float), ("w_norm", _c_float), ("b_max", _c_float), ("w_replicas", _c_int)]


HALO_MAX_PEERS, HALO_MAX_FIELDS = 8, 4


class HaloDesc(ctypes.Structure):
    """struct fnx_halo_desc (include/fluidstep.h)."""
    _fields_ = [("n_peers", _c_int), ("n_fields", _c_int),
                ("src", (_c_void_p * HALO_MAX_FIELDS) * HALO_MAX_PEERS),
                ("dst", (_c_void_p * HALO_MAX_FIELDS) * HALO_MAX_PEERS),
                ("count", _c_size_t * HALO_MAX_PEERS),
                ("flag_out", _c_void_p * HALO_MAX_PEERS), ("flag_in", _c_void_p * HALO_MAX_PEERS),
                ("epoch", _c_void_p), ("done", _c_void_p)]


class ProfileRec(ctypes.Structure):
    """struct fnx_profile_rec (include/fluidstep.h)."""
    _fields_ = [("cin", _c_int), ("cout", _c_int), ("ksize", _c_int), ("h", _c_int), ("w", _c_int),
                ("tensor", _c_int), ("ms", _c_float)]


class MsnetPlan(ctypes.Structure):
    """struct fnx_msnet_plan (include/fluidstep.h)."""
    _fields_ = [("data_channels", _c_int), ("quarter", ConvLayer * 4), ("half", ConvLayer * 6),
                ("full", ConvLayer * 6), ("final_conv", ConvLayer)]


# name -> (restype, argtypes); every symbol include/fluidstep.h declares
_P, _I, _F, _S = _c_void_p, _c_int, _c_float, _c_size_t
_GRID = [_I, _I, _I, _I, _I]  # B, D, H, W, is3d
SIGNATURES = {
    "fnx_last_error": (ctypes.c_char_p, []),
    "fnx_build_info": (ctypes.c_char_p, []),
    "fnx_abi_version": (_I, []),
    "fnx_launch_count": (ctypes.c_longlong, []),
    "fnx_advect_scalar_workspace": (_S, [_I, _I, _I, _I]),
    "fnx_advect_scalar": (_I, [_F, _P, _P, _P, _P] + _GRID + [_I, _I, _I, _F, _P, _S, _P]),
    "fnx_advect_vel_workspace": (_S, [_I, _I, _I, _I, _I]),
    "fnx_advect_vel": (_I, [_F, _P, _P, _P, _P] + _GRID + [_I, _I, _F, _P, _S, _P]),
    "fnx_jacobi_workspace": (_S, [_I, _I, _I, _I, _I]),
    "fnx_solve_linear_system_jacobi": (_I, [_P, _P, _P, _P] + _GRID + [_F, _I, ctypes.POINTER(_I), _P, _S, _P]),
    "fnx_jacobi_iterate": (_I, [_P, _P, _P, _P] + _GRID + [_I, _I, _I, _P, _S, _P]),
    "fnx_jacobi_iterate_resid": (_I, [_P, _P, _P, _P] + _GRID + [_I, _I, _I, _I, _I, _P, _P, _S, _P]),
    "fnx_step_project_bcs_rows": (_I, [_P] * 6 + [_I] + _GRID + [_I, _I, _P]),
    "fnx_jacobi_iterate_held": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _S, _P]),
    "fnx_jacobi_tilemask_bytes": (_S, [_I, _I, _I]),
    "fnx_jacobi_tilemask_held": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _S, _P]),
    "fnx_jacobi_iterate_held_masked": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _S, _P]),
    "fnx_halo_exchange": (_I, [ctypes.POINTER(HaloDesc), _P]),
    "fnx_step_project_bcs_held": (_I, [_P] * 6 + [_I] + _GRID + [_I, _I, _I, _I, _P]),
    "fnx_output_fields": (_I, [_P, _P, _P, _P, _P] + _GRID + [_I, _P]),
    "fnx_velocity_divergence": (_I, [_P, _P, _P] + _GRID + [_P]),
    "fnx_velocity_update": (_I, [_P, _P, _P] + _GRID + [_P]),
    "fnx_set_wall_bcs": (_I, [_P, _P] + _GRID + [_P]),
    "fnx_add_buoyancy": (_I, [_P, _P, _P, ctypes.POINTER(_F), _F, _F] + _GRID + [_P]),
    "fnx_add_gravity": (_I, [_P, _P, ctypes.POINTER(_F), _F] + _GRID + [_P]),
    "fnx_add_viscosity": (_I, [_P, _P, ctypes.c_double, ctypes.c_double, _I, _I, _I, _P, _S, _P]),
    "fnx_correct_scalar": (_I, [_P, _P, _P, ctypes.c_double, _S, _P]),
    "fnx_flags_to_occupancy": (_I, [_P, _P, _S, _P]),
    "fnx_set_const_vals": (_I, [_P, _P, _P, _S, _P]),
    "fnx_empty_domain": (_I, [_P] + _GRID + [_I, _P]),
    "fnx_get_centered": (_I, [_P, _P] + _GRID + [_P]),
    "fnx_step_workspace": (_S, _GRID),
    "fnx_mask_rows": (_I, [_P, _P, _P, _P, _P] + _GRID + [_P]),
    "fnx_step_advect_forces_div": (_I, [ctypes.POINTER(StepParams)] + [_P] * 11 + _GRID + [_P, _S, _P]),
    "fnx_step_project_bcs": (_I, [_P] * 6 + [_I] + _GRID + [_P]),
    "fnx_step_jacobi": (_I, [ctypes.POINTER(StepParams)] + [_P] * 12 + _GRID + [_P, _S, _P]),
    "fnx_scale_std_workspace": (_S, [_I]),
    "fnx_scale_std": (_I, [_P, _S, _I, _F, _P, _P, _S, _P]),
    "fnx_conv2d": (_I, [_P, _P, _P, _P] + [_I] * 9 + [_P]),
    "fnx_resize_bilinear": (_I, [_P, _P] + [_I] * 8 + [_P]),
    "fnx_tc_act_bytes": (_S, [_I, _I, _I]),
    "fnx_tc_weight_bytes": (_S, [_I, _I, _I]),
    "fnx_tc_pack_weights": (_I, [_P, _I, _I, _I, _F, _P, _P]),
    "fnx_tc_set_debug": (_I, [_P]),
    "fnx_tc_amax": (_I, [_P, _S, _P, _P]),
    "fnx_tc_pack_split": (_I, [_P, _I, _I, _I, _P, _P, _P, _P]),
    "fnx_tc_unpack_split": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "fnx_conv_tc": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _F, _F, _I, _P, _P, _I, _I, _P]),
    "fnx_msnet_workspace": (_S, [ctypes.POINTER(MsnetPlan), _I, _I]),
    "fnx_msnet_workspace_n": (_S, [ctypes.POINTER(MsnetPlan), _I, _I, _I]),
    "fnx_msnet_workspace_init": (_I, [_P, _S, _P]),
    "fnx_msnet_forward": (_I, [ctypes.POINTER(MsnetPlan), _P, _P, _I, _I, _I, _P, _S, _P]),
    "fnx_profile_enable": (_I, [_I]),
    "fnx_profile_fetch": (_I, [ctypes.POINTER(ProfileRec), _I]),
    "fnx_fluidnet_input": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "fnx_fluidnet_output": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "fnx_fluidnet_input_3d": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "fnx_fluidnet_output_3d": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
}

_lib = None
_lock = threading.Lock()


def load():
    """Load the shared library (once).  Raises if it is missing: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is not built. Run `python -m fluidnet_cxx_b200.build` (needs nvcc); "
                "fluidnet_cxx_b200 has no CPU/PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError = ABI mismatch, fail loudly
            fn.restype = res
            fn.argtypes = args
        for name in OPTIONAL_SIGNATURES:
            if hasattr(lib, name):
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = OPTIONAL_SIGNATURES[name]
        _lib = lib
    return _lib


OPTIONAL_SIGNATURES = {}


def check(err, who=""):
    if err != 0:
        msg = load().fnx_last_error().decode("utf-8", "replace")
        if err == -1:
            raise RuntimeError(f"{who}: {msg}" if who else msg)
        raise RuntimeError(f"{who}: fluidstep error {err}: {msg}")


def ptr(t):
    """Device pointer of a contiguous fp32 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("fluidnet_cxx_b200 runs on CUDA tensors only (no CPU fallback); got "
                           f"{type(t).__name__} on {getattr(t, 'device', '?')}")
    if t.dtype != torch.float32:
        raise RuntimeError(f"fluidnet_cxx_b200 expects float32 tensors, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError("fluidnet_cxx_b200 expects contiguous tensors")
    if t.device.index != torch.cuda.current_device():
        # the C-ABI launches on the current device / its current stream (the reference's ATen ops guard the
        # device themselves): refuse loudly instead of launching on the wrong GPU
        raise RuntimeError(f"tensor lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                           "wrap the call in `with torch.cuda.device(tensor.device):`")
    return t.data_ptr()


def stream_of(t):
    return torch.cuda.current_stream(t.device).cuda_stream


class _Workspaces:
    """Grow-only per-(device, tag) scratch buffers (nothing is allocated inside the C-ABI)."""

    def __init__(self):
        self._bufs = {}
        self._retired = []   # outgrown buffers stay alive: captured CUDA graphs may still point at them

    def get(self, device, tag, nbytes):
        # one buffer per (device, STREAM, tag): two simulations on different streams or threads never share
        # scratch (work on one stream is ordered, so re-use within a stream is safe; a CUDA-graph capture runs
        # on its own stream and therefore gets buffers of its own, which live as long as this registry)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = (idx, torch.cuda.current_stream(idx).cuda_stream, tag)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            if buf is not None:
                self._retired.append(buf)
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._bufs[key] = buf
        return buf

    def clear(self):
        self._bufs.clear()
        self._retired.clear()


workspaces = _Workspaces()


def grid_of(flags):
    """(B, D, H, W) of a 5-D (B, 1, D, H, W) tensor."""
    B, _, D, H, W = flags.shape
    return int(B), int(D), int(H), int(W)
