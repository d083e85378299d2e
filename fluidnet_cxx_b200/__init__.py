"""fluidnet_cxx_b200 -- B200-native per-timestep fluid step behind the `lib.fluid` surface of
jolibrain/fluidnet_cxx.  Hand-written sm_100a CUDA (csrc/) behind a C-ABI
(include/fluidstep.h); `fluidnet_cxx_b200.lib` mirrors the reference's `lib` package
(pytorch/lib/__init__.py:1-7) for the hot path; `fluidnet_cxx_b200.fluidnet_cpp` mirrors its
pybind module (pytorch/lib/fluid/cpp/fluids_init.cpp:1009-1014)."""
__version__ = "0.1.0"
