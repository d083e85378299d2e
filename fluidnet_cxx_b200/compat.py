"""Drop-in registration: makes `import lib`, `import lib.fluid as fluid`, `from lib import fluid,
MultiScaleNet` and `import fluidnet_cpp` resolve to this package, so the reference drivers
(pytorch/plume.py:18-21, *_saved.py:5) run against the B200 path without edits."""
import importlib
import sys


def install():
    import fluidnet_cxx_b200.lib as L
    sys.modules["lib"] = L
    for sub in ("fluid", "simulate", "multi_scale_net", "model", "dataset_load", "util_print", "plot_field",
                "argument_parser", "load_manta_data"):
        sys.modules["lib." + sub] = importlib.import_module("fluidnet_cxx_b200.lib." + sub)
    sys.modules["lib.fluid.cell_type"] = importlib.import_module("fluidnet_cxx_b200.lib.fluid.cell_type")
    sys.modules["fluidnet_cpp"] = importlib.import_module("fluidnet_cxx_b200.fluidnet_cpp")
    return L
