"""Import the patched reference build staged by oracle/build_ref.py.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): used by tests/, by
tools that generate tests/golden/*, and by bench.py's reference arm.
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF_DIR, "build", "fluidnet_cpp.so")) and \
        os.path.isdir(os.path.join(REF_DIR, "reflib"))


def load():
    """Return the reference package (`reflib`: .fluid, .simulate, .MultiScaleNet)."""
    if not available():
        raise RuntimeError("oracle/_ref is not built; run `python oracle/build_ref.py` "
                           "in the build container (needs /root/reference)")
    for p in (os.path.join(REF_DIR, "build"), REF_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch  # noqa: F401  (fluidnet_cpp.so links against libtorch)
    return importlib.import_module("reflib")


def load_scalenet():
    """The shipped ScaleNet (FluidNet wrapper + MultiScaleNet) with trained weights, CPU, eval()."""
    import torch
    reflib = load()
    model_saved = importlib.import_module("reflib.model_saved")
    mconf = torch.load(os.path.join(REF_DIR, "convModel_mconf.pth"), weights_only=False)
    state = torch.load(os.path.join(REF_DIR, "convModel_lastEpoch_best.pth"),
                       weights_only=False, map_location="cpu")
    net = model_saved.FluidNet(mconf, dropout=False)
    net.load_state_dict(state["state_dict"])
    net.eval()
    return reflib, net, mconf
