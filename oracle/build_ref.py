#!/usr/bin/env python
"""Build the UNMODIFIED-ARITHMETIC reference (jolibrain/fluidnet_cxx) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product
package (fluidnet_cxx_b200/); only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may execute it.

What it does (needs /root/reference, i.e. runs in the build container only):
  1. copies the reference's `pytorch/lib/fluid` package (Python + the C++ ATen
     extension sources), `lib/multi_scale_net.py`, and the shipped ScaleNet
     (`trained_models/ScaleNet_ShortTerm_LongTermLoss/*_saved.py` + mconf)
     into oracle/_ref/  (git-ignored: reference sources never enter history);
  2. applies the mechanical torch-2.x compatibility patch of SURVEY.md §8(c)
     (uint8 masks -> bool, `max_values` -> `amax`, `1 - mask` -> logical_not)
     and the CPU-device patch (`torch.device('cuda')` -> 'cpu', drop `.cuda()`).
     No arithmetic is changed;
  3. compiles the `fluidnet_cpp` pybind extension with torch's cpp_extension
     into oracle/_ref/build/fluidnet_cpp.so.

oracle/_ref/ travels to the GPU box with the gpurun snapshot (it is git-ignored
but not gpurun-ignored) so `bench.py --impl reference` can time the reference's
own CPU path there.
"""
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FLUIDNET_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
MODEL_DIR = "trained_models/ScaleNet_ShortTerm_LongTermLoss"


def _sub(path, pairs, count_min=1):
    with open(path) as f:
        s = f.read()
    for pat, rep in pairs:
        s2, n = re.subn(pat, rep, s)
        if n < count_min:
            raise RuntimeError(f"patch {pat!r} did not apply to {path}")
        s = s2
    with open(path, "w") as f:
        f.write(s)


def stage_sources():
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    # staged under the package name `reflib` so it never collides with the
    # product's own top-level `lib` package (the reference-facing surface)
    lib = os.path.join(OUT, "reflib")
    os.makedirs(lib)
    shutil.copytree(os.path.join(REF, "pytorch/lib/fluid"), os.path.join(lib, "fluid"))
    for f in ("multi_scale_net.py", "simulate.py"):
        shutil.copy(os.path.join(REF, "pytorch/lib", f), lib)
    shutil.copy(os.path.join(REF, MODEL_DIR, "ScaleNet_ShortTerm_LongTermLoss_saved.py"),
                os.path.join(lib, "model_saved.py"))
    # lib/__init__.py of the reference also imports matplotlib-based modules that
    # are not installed in this image; export only what the step needs.
    with open(os.path.join(lib, "__init__.py"), "w") as f:
        f.write("from .multi_scale_net import MultiScaleNet\n"
                "from . import fluid\n"
                "from .simulate import simulate, setConstVals\n")
    for f in ("convModel_mconf.pth", "convModel_conf.pth", "convModel_lastEpoch_best.pth"):
        shutil.copy(os.path.join(REF, MODEL_DIR, f), OUT)
    for f in ("plumeConfig.yaml", "rayleighTaylorConfig.yaml"):
        shutil.copy(os.path.join(REF, "pytorch", f), OUT)

    cpp = os.path.join(lib, "fluid", "cpp")
    # --- torch 2.x compat; no arithmetic changes -------------
    _sub(os.path.join(cpp, "calc_line_trace.cpp"), [
        (r"maxT\.max_values\(1, true\)", "maxT.amax(1, true)"),
        (r"at::kByte", "at::kBool"),
    ])
    _sub(os.path.join(cpp, "fluids_init.cpp"), [
        (r"at::kByte", "at::kBool"),
        (r"maskSolid\.equal\(1-maskFluid\)", "maskSolid.equal(maskFluid.logical_not())"),
    ])
    _sub(os.path.join(cpp, "grid.cpp"), [
        (r"T m3 = 1 - \(m0\.__or__\(m1\)\.__or__\(m2\)\);",
         "T m3 = (m0.__or__(m1).__or__(m2)).logical_not();"),
    ])
    fl = os.path.join(lib, "fluid")
    for f in ("set_wall_bcs.py", "source_terms.py", "set_wall_bcs_stick.py",
              "set_wall_bcs_inflow.py", "source_terms_test.py"):
        _sub(os.path.join(fl, f), [(r"dtype=torch\.uint8", "dtype=torch.bool")])
    # --- CPU device patch -----------------------------------------------------
    for root, _, files in os.walk(lib):
        for f in files:
            if not f.endswith(".py"):
                continue
            p = os.path.join(root, f)
            with open(p) as fh:
                s = fh.read()
            s2 = s.replace("torch.device('cuda')", "torch.device('cpu')").replace(".cuda()", "")
            if s2 != s:
                with open(p, "w") as fh:
                    fh.write(s2)
    # make the package name-independent (it is imported as `reflib`)
    _sub(os.path.join(lib, "simulate.py"), [(r"import lib\.fluid as fluid", "from . import fluid")])
    _sub(os.path.join(lib, "model_saved.py"),
         [(r"from lib import fluid, MultiScaleNet",
           "from . import fluid\nfrom .multi_scale_net import MultiScaleNet")])
    # the python wrappers `import fluidnet_cpp` (top-level module name, build/)


def stage_drivers():
    """UNMODIFIED copies of the reference's two simulation drivers, their YAML configs and the shipped
    model directory under oracle/_ref/drivers/ (git-ignored like the rest of oracle/_ref; it travels to the
    GPU box so tests/test_gpu_drivers.py can run `pytorch/plume.py` / `pytorch/rayleighTaylor.py` as they
    are against the B200 library through tools/run_reference_driver.py)."""
    d = os.path.join(OUT, "drivers")
    if os.path.isdir(d):
        shutil.rmtree(d)
    os.makedirs(os.path.join(d, "pytorch"))
    for f in ("plume.py", "rayleighTaylor.py", "plumeConfig.yaml", "rayleighTaylorConfig.yaml", "trainConfig.yaml"):
        shutil.copy(os.path.join(REF, "pytorch", f), os.path.join(d, "pytorch"))
    md = os.path.join(d, MODEL_DIR)
    os.makedirs(md)
    for f in ("ScaleNet_ShortTerm_LongTermLoss_saved.py", "convModel_mconf.pth", "convModel_conf.pth",
              "convModel_lastEpoch_best.pth"):
        shutil.copy(os.path.join(REF, MODEL_DIR, f), md)


def build_extension():
    import torch  # noqa: F401
    from torch.utils.cpp_extension import load
    cpp = os.path.join(OUT, "reflib", "fluid", "cpp")
    bdir = os.path.join(OUT, "build")
    os.makedirs(bdir, exist_ok=True)
    srcs = [os.path.join(cpp, f) for f in
            ("grid.cpp", "advect_type.cpp", "calc_line_trace.cpp", "fluids_init.cpp")]
    load(name="fluidnet_cpp", sources=srcs, build_directory=bdir,
         extra_cflags=["-O2", "-w"], verbose=False)
    assert os.path.exists(os.path.join(bdir, "fluidnet_cpp.so"))


def main():
    if not os.path.isdir(REF):
        print(f"[oracle/build_ref] {REF} not present; keeping prebuilt oracle/_ref as is")
        return 0
    if "--drivers-only" in sys.argv:
        stage_drivers()
        print("[oracle/build_ref] staged", os.path.join(OUT, "drivers"))
        return 0
    stage_sources()
    stage_drivers()
    build_extension()
    print("[oracle/build_ref] built", os.path.join(OUT, "build", "fluidnet_cpp.so"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
