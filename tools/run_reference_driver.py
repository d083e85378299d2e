#!/usr/bin/env python
"""Run an UNMODIFIED driver of the reference (pytorch/plume.py, pytorch/rayleighTaylor.py) against the B200
library (SURVEY.md section 7 "hazards" H1-H4):

    python tools/run_reference_driver.py /path/to/fluidnet_cxx/pytorch/plume.py --simConf my.yaml ...

  H1  matplotlib / mpl_toolkits.axes_grid1.colorbar / pyevtk may be missing or too new: headless stand-ins are
      registered for whatever cannot be imported (set realTimePlot / saveVTK to false in the YAML);
  H2  `yaml.load(f)` without Loader (PyYAML >= 6) and `torch.load` of the `_mconf.pth` / `restart.pth` pickles
      (torch >= 2.6 defaults to weights_only) are shimmed;
  H3  `lib.FluidNetDataset` tolerates a missing dataset; `lib.simulate` accepts the legacy 5-argument call;
  H4  the driver copies `<modelDir>/<name>_saved.py` into ./lib/ and executes it: the launcher runs the driver
      from a scratch directory that has a ./lib/ folder; `import lib` resolves to fluidnet_cxx_b200.lib
      (compat.install), so the saved model's `from lib import fluid, MultiScaleNet` gets the tcgen05 network;
  H5  rayleighTaylor.py:112-115 updates `conf` with the pickled training configuration `<model>_conf.pth`, whose
      `modelDir` is the training machine's RELATIVE path (`data2/model_divL2_5divLT_ScaleNet_New` for the shipped
      model) and then loads the weights from there: the launcher links that relative path, inside the scratch
      directory the driver runs from, to the model directory given by --modelDir / the YAML;
  H6  rayleighTaylor.py:253-257 appends `[[it*dt, distance]]` -- a Python float next to a (n, 1) CUDA tensor -- to a
      NumPy array every step.  NumPy of the reference's day built an object array out of that; NumPy >= 1.24 refuses
      the ragged list and never converts CUDA tensors: `np.append` falls back to the old object-array result (tensor
      moved to the host) when NumPy rejects the arguments.
"""
import importlib
import os
import runpy
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class _Anything(types.ModuleType):
    """module stand-in: any attribute is a callable returning another stand-in"""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Callable(f"{self.__name__}.{name}")


class _Callable:
    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        return _Callable(self._name + "()")

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Callable(f"{self._name}.{name}")

    def __iter__(self):
        return iter(())


def _ensure(modname):
    try:
        importlib.import_module(modname)
        return False
    except Exception:       # noqa: BLE001 - missing or incompatible: stand in for it
        parts = modname.split(".")
        for i in range(1, len(parts) + 1):
            name = ".".join(parts[:i])
            if name not in sys.modules or i == len(parts):
                sys.modules[name] = _Anything(name)
            if i > 1:
                setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], sys.modules[name])
        return True


def prepare():
    """install the shims; returns the list of stubbed modules"""
    stubbed = [m for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.cm",
                           "mpl_toolkits.axes_grid1.axes_divider", "mpl_toolkits.axes_grid1.colorbar", "pyevtk.hl")
               if _ensure(m)]
    import yaml
    if not getattr(yaml.load, "_fnx_shim", False):
        _load = yaml.load

        def load(stream, Loader=None, **kw):
            return _load(stream, Loader=Loader or yaml.SafeLoader, **kw)
        load._fnx_shim = True
        yaml.load = load
    import torch
    if not getattr(torch.load, "_fnx_shim", False):
        _tload = torch.load

        def tload(*a, **kw):
            kw.setdefault("weights_only", False)
            return _tload(*a, **kw)
        tload._fnx_shim = True
        torch.load = tload
    import numpy as np
    if not getattr(np.append, "_fnx_shim", False):
        _append = np.append

        def append(arr, values, axis=None):
            try:
                return _append(arr, values, axis)
            except (ValueError, TypeError):
                rows = [[(v.detach().cpu() if isinstance(v, torch.Tensor) else v) for v in row] for row in values]
                obj = np.empty((len(rows), len(rows[0])), dtype=object)
                for i, row in enumerate(rows):
                    for j, v in enumerate(row):
                        obj[i, j] = v
                return _append(np.asarray(arr, dtype=object), obj, axis)
        append._fnx_shim = True
        np.append = append
    import fluidnet_cxx_b200.compat as compat
    compat.install()
    return stubbed


def _link_pickled_model_dir(argv, work):
    """hazard H5: make the relative `modelDir` pickled in <modelDir>/<modelFilename>_conf.pth resolve, from the
    scratch directory, to the real model directory"""
    import torch
    import yaml
    model_dir = argv[argv.index("--modelDir") + 1] if "--modelDir" in argv else None
    model_name = argv[argv.index("--modelFilename") + 1] if "--modelFilename" in argv else None
    if "--simConf" in argv:
        try:
            with open(argv[argv.index("--simConf") + 1]) as f:
                sc = yaml.safe_load(f) or {}
            model_dir = model_dir or sc.get("modelDir")
            model_name = model_name or sc.get("modelFilename")
        except (OSError, yaml.YAMLError):
            pass
    if not model_dir or not model_name:
        return
    cpath = os.path.join(model_dir, model_name + "_conf.pth")
    if not os.path.isfile(cpath):
        return
    try:
        pickled = torch.load(cpath, weights_only=False, map_location="cpu").get("modelDir")
    except Exception:       # noqa: BLE001 - nothing to link
        return
    if not pickled or os.path.isabs(pickled) or os.path.exists(os.path.join(work, pickled)):
        return
    link = os.path.join(work, pickled)
    os.makedirs(os.path.dirname(link) or work, exist_ok=True)
    os.symlink(os.path.abspath(model_dir), link)


def main():
    if len(sys.argv) < 2:
        print(__doc__)
        return 2
    driver = os.path.abspath(sys.argv[1])
    stubbed = prepare()
    if stubbed:
        print("[run_reference_driver] headless stand-ins for:", ", ".join(stubbed), file=sys.stderr)
    work = tempfile.mkdtemp(prefix="fnx_driver_")
    os.makedirs(os.path.join(work, "lib"), exist_ok=True)
    # relative paths on the command line (configs) refer to the caller's directory
    argv = [driver] + [os.path.abspath(a) if (not a.startswith("-") and os.path.exists(a)) else a for a in sys.argv[2:]]
    # the driver's default --simConf is relative to its own directory
    if "--simConf" not in argv:
        default = {"plume.py": "plumeConfig.yaml", "rayleighTaylor.py": "rayleighTaylorConfig.yaml"}.get(
            os.path.basename(driver))
        if default and os.path.exists(os.path.join(os.path.dirname(driver), default)):
            argv += ["--simConf", os.path.join(os.path.dirname(driver), default)]
    _link_pickled_model_dir(argv, work)
    os.chdir(work)
    sys.argv = argv
    runpy.run_path(driver, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
