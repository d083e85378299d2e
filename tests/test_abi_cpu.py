"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/fluidstep.h declares; argument validation fails loudly; no compute calls."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from fluidnet_cxx_b200 import build, _native
    build.build()
    return _native.load()


def declared_symbols():
    with open(os.path.join(ROOT, "include", "fluidstep.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fnx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fluidstep.h but not exported"


def test_binding_covers_header(lib):
    from fluidnet_cxx_b200 import _native
    bound = set(_native.SIGNATURES) | set(_native.OPTIONAL_SIGNATURES)
    assert set(declared_symbols()) <= bound


def test_build_info(lib):
    assert b"sm_100a" in lib.fnx_build_info()
    assert lib.fnx_abi_version() >= 1


def test_workspace_sizes(lib):
    assert lib.fnx_advect_scalar_workspace(1, 1, 128, 128) == 128 * 128 * 8
    assert lib.fnx_advect_vel_workspace(1, 1, 128, 128, 0) == 128 * 128 * 8
    assert lib.fnx_advect_vel_workspace(1, 16, 16, 16, 1) == 16 ** 3 * 12
    assert lib.fnx_jacobi_workspace(2, 1, 64, 64, 10) >= 2 * 64 * 64 * 4
    assert lib.fnx_step_workspace(1, 1, 64, 64, 0) > 5 * 64 * 64 * 4


def test_argument_errors(lib):
    # shape validation happens before any CUDA call, so it is testable without a GPU
    assert lib.fnx_velocity_divergence(None, None, None, 1, 4, 8, 8, 0, None) == -1   # 2-D with D > 1
    assert b"unsupported grid" in lib.fnx_last_error()
    assert lib.fnx_advect_scalar(0.1, None, None, None, None, 1, 1, 8, 8, 0, 7, 1, 0, 0.5, None, 0, None) == -1
    assert lib.fnx_advect_scalar(0.1, None, None, None, None, 1, 1, 8, 8, 0, 1, 2, 0, 0.5, None, 0, None) == -1
    it = ctypes.c_int(0)
    assert lib.fnx_solve_linear_system_jacobi(None, None, None, None, 1, 1, 8, 8, 0, 0.0, 0, ctypes.byref(it),
                                              None, 0, None) == -1
    assert b"At least 1 iteration" in lib.fnx_last_error()


def test_no_cpu_fallback():
    """CPU tensors are rejected loudly (the product never routes through the oracle or torch ops)."""
    from fluidnet_cxx_b200.lib import fluid
    U = torch.zeros(1, 2, 1, 8, 8)
    flags = torch.ones(1, 1, 1, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        fluid.velocityDivergence(U, flags)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        fluid.advectScalar(0.1, flags.clone(), U, flags)


def test_reference_asserts():
    """Same assertion behaviour as the reference wrappers (advection.py:44-62, velocity_divergence.py:18-36)."""
    from fluidnet_cxx_b200.lib import fluid
    U = torch.zeros(1, 2, 1, 8, 8)
    flags = torch.ones(1, 1, 1, 8, 8)
    with pytest.raises(AssertionError, match="Dimension mismatch"):
        fluid.velocityDivergence(U[0], flags)
    with pytest.raises(AssertionError, match="flags is not scalar"):
        fluid.velocityDivergence(U, U)
    with pytest.raises(AssertionError, match="Size mismatch"):
        fluid.velocityDivergence(torch.zeros(1, 2, 1, 8, 9), flags)
    with pytest.raises(AssertionError, match="Advection method not supported"):
        fluid.advectScalar(0.1, flags.clone(), U, flags, method="rk4")
    with pytest.raises(AssertionError, match="Input is not contiguous"):
        fluid.setWallBcs(torch.zeros(1, 2, 1, 8, 16)[..., ::2], flags)


def test_product_does_not_import_oracle():
    """Nothing under fluidnet_cxx_b200/ may import, link or execute oracle/."""
    pkg = os.path.join(ROOT, "fluidnet_cxx_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                with open(os.path.join(root, f)) as fh:
                    s = fh.read()
                assert "import oracle" not in s and "oracle/" not in s and "ref_loader" not in s and \
                    "fluid_oracle" not in s, os.path.join(root, f)


def test_msnet_workspace_is_host_only(lib):
    """fnx_msnet_workspace is pure host arithmetic over the plan (dry run of the forward's layer
    schedule): usable without a GPU, grows with the grid, rejects grids with an empty 1/4 scale."""
    from fluidnet_cxx_b200 import _native
    plan = _native.MsnetPlan()
    plan.data_channels = 2
    spec = {"quarter": ([2, 32, 64, 32, 1], [3, 3, 3, 3], 2), "half": ([3, 32, 64, 128, 64, 32, 1], [5, 3, 3, 3, 3, 3], 4),
            "full": ([3, 32, 64, 128, 64, 32, 8], [5, 3, 3, 3, 3, 5], 4)}
    for name, (ch, ks, nrelu) in spec.items():
        arr = getattr(plan, name)
        for i, k in enumerate(ks):
            arr[i].cin, arr[i].cout, arr[i].ksize, arr[i].relu = ch[i], ch[i + 1], k, int(i < nrelu)
            arr[i].w_tc = 1 if (k == 3 and ch[i] % 16 == 0 and ch[i + 1] in (32, 64, 128)) else None
    plan.final_conv.cin, plan.final_conv.cout, plan.final_conv.ksize = 8, 1, 1
    small = lib.fnx_msnet_workspace(ctypes.byref(plan), 64, 64)
    big = lib.fnx_msnet_workspace(ctypes.byref(plan), 512, 512)
    assert 0 < small < big
    # 512^2: split activations 32+64+128+64 channels x 4 B x (516^2) at full res dominate
    assert big > (32 + 64 + 128 + 64) * 4 * 516 * 516
    assert lib.fnx_msnet_workspace(ctypes.byref(plan), 3, 3) == 0


def test_manta_file_roundtrip_and_reference_reader(tmp_path):
    """lib.load_manta_data.loadMantaFile reads the Mantaflow .bin layout exactly like the reference's
    struct.unpack reader (pytorch/lib/load_manta_data.py:4-41, re-stated here byte by byte)."""
    import struct
    import numpy as np
    from fluidnet_cxx_b200.lib.load_manta_data import loadMantaFile, saveMantaFile
    for is3d, (nz, ny, nx) in ((False, (1, 7, 5)), (True, (3, 4, 6))):
        rng = np.random.RandomState(nz)
        n = nz * ny * nx
        p = torch.from_numpy(rng.randn(1, 1, nz, ny, nx).astype(np.float32))
        U = torch.from_numpy(rng.randn(1, 3 if is3d else 2, nz, ny, nx).astype(np.float32))
        flags = torch.from_numpy(rng.choice([1, 2, 4], size=(1, 1, nz, ny, nx)).astype(np.float32))
        rho = torch.from_numpy(rng.rand(1, 1, nz, ny, nx).astype(np.float32))
        path = str(tmp_path / f"f{int(is3d)}.bin")
        saveMantaFile(path, p, U, flags, rho)
        # the reference reader, field by field
        with open(path, "rb") as f:
            head = struct.unpack("i" * 5, f.read(20))
            assert head[1:] == (nx, ny, nz, int(is3d))
            arr = struct.unpack("f" * 3 * n, f.read(12 * n))
            assert np.array_equal(np.float32(arr[:n]), U[0, 0].numpy().ravel())
            assert np.array_equal(np.float32(arr[2 * n:]), p.numpy().ravel())
        p2, U2, f2, r2, is3d2 = loadMantaFile(path)
        assert is3d2 == is3d and U2.shape == U.shape and f2.dtype == torch.float32
        assert torch.equal(p2, p) and torch.equal(U2, U) and torch.equal(f2, flags) and torch.equal(r2, rho)
    with open(str(tmp_path / "bad.bin"), "wb") as f:
        f.write(struct.pack("i" * 5, 0, 4, 4, 1, 0) + b"\0" * 10)
    with pytest.raises(AssertionError):
        loadMantaFile(str(tmp_path / "bad.bin"))


def test_reference_surface_and_driver_launcher(tmp_path):
    """`lib` exports every name of pytorch/lib/__init__.py:1-7; the headless launcher's shims let the
    unmodified plume.py get through imports, argparse, the YAML (yaml.load without Loader) and the output
    folder set-up -- on this CPU box it then stops at its first CUDA call, not at an import."""
    import subprocess
    import sys
    import fluidnet_cxx_b200.lib as L
    for name in ("FluidNetDataset", "summary", "MultiScaleNet", "FluidNet", "simulate", "plotField", "SmartFormatter",
                 "fluid"):
        assert hasattr(L, name), name
    ds = L.FluidNetDataset({"modelParam": {"dt": 0.1}, "dataDir": "/nonexistent", "dataset": "x"}, "te", save_dt=4)
    conf, mconf = ds.createConfDict()
    assert mconf == {"dt": 0.1} and "modelParam" not in conf and len(ds) == 0
    driver = "/root/reference/pytorch/plume.py"
    if not os.path.exists(driver):
        pytest.skip("reference checkout not present")
    cfg = tmp_path / "cfg.yaml"
    model_dir = "/root/reference/trained_models/ScaleNet_ShortTerm_LongTermLoss"
    import yaml
    with open("/root/reference/pytorch/plumeConfig.yaml") as f:
        conf = yaml.safe_load(f)
    conf.update({"outputFolder": f"{tmp_path}/out", "modelDir": model_dir, "realTimePlot": False, "saveVTK": False,
                 "maxIter": 2, "statIter": 1})
    cfg.write_text(yaml.safe_dump(conf))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_driver.py"), driver,
                        "--simConf", str(cfg)], capture_output=True, text=True, timeout=300)
    tail = (r.stderr or "")[-2000:]
    assert "ModuleNotFoundError" not in tail and "ImportError" not in tail, tail
    assert "load() missing" not in tail, tail                      # yaml.load shim in place
    if not torch.cuda.is_available():
        assert r.returncode != 0 and ("cuda" in tail.lower() or "nvidia" in tail.lower()), tail
        assert os.path.isdir(tmp_path / "out")                     # got past the config / folder set-up
