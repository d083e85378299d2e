"""The reference's UNMODIFIED drivers (pytorch/plume.py, pytorch/rayleighTaylor.py -- verbatim copies staged by
oracle/build_ref.py under oracle/_ref/drivers/) run on the GPU against this library through
tools/run_reference_driver.py, and what they write (restart.pth; growth.npy / avg_density.npy) is compared with
the same simulation stepped on CPU by the reference's own lib (oracle/_ref/reflib):
  Jacobi projection: bit-exact;  ScaleNet projection: 5e-5 relative (max norm) after the run's steps.
The drivers instantiate the model class saved with the weights; lib.simulate routes it to the fused, graph-replayed
path (simulate._native_net), which the timing line at the end of each test reports."""
import os
import subprocess
import sys
import time

import numpy as np
import pytest
import torch
import yaml

from conftest import ROOT

pytestmark = pytest.mark.gpu
DRV = os.path.join(ROOT, "oracle", "_ref", "drivers")
MODEL_DIR = os.path.join(DRV, "trained_models", "ScaleNet_ShortTerm_LongTermLoss")


def _need():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    if not os.path.isdir(DRV):
        pytest.skip("oracle/_ref/drivers (staged copies of the reference drivers) not present")
    import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref not built")


def _run_driver(script, cfg_path, extra=()):
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_driver.py"),
                        os.path.join(DRV, "pytorch", script), "--simConf", str(cfg_path)] + list(extra),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    return time.perf_counter() - t0


def _rel(got, ref):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    return float(np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-30))


@pytest.mark.parametrize("method", ["jacobi", "convnet"])
def test_plume_driver_unmodified(tmp_path, method):
    _need()
    import ref_loader
    with open(os.path.join(DRV, "pytorch", "plumeConfig.yaml")) as f:
        conf = yaml.safe_load(f)
    steps = 10
    conf.update({"outputFolder": str(tmp_path / "out"), "modelDir": MODEL_DIR, "realTimePlot": False, "saveVTK": False,
                 "maxIter": steps, "statIter": steps - 1, "simMethod": method, "jacobiIter": 28, "resX": 128, "resY": 128})
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text(yaml.safe_dump(conf))
    wall = _run_driver("plume.py", cfg)
    got = torch.load(tmp_path / "out" / "restart.pth", map_location="cpu", weights_only=False)
    assert got["it"] == steps - 1
    # the same run on CPU with the reference's own lib
    reflib, net, mconf_model = ref_loader.load_scalenet()
    mconf = dict(mconf_model); mconf.update(conf)
    net.mconf = mconf; net.scale.mconf = mconf
    bd = {k: torch.zeros(1, c, 1, 128, 128) for k, c in (("p", 1), ("U", 2), ("flags", 1), ("density", 1))}
    reflib.fluid.emptyDomain(bd["flags"])
    reflib.fluid.createPlumeBCs(bd, conf["injectionDensity"], conf["injectionVelocity"], conf["sourceRadius"])
    with torch.no_grad():
        for _ in range(steps):
            reflib.simulate(mconf, bd, net, method)
    for k in ("p", "U", "density", "flags"):
        a, b = got["batch_dict"][k].numpy(), bd[k].numpy()
        if method == "jacobi" or k == "flags":
            assert int(np.sum(~((a == b) | (np.isnan(a) & np.isnan(b))))) == 0, (method, k)
        else:
            assert _rel(a, b) < 5e-5, (method, k, _rel(a, b))
    print(f"[drivers] plume.py unmodified, 128x128 {method}, {steps} steps incl. start-up: {wall:.1f} s wall")


def test_rayleigh_taylor_driver_unmodified(tmp_path):
    _need()
    import ref_loader
    with open(os.path.join(DRV, "pytorch", "rayleighTaylorConfig.yaml")) as f:
        conf = yaml.safe_load(f)
    steps, resX, resY = 8, 64, 128
    conf.update({"outputFolder": str(tmp_path / "out"), "modelDir": MODEL_DIR, "realTimePlot": False, "saveVTK": False,
                 "maxIter": steps, "statIter": 2, "simMethod": "convnet", "resX": resX, "resY": resY})
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text(yaml.safe_dump(conf))
    train = tmp_path / "train.yaml"       # rayleighTaylor.py:64,105-107 only splits it into (conf, mconf)
    train.write_text(yaml.safe_dump({"dataDir": "/nonexistent", "dataset": "none", "modelDir": MODEL_DIR,
                                     "modelFilename": "convModel", "modelParam": {"dt": 0.1}}))
    wall = _run_driver("rayleighTaylor.py", cfg, ["--trainingConf", str(train)])
    growth = np.load(tmp_path / "out" / "growth.npy", allow_pickle=True)
    avg = np.load(tmp_path / "out" / "avg_density.npy", allow_pickle=True)
    # the same run on CPU with the reference's own lib (rayleighTaylor.py:140-260)
    reflib, net, mconf_model = ref_loader.load_scalenet()
    mconf = {"dt": 0.1}; mconf.update(mconf_model); mconf.update(conf)
    mconf["periodic-y"] = True; mconf["periodic-x"] = False
    net.mconf = mconf; net.scale.mconf = mconf
    bd = {k: torch.zeros(1, c, 1, resY, resX) for k, c in (("p", 1), ("U", 2), ("flags", 1), ("density", 1))}
    reflib.fluid.emptyDomain(bd["flags"])
    reflib.fluid.createRayleighTaylorBCs(bd, mconf, rho1=mconf["rho1"], rho2=mconf["rho2"])
    means = []
    with torch.no_grad():
        for it in range(steps):
            reflib.simulate(mconf, bd, net, "convnet")
            if it % 2 == 0:
                means.append([it, float(torch.mean(bd["density"]))])
    got_means = np.asarray(avg, np.float64).reshape(-1, 2)[-len(means):]
    assert got_means.shape == (len(means), 2)
    assert np.allclose(got_means[:, 0], [m[0] for m in means])
    ref_m = np.array([m[1] for m in means])
    assert np.max(np.abs(got_means[:, 1] - ref_m)) <= 5e-5 * max(np.max(np.abs(ref_m)), 1e-30) + 1e-9
    assert len(growth) >= steps
    print(f"[drivers] rayleighTaylor.py unmodified, {resY}x{resX} convnet, {steps} steps incl. start-up: {wall:.1f} s wall")
