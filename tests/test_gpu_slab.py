"""The window-based slab step (fluidnet_cxx_b200.lib.slab) with the REAL kernels and the REAL halo-exchange
kernel (csrc/halo.cu):
  * on one GPU, N virtual ranks (own arena + stream each, peers = each other's device pointers) must
    reproduce the single-GPU fused step bit for bit on their owned rows, holding only their window;
  * on >= 2 GPUs (skipped otherwise), one process per GPU over symmetric memory, whole step in CUDA
    graphs, >= 200 steps at world 4 when 4 GPUs are there (tools/slab_check.py under torchrun)."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def make_state(H, W, seed, vscale=0.5):
    from fluidnet_cxx_b200.lib import fluid
    from test_gpu_parity import plume_mconf, random_case
    mconf = plume_mconf(jacobiIter=28)
    f, U, rho, p = random_case(seed, 1, H, W, "obstacle", 10, vscale, True)
    bd = {"p": torch.zeros(1, 1, 1, H, W, device="cuda"), "U": torch.zeros(1, 2, 1, H, W, device="cuda"),
          "flags": torch.zeros(1, 1, 1, H, W, device="cuda"), "density": torch.zeros(1, 1, 1, H, W, device="cuda")}
    fluid.emptyDomain(bd["flags"])
    fluid.createPlumeBCs(bd, mconf["injectionDensity"], mconf["injectionVelocity"], mconf["sourceRadius"])
    ff = torch.from_numpy(f).cuda()
    ff[:, :, :, 0:6] = bd["flags"][:, :, :, 0:6]          # keep the inlet rows free of random boxes
    bd["flags"] = ff.contiguous()
    bd["U"] = torch.from_numpy(U).cuda()
    bd["density"] = torch.from_numpy(rho).cuda()
    return mconf, bd


@pytest.mark.parametrize("world,H,W,K", [(2, 128, 96, 1), (4, 256, 132, 1), (3, 192, 64, 2), (1, 64, 64, 1)])
def test_virtual_ranks_equal_single_gpu(world, H, W, K):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    sim = importlib.import_module("fluidnet_cxx_b200.lib.simulate")
    from fluidnet_cxx_b200.lib import slab
    mconf, bd = make_state(H, W, seed=world * 10 + K)
    steps = 3
    ref = {k: v.clone() for k, v in bd.items()}
    sim.clear_graph_cache()
    refs = []
    for _ in range(steps):
        sim._simulate_fused(mconf, ref, None, "jacobi", float(mconf["dt"]), False)
        refs.append({k: ref[k].clone() for k in ("p", "U", "density")})
    vw = slab.VirtualWorld(world, torch.device("cuda", torch.cuda.current_device()))

    def held(name, a, b):
        return bd[name][:, :, :, a:b].contiguous() if name in bd else None
    steppers = [slab.SlabJacobiStep(t, mconf, H, W, held, K=K, use_graph=False) for t in vw.topos]
    # memory per rank: held rows only
    for s in steppers:
        assert s.U[0].shape[3] == s.g["rows_held"] <= H // world + 2 * s.g["G"]
    for it in range(steps):
        slab.run_virtual(steppers, 1)
        for k in ("p", "U", "density"):
            got = torch.cat([s.owned(k) for s in steppers], dim=3)
            assert torch.equal(got, refs[it][k]), (it, k, int((got != refs[it][k]).sum()))


def _torchrun(n, args, timeout):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + n), os.path.join(ROOT, "tools", "slab_check.py")] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.parametrize("world", [2, 4])
def test_multi_gpu_graphed_slab_step(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = _torchrun(world, ["--res", "1024", "--steps", "200", "--iters", "28"], timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "SLAB_CHECK_OK" in r.stdout
