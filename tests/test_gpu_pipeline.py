"""lib.HostStepPipeline: steps on host-resident states with upload / compute / download overlapped on three streams.
Each submitted state must come back exactly as lib.simulate computes it on the device -- distinct states in flight at
once, slots reused several times, graph-replayed (small grid) and direct-launch (large grid) paths."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("method,H,W", [("jacobi", 96, 160), ("jacobi", 2304, 2048), ("convnet", 128, 128)])
def test_pipeline_equals_device_steps(method, H, W):
    from fluidnet_cxx_b200.lib import fluid
    from fluidnet_cxx_b200.lib.host_pipeline import HostStepPipeline
    from fluidnet_cxx_b200.lib.pretrained import load_scalenet
    sim = importlib.import_module("fluidnet_cxx_b200.lib.simulate")
    from test_gpu_parity import plume_mconf
    net = None
    mconf = plume_mconf(simMethod=method)
    mconf["jacobiIter"] = 12
    if method == "convnet":
        net, mconf_net = load_scalenet("cuda")
        m = dict(mconf_net); m.update(mconf); mconf = m
        net.mconf = mconf; net.scale.mconf = mconf
    bd = {"p": torch.zeros(1, 1, 1, H, W, device="cuda"), "U": torch.zeros(1, 2, 1, H, W, device="cuda"),
          "flags": torch.zeros(1, 1, 1, H, W, device="cuda"), "density": torch.zeros(1, 1, 1, H, W, device="cuda")}
    fluid.emptyDomain(bd["flags"])
    bd["flags"][..., H // 2:H // 2 + 6, W // 3:W // 3 + 11] = 2.0
    fluid.createPlumeBCs(bd, mconf["injectionDensity"], mconf["injectionVelocity"], mconf["sourceRadius"])
    masks = {k: bd[k] for k in ("UBC", "UBCInvMask", "densityBC", "densityBCInvMask")}
    n = 7
    g = torch.Generator().manual_seed(H + W)
    ins, outs, want = [], [], []
    for i in range(n):
        st = {"p": torch.full((1, 1, 1, H, W), float("nan")).pin_memory(),       # p must never be read
              "U": (torch.randn(1, 2, 1, H, W, generator=g) * (0.2 + 0.1 * i)).pin_memory(),
              "flags": bd["flags"].cpu().pin_memory(), "density": torch.rand(1, 1, 1, H, W, generator=g).pin_memory()}
        ins.append(st)
        outs.append({k: torch.empty_like(st[k]).pin_memory() for k in ("p", "U", "density")})
        d = {k: v.cuda() for k, v in st.items()}
        d["p"] = torch.zeros_like(d["p"])
        d.update(masks)
        sim.clear_graph_cache()
        with torch.no_grad():
            sim.simulate(mconf, d, net, method)
        want.append({k: d[k].cpu() for k in ("p", "U", "density")})
    sim.clear_graph_cache()
    pipe = HostStepPipeline(mconf, net, method, like=ins[0], device="cuda", masks=masks, depth=2)
    for i in range(n):
        pipe.submit(ins[i], outs[i])
    pipe.flush()
    for i in range(n):
        for k in ("p", "U", "density"):
            a, b = outs[i][k], want[i][k]
            same = (a == b) | (torch.isnan(a) & torch.isnan(b))
            assert bool(same.all()), (i, k, int((~same).sum()))
    assert pipe.h2d_bytes_per_step == 4 * 4 * H * W and pipe.d2h_bytes_per_step == 4 * 4 * H * W
    sim.clear_graph_cache()
