"""Multi-GPU check of the window-based slab step (run under torchrun, one rank per GPU):
N steps of SlabJacobiStep (symmetric memory, halo-exchange kernel, whole step as CUDA graphs) against the
single-GPU fused step run on rank 0, owned rows bit for bit.  Prints SLAB_CHECK_OK on success.
    torchrun --nproc-per-node N tools/slab_check.py [--res 1024] [--steps 200] [--iters 28] [--K 1]"""
import argparse
import datetime
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--iters", type=int, default=28)
    ap.add_argument("--K", type=int, default=1)
    ap.add_argument("--no-graph", action="store_true")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from fluidnet_cxx_b200.lib.hang_guard import HangGuard
    guard = HangGuard(90, who=f"slab_check rank {rank}/{world}")
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    from fluidnet_cxx_b200.lib import fluid, slab
    sim = importlib.import_module("fluidnet_cxx_b200.lib.simulate")
    H = W = a.res
    wl = dict(bench.WORKLOADS["plume4096_jacobi100"], res=(1, H, W), jacobi_iters=a.iters)
    mconf = bench.workload_mconf(wl)
    # the global initial state on the HOST; only held rows ever go to a GPU
    U_np, rho_np = bench.synthetic_state_numpy(1, H, W, 0)
    host = {"p": torch.zeros(1, 1, 1, H, W), "U": torch.zeros(1, 2, 1, H, W), "flags": torch.zeros(1, 1, 1, H, W),
            "density": torch.zeros(1, 1, 1, H, W)}
    if rank == 0:
        bd = {k: v.to(dev) for k, v in host.items()}
        bench.init_state(fluid, wl, mconf, bd, U_np, rho_np, lambda x: torch.from_numpy(x).to(dev))
        full = {k: v.cpu() for k, v in bd.items()}
    else:
        full = None
    objs = [full]
    dist.broadcast_object_list(objs, src=0)
    full = objs[0]
    guard.beat("set-up")
    step = slab.SlabJacobiStep(slab.ProcessTopology(dev), mconf, H, W,
                               lambda name, r0, r1: full[name][:, :, :, r0:r1].contiguous() if name in full else None,
                               K=a.K, use_graph=not a.no_graph)
    for i in range(2):
        guard.beat(f"direct step {i}")
        step.step()
    graphed = step.capture()
    for i in range(2, a.steps):
        if i % 20 == 0:
            guard.beat(f"step {i}")
        step.step()
    torch.cuda.synchronize()
    guard.beat("gather")
    ok = True
    for k in ("p", "U", "density"):
        mine = step.owned(k).contiguous()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        if rank == 0:
            got = torch.cat(parts, dim=3)
            if k == "p":
                ref = {kk: v.to(dev) for kk, v in full.items()}
                sim.clear_graph_cache()
                os.environ["FLUIDNET_B200_GRAPHS"] = "0"
                for _ in range(a.steps):
                    sim.simulate(mconf, ref, None, "jacobi")
            bad = int((got != ref[k]).sum())
            print(f"[slab_check] world {world} {H}x{W} steps {a.steps} graphs {graphed}: {k} differs in {bad} cells", flush=True)
            ok = ok and bad == 0
    reach = step.max_reach()
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.broadcast(t, src=0)
    guard.stop()
    if rank == 0 and t.item() == 1.0:
        print(f"SLAB_CHECK_OK world={world} graphs={graphed} max|u|dt={reach:.3f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
