"""Time of ONE 8-iteration launch of the blocked 2-D Jacobi kernel on H x 4096 grids (H = a slab of a strong-scaled
4096^2 grid at 8 / 4 / 2 / 1 GPUs).  FNX_JACOBI_PACKED=0/1 and FNX_JACOBI_NW=8/12 select the variant.
    python tools/jacobi_launch_time.py [H ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fluidnet_cxx_b200 import _native as N
lib = N.load()
W = 4096
L2 = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
for H in [int(a) for a in sys.argv[1:]] or [546, 1058, 2082, 4096]:
    g = torch.Generator(device="cuda").manual_seed(1)
    flags = torch.ones(1, 1, 1, H, W, device="cuda")
    flags[..., 0, :] = 2; flags[..., -1, :] = 2; flags[..., :, 0] = 2; flags[..., :, -1] = 2
    div = torch.randn(1, 1, 1, H, W, device="cuda", generator=g)
    p0 = torch.randn(1, 1, 1, H, W, device="cuda", generator=g)
    p1 = torch.empty_like(p0)
    ws = N.workspaces.get(flags.device, "jacobi", lib.fnx_jacobi_workspace(1, 1, H, W, 8))
    ts = []
    for rep in range(12):
        L2.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        N.check(lib.fnx_jacobi_iterate(N.ptr(flags), N.ptr(div), N.ptr(p0), N.ptr(p1), 1, 1, H, W, 0, 8, 0, 0,
                                       ws.data_ptr(), ws.numel(), N.stream_of(flags)), "jacobi")
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:])
    print(f"H={H:5d} packed={os.environ.get('FNX_JACOBI_PACKED', 'auto')} nw={os.environ.get('FNX_JACOBI_NW', 'auto')}: "
          f"median {ts[len(ts) // 2]:.1f} us  min {ts[0]:.1f} us  (1 launch = tile masks + 8 iterations)")
