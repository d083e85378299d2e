"""Run the MultiScaleNet forward a few times on a synthetic input (for ncu captures)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from fluidnet_cxx_b200.lib.pretrained import load_scalenet
res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
model, _ = load_scalenet("cuda")
x = torch.randn(1, 2, res, res, device="cuda")
x[:, 1] = (x[:, 1] > 0.8).float()
with torch.no_grad():
    for _ in range(reps):
        y = model.multiScale(x)
torch.cuda.synchronize()
# timing without a profiler
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
with torch.no_grad():
    ev[0].record()
    for _ in range(10):
        y = model.multiScale(x)
    ev[1].record()
torch.cuda.synchronize()
print(f"msnet forward {res}x{res}: {ev[0].elapsed_time(ev[1]) / 10:.3f} ms")
