"""Per-layer time of the MultiScaleNet forward for forced rows-per-block values (FNX_TC_ROWS) next to
the model's own choice: checks conv_tc.cu::choose_rows against measurements."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from fluidnet_cxx_b200 import _native as N
from fluidnet_cxx_b200.lib.pretrained import load_scalenet
res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lib = N.load()
model, _ = load_scalenet("cuda")
x = torch.randn(1, 2, res, res, device="cuda")
x[:, 1] = (x[:, 1] > 0.8).float()
table = {}
for forced in ("model", "1", "2", "3", "4", "6", "8"):
    if forced == "model":
        os.environ.pop("FNX_TC_ROWS", None)
    else:
        os.environ["FNX_TC_ROWS"] = forced
    with torch.no_grad():
        for _ in range(3):
            model.multiScale(x)
        torch.cuda.synchronize()
        lib.fnx_profile_enable(1)
        for _ in range(10):
            model.multiScale(x)
        torch.cuda.synchronize()
    buf = (N.ProfileRec * 4096)()
    n = lib.fnx_profile_fetch(buf, 4096)
    lib.fnx_profile_enable(0)
    for i in range(n):
        r = buf[i]
        key = f"{r.cin}->{r.cout} k{r.ksize} @{r.h}"
        table.setdefault(key, {}).setdefault(forced, []).append(r.ms * 1e3)
print(f"{'layer':22s}" + "".join(f"{c:>9s}" for c in ("model", "1", "2", "3", "4", "6", "8")))
for key, cols in table.items():
    print(f"{key:22s}" + "".join(f"{sum(cols[c]) / len(cols[c]):9.1f}" for c in ("model", "1", "2", "3", "4", "6", "8")))
