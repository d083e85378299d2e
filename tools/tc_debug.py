"""First-light check of the tensor-core conv path: prints relative errors instead of asserting."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
from fluidnet_cxx_b200 import _native as N
from test_gpu_cnn import _tc_conv, cu, rel_err, TC_SHAPES
import torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
lib = N.load()
for cin, cout, h, w, k in TC_SHAPES + [(64, 128, 512, 512, 3), (32, 8, 512, 512, 5)]:
    rng = np.random.RandomState(cin + cout)
    x = rng.randn(cin, h, w).astype(np.float32)
    wt = (rng.randn(cout, cin, k, k) / np.sqrt(cin * k * k)).astype(np.float32)
    b = rng.randn(cout).astype(np.float32)
    ref = F.conv2d(cu(x)[None].double(), cu(wt).double(), cu(b).double(), padding=k // 2)[0].cpu().numpy()
    y = _tc_conv(lib, N, cu(x), cu(wt), cu(b), 0, 1).cpu().numpy()
    e = np.abs(y[2:2 + cout] - ref)
    print(f"{cin}->{cout} k{k} {h}x{w}: rel_err={e.max() / np.abs(ref).max():.3e}  worst at {np.unravel_index(e.argmax(), e.shape)}  "
          f"fp32 cudnn rel_err={rel_err(F.conv2d(cu(x)[None], cu(wt), cu(b), padding=k // 2)[0].cpu().numpy(), ref):.3e}", flush=True)
    if e.max() / np.abs(ref).max() > 1e-4:
        bad = e > 1e-4 * np.abs(ref).max()
        print("   bad fraction", bad.mean(), "bad channels", np.unique(np.nonzero(bad)[0])[:20], "rows", np.unique(np.nonzero(bad)[1])[:20],
              "cols", np.unique(np.nonzero(bad)[2])[:20])
