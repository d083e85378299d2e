"""Where does the MMA warp of k_conv_tc wait?  Runs single layers with fnx_tc_set_debug and prints, per layer,
the mean clocks per CTA spent waiting for activation stages / weight slots / accumulators vs the total."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from fluidnet_cxx_b200 import _native as N
from test_gpu_cnn import _tc_conv, cu
lib = N.load()
res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dbg = torch.zeros(4 * 148, dtype=torch.int64, device="cuda")
for cin, cout, k in ((64, 128, 3), (128, 64, 3), (64, 32, 3), (32, 64, 3), (32, 8, 5), (3, 32, 5)):
    rng = np.random.RandomState(cin + cout)
    x = cu(rng.randn(cin, res, res).astype(np.float32))
    w = cu((rng.randn(cout, cin, k, k) / np.sqrt(cin * k * k)).astype(np.float32))
    b = cu(rng.randn(cout).astype(np.float32))
    _tc_conv(lib, N, x, w, b, 1, 1)            # warm-up
    dbg.zero_()
    lib.fnx_tc_set_debug(dbg.data_ptr())
    _tc_conv(lib, N, x, w, b, 1, 1)
    lib.fnx_tc_set_debug(None)
    d = dbg.view(148, 4).double().cpu().numpy()
    d = d[d[:, 3] > 0]
    m = d.mean(0)
    print(f"{cin:3d}->{cout:3d} k{k} @{res}: total {m[3]:9.0f} clk/CTA | wait A {m[0]:8.0f} ({100*m[0]/m[3]:4.1f}%)  "
          f"W {m[1]:8.0f} ({100*m[1]/m[3]:4.1f}%)  ACC {m[2]:8.0f} ({100*m[2]/m[3]:4.1f}%)  issue+other "
          f"{100*(m[3]-m[0]-m[1]-m[2])/m[3]:4.1f}%")
