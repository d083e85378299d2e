"""MultiScaleNet: the 3-scale pressure CNN (reference: pytorch/lib/multi_scale_net.py:21-127).

Same module tree and parameter names as the reference (`convN_4.encode.{0,2,4,6}`,
`convN_2.encode.{0,...,10}`, `convN_1.encode.{0,...,10}`, `final`; the Dropout placeholders keep
the Sequential indices), so the shipped `convModel_lastEpoch_best.pth` loads with
`load_state_dict`.  `forward` runs the hand-written sm_100a convolution / resize kernels through
the C-ABI (inference only; there is no torch.nn fallback on the forward path).
"""
import ctypes
import os

import torch
import torch.nn as nn

from .. import _native as N


def _block(channels, kernels, relu_count, dropout=True):
    """Conv2d chain; ReLU after the first `relu_count` convs; Dropout before the last conv."""
    layers = []
    n = len(kernels)
    for i in range(n):
        if i == n - 1 and dropout:
            layers.append(nn.Dropout())
        layers.append(nn.Conv2d(channels[i], channels[i + 1], kernel_size=kernels[i], padding=kernels[i] // 2))
        if i < relu_count:
            layers.append(nn.ReLU(inplace=(i == 0)))
    return nn.Sequential(*layers)


class _ConvBlock(nn.Module):
    def __init__(self, channels, kernels, relu_count):
        super().__init__()
        self.encode = _block(channels, kernels, relu_count)
        self.relu_count = relu_count

    def convs(self):
        return [m for m in self.encode if isinstance(m, nn.Conv2d)]


class MultiScaleNet(nn.Module):
    def __init__(self, data_channels):
        super().__init__()
        c = data_channels
        # multi_scale_net.py:104-107
        self.convN_4 = _ConvBlock([c, 32, 64, 32, 1], [3, 3, 3, 3], 2)
        self.convN_2 = _ConvBlock([c + 1, 32, 64, 128, 64, 32, 1], [5, 3, 3, 3, 3, 3], 4)
        self.convN_1 = _ConvBlock([c + 1, 32, 64, 128, 64, 32, 8], [5, 3, 3, 3, 3, 5], 4)
        self.final = nn.Conv2d(8, 1, kernel_size=1)

    # -- native plan -------------------------------------------------------------------------
    # fnx_msnet_plan (include/fluidstep.h): device pointers to the fp32 weights / biases, the
    # split-fp16 packed weights of the tensor-core layers and their range scalars.  Built lazily on
    # the first forward on a device and rebuilt whenever a parameter tensor changes.
    USE_TENSOR_CORES = True   # False: every layer on the fp32 direct kernel (numerical cross-check)
    WEIGHT_REPLICAS = int(os.environ.get("FNX_TC_WEIGHT_REPLICAS", "1"))

    def _layers(self):
        out = []
        for blk in (self.convN_4, self.convN_2, self.convN_1):
            convs = blk.convs()
            out.append([(c, i < blk.relu_count) for i, c in enumerate(convs)])
        out.append([(self.final, False)])
        return out

    def _plan_key(self, device):
        return (str(device), bool(self.USE_TENSOR_CORES), int(self.WEIGHT_REPLICAS)) + tuple(
            (p.data_ptr(), p._version) for p in self.parameters())

    def _build_plan(self, device):
        import math
        lib = N.load()
        st = torch.cuda.current_stream(device).cuda_stream
        keep = []
        plan = N.MsnetPlan()
        plan.data_channels = self.convN_4.convs()[0].in_channels

        def fill(dst, conv, relu):
            w = conv.weight.detach().contiguous().float()
            b = conv.bias.detach().contiguous().float()
            if w.device != device or b.device != device:
                raise RuntimeError(f"MultiScaleNet parameters live on {w.device} but the input is on {device}: "
                                   "move the model first (net.to(device)); there is no CPU path")
            keep.extend([w, b])
            cout, cin, k, _ = w.shape
            dst.weight, dst.bias = w.data_ptr(), b.data_ptr()
            dst.cin, dst.cout, dst.ksize, dst.relu = cin, cout, k, int(relu)
            dst.w_tc = None
            dst.w_replicas = 1
            dst.w_scale, dst.w_norm, dst.b_max = 1.0, 0.0, 0.0
            if self.USE_TENSOR_CORES and (k == 3 or (k == 5 and cout <= 32)) and cin <= 128 and cout <= 128:
                wmax = float(w.abs().max())
                # power of two with max|w| * w_scale <= 2^14 (same rule as the device side)
                w_scale = 1.0 if not (wmax > 0 and math.isfinite(wmax)) else 2.0 ** (14 - math.frexp(wmax)[1])
                nbytes = lib.fnx_tc_weight_bytes(cin, cout, k)
                nrep = self.WEIGHT_REPLICAS
                packed = torch.empty(nrep * nbytes, dtype=torch.uint8, device=device)
                N.check(lib.fnx_tc_pack_weights(w.data_ptr(), cin, cout, k, w_scale, packed.data_ptr(), st),
                        "MultiScaleNet.pack_weights")
                for r in range(1, nrep):      # identical copies: every SM streams the same slots at the same time
                    packed[r * nbytes:(r + 1) * nbytes].copy_(packed[:nbytes])
                dst.w_replicas = nrep
                keep.append(packed)
                dst.w_tc = packed.data_ptr()
                dst.w_scale = w_scale
                dst.w_norm = float(w.abs().sum(dim=(1, 2, 3)).max())
                dst.b_max = float(b.abs().max())

        groups = self._layers()
        for dst_arr, grp in zip((plan.quarter, plan.half, plan.full), groups[:3]):
            for i, (conv, relu) in enumerate(grp):
                fill(dst_arr[i], conv, relu)
        fill(plan.final_conv, *groups[3][0])
        return plan, keep

    def _get_plan(self, device):
        key = self._plan_key(device)
        cache = self.__dict__.setdefault("_fnx_plan_cache", {})
        if cache.get("key") != key:
            cache.clear()
            cache["plan"], cache["keep"] = self._build_plan(device)
            cache["key"] = key
            cache["ws"] = {}
        return cache

    def __getstate__(self):
        d = self.__dict__.copy()
        d.pop("_fnx_plan_cache", None)      # ctypes plan + device buffers: rebuilt on demand
        return d

    def forward(self, x):
        if self.training:
            raise RuntimeError("fluidnet_cxx_b200.MultiScaleNet is inference-only: call .eval() "
                               "(training is outside the B200 hot path)")
        assert x.dim() == 4, "MultiScaleNet expects (batch, channels, height, width)"
        x = x.contiguous()
        lib = N.load()
        n, c, h, w = x.shape
        cache = self._get_plan(x.device)
        plan = cache["plan"]
        assert c == plan.data_channels, "MultiScaleNet: wrong number of input channels"
        st = N.stream_of(x)
        # one workspace per (resolution, stream): simulations on different streams never share activations.  A
        # CUDA-graph capture (its own stream) re-uses the workspace of the eager warm-up call that precedes it --
        # zero-filling a fresh one would put the memset of the whole workspace into every replay.
        wkey = (n, h, w, st)
        ws = cache["ws"].get(wkey)
        if ws is None and torch.cuda.is_current_stream_capturing():
            ws = next((v for k, v in cache["ws"].items() if k[:3] == (n, h, w)), None)
        if ws is None:
            nbytes = lib.fnx_msnet_workspace_n(ctypes.byref(plan), n, h, w)
            if nbytes == 0:
                N.check(-1, "MultiScaleNet.workspace")
            ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
            # zero borders of the split activation buffers: written once, never touched again
            N.check(lib.fnx_msnet_workspace_init(ws.data_ptr(), ws.numel(), st), "MultiScaleNet.workspace_init")
            if len(cache["ws"]) >= 4:    # a few resolutions stay resident (captured graphs point at them)
                cache["ws"].pop(next(iter(cache["ws"])))
            cache["ws"][wkey] = ws
        y = torch.empty((n, 1, h, w), dtype=torch.float32, device=x.device)
        N.check(lib.fnx_msnet_forward(ctypes.byref(plan), N.ptr(x), N.ptr(y), n, h, w, ws.data_ptr(), ws.numel(), st),
                "MultiScaleNet.forward")
        return y
