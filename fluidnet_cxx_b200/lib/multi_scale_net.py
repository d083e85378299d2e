"""MultiScaleNet: the 3-scale pressure CNN (reference: pytorch/lib/multi_scale_net.py:21-127).

Same module tree and parameter names as the reference (`convN_4.encode.{0,2,4,6}`,
`convN_2.encode.{0,...,10}`, `convN_1.encode.{0,...,10}`, `final`; the Dropout placeholders keep
the Sequential indices), so the shipped `convModel_lastEpoch_best.pth` loads with
`load_state_dict`.  `forward` runs the hand-written sm_100a convolution / resize kernels through
the C-ABI (inference only; there is no torch.nn fallback on the forward path).
"""
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

    # -- C-ABI helpers ----------------------------------------------------------------------
    @staticmethod
    def _conv(lib, x, conv, relu, out=None, coff=0):
        n, cin, h, w = x.shape
        cout, k = conv.out_channels, conv.kernel_size[0]
        if out is None:
            out = torch.empty((n, cout, h, w), dtype=torch.float32, device=x.device)
        wt, bs = conv.weight.detach(), conv.bias.detach()
        N.check(lib.fnx_conv2d(N.ptr(x), N.ptr(wt), N.ptr(bs), N.ptr(out), n, cin, h, w, cout, k, int(relu),
                               out.size(1), coff, N.stream_of(x)), "MultiScaleNet.conv")
        return out

    @staticmethod
    def _resize(lib, x, size, out, coff):
        n, c, h, w = x.shape
        N.check(lib.fnx_resize_bilinear(N.ptr(x), N.ptr(out), n, c, h, w, size[0], size[1], out.size(1), coff,
                                        N.stream_of(x)), "MultiScaleNet.resize")

    def _run_block(self, lib, block, x):
        convs = block.convs()
        for i, conv in enumerate(convs):
            x = self._conv(lib, x, conv, relu=i < block.relu_count)
        return x

    def forward(self, x):
        if self.training:
            raise RuntimeError("fluidnet_cxx_b200.MultiScaleNet is inference-only: call .eval() "
                               "(training is outside the B200 hot path)")
        assert x.dim() == 4, "MultiScaleNet expects (batch, channels, height, width)"
        x = x.contiguous()
        lib = N.load()
        n, c, h, w = x.shape
        quarter = (int(h * 0.25), int(w * 0.25))
        half = (int(h * 0.5), int(w * 0.5))
        dev = x.device
        # quarter scale
        x4 = torch.empty((n, c) + quarter, dtype=torch.float32, device=dev)
        self._resize(lib, x, quarter, x4, 0)
        o4 = self._run_block(lib, self.convN_4, x4)
        # half scale: cat(down(x), up(o4))
        in2 = torch.empty((n, c + 1) + half, dtype=torch.float32, device=dev)
        self._resize(lib, x, half, in2, 0)
        self._resize(lib, o4, half, in2, c)
        o2 = self._run_block(lib, self.convN_2, in2)
        # full scale: cat(x, up(o2))
        in1 = torch.empty((n, c + 1, h, w), dtype=torch.float32, device=dev)
        self._resize(lib, x, (h, w), in1, 0)
        self._resize(lib, o2, (h, w), in1, c)
        o1 = self._run_block(lib, self.convN_1, in1)
        return self._conv(lib, o1, self.final, relu=False)
