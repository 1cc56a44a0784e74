"""FP8 (e4m3) 1x1 convolutions with delayed scaling for ResNet-50 local training (BASELINE.json
configs[4]). A 1x1 convolution on NHWC activations is a plain GEMM ``[pixels, Cin] x [Cin, Cout]``:
operands are quantised by the hand-written one-pass kernel ``fp8_quantize`` (scale from the previous
step's amax, this step's amax recorded for the next — csrc/nn/resnet_kernels.cu) and multiplied by the
library fp8 GEMM (cuBLASLt through ``torch._scaled_mm``); gradients stay in bf16.

Off by default; ``FLConfig(dtype="fp8")`` turns it on per model (``set_model_fp8``) for the ResNets
(everything else — 3x3/7x7 convolutions, BatchNorm, the head — runs in bf16)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _ext

ENABLE = False
E4M3_MAX = 448.0
MARGIN = 2.0            # head-room for values that grow between two steps (delayed scaling)


def set_model_fp8(model: nn.Module, on: bool) -> int:
    """Switch every ``Conv1x1`` of ``model`` to fp8 (or back); returns how many there are."""
    n = 0
    for m in model.modules():
        if isinstance(m, Conv1x1):
            m.use_fp8 = bool(on)
            n += 1
    return n


class DelayedScale:
    """scale (quantisation multiplier) and running amax of one tensor role, device-resident so the
    whole step stays CUDA-graph capturable."""

    def __init__(self, device):
        self.scale = torch.ones(1, dtype=torch.float32, device=device)
        self.inv = torch.ones(1, dtype=torch.float32, device=device)
        self.amax = torch.zeros(1, dtype=torch.float32, device=device)
        self.ready = False

    def refresh(self, x: torch.Tensor) -> None:
        if not self.ready:                      # first use: take the scale from the data itself
            self.amax.copy_(x.detach().abs().amax().float().reshape(1))
            self.ready = True
        # scale <- (448 / margin) / amax (kept if amax == 0), inv <- 1 / scale, amax <- 0: one tiny launch
        _ext.ops().fp8_scale_update(self.amax, self.scale, self.inv, E4M3_MAX / MARGIN)


def _quantize(x2: torch.Tensor, st: DelayedScale) -> torch.Tensor:
    st.refresh(x2)
    q = torch.empty(x2.shape, dtype=torch.uint8, device=x2.device)
    _ext.ops().fp8_quantize(x2, q, st.scale, st.amax)
    return q.view(torch.float8_e4m3fn)


class _FP8Conv1x1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, sx, sw):
        B, Cin, H, W = x.shape
        Cout = w.shape[0]
        x2 = x.permute(0, 2, 3, 1).reshape(B * H * W, Cin)                 # zero-copy for channels_last input
        wb = w.reshape(Cout, Cin).to(torch.bfloat16).contiguous()
        xq = _quantize(x2, sx)
        wq = _quantize(wb, sw)
        y = torch._scaled_mm(xq, wq.t(), scale_a=sx.inv, scale_b=sw.inv, out_dtype=torch.bfloat16)
        ctx.save_for_backward(x2, wb)
        ctx.shape = (B, Cin, H, W, Cout)
        return y.view(B, H, W, Cout).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        x2, wb = ctx.saved_tensors
        B, Cin, H, W, Cout = ctx.shape
        g2 = gy.permute(0, 2, 3, 1).reshape(B * H * W, Cout).to(torch.bfloat16)
        dx = (g2 @ wb).view(B, H, W, Cin).permute(0, 3, 1, 2)
        dw = (g2.t() @ x2).float().view(Cout, Cin, 1, 1)
        return dx, dw, None, None


class Conv1x1(nn.Conv2d):
    """1x1 convolution (no bias) that runs as an fp8 GEMM when ``fp8.ENABLE`` is set and the input is
    NHWC bf16 on CUDA; otherwise a regular ``nn.Conv2d``."""

    def __init__(self, cin: int, cout: int, stride: int = 1):
        super().__init__(cin, cout, 1, stride, 0, bias=False)
        self._fp8_state = None
        self.use_fp8 = None          # None: follow the module-level ENABLE; True / False: per-model choice (trainer)
        self.use_tc = None           # hand-written tcgen05 GEMM (ops/tc_conv.py) instead of cuDNN / cuBLASLt

    def forward(self, x):
        on = ENABLE if self.use_fp8 is None else self.use_fp8
        from . import tc_conv

        if (tc_conv.ENABLE if self.use_tc is None else self.use_tc) and tc_conv.eligible(x, self.in_channels, self.out_channels):
            # fp8 on the hand-written path = block-scaled e4m3 (MX): scales come from the data of this very call
            return tc_conv.conv1x1(x, self.weight, self.stride[0], "mx" if on else None)
        ok = (on and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4
              and self.in_channels % 16 == 0 and self.out_channels % 16 == 0)
        if not ok:
            return super().forward(x)
        xs = x[:, :, ::self.stride[0], ::self.stride[1]] if self.stride[0] > 1 else x
        if (xs.shape[0] * xs.shape[2] * xs.shape[3]) % 16 != 0:
            return super().forward(x)
        xs = xs.contiguous(memory_format=torch.channels_last)
        if self._fp8_state is None:
            self._fp8_state = (DelayedScale(x.device), DelayedScale(x.device))
        return _FP8Conv1x1.apply(xs, self.weight, *self._fp8_state)
