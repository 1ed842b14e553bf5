"""TEST / BASELINE INFRASTRUCTURE ONLY -- conv-form CPU port of the reference's cosine NFP.

The gather-form oracle (``nfp_oracle.py``) is the parity checker.  This module restates the SAME
algorithm with the SAME ATen operator sequence the reference executes, so that timing it on host
cores is a faithful stand-in for the reference's CPU path (``bench.py``'s ``cpu_baseline`` leg and
``--impl reference`` arm; the Python reference itself cannot travel to the GPU box):

    reflect-pad + depthwise conv with one-hot "centre" kernels      (reference nfp.py:53-61, 152)
    reflect-pad + grouped conv C -> K*C with one-hot "tap" kernels  (reference nfp.py:42-50, 64-82, 153)
    view (B, K*C, H', W') as (B, C, K, H', W')                      (reference nfp.py:136-139, 154)
    F.cosine_similarity(centre[:, :, None], taps, dim=1, eps)       (reference nfp.py:155-159)

and autograd through that chain for the backward.  Checked against the gather-form oracle in
``tests/test_oracle_golden.py`` and against the reference in ``oracle/check_against_reference.py``.
Never imported by the product package.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .nfp_oracle import tap_offsets


class ConvFormCosineNFP(torch.nn.Module):
    def __init__(self, channels: int, R: int = 1, stride: int = 1, padding: int = 0, dilation: int = 1,
                 padding_mode: str = "reflect", similarity: bool = True, eps: float = 1e-6):
        super().__init__()
        k = 2 * R + 1
        taps = tap_offsets(R)
        K = len(taps)
        w_tap = torch.zeros(channels, K, k, k)
        for n, (a, b) in enumerate(taps):
            w_tap[:, n, a, b] = 1.0
        w_ctr = torch.zeros(channels, 1, k, k)
        w_ctr[:, 0, R, R] = 1.0
        # frozen Parameters, as in the reference (they show up in named_parameters / state_dict)
        self.w_tap = torch.nn.Parameter(w_tap.reshape(channels * K, 1, k, k), requires_grad=False)
        self.w_ctr = torch.nn.Parameter(w_ctr, requires_grad=False)
        self.C, self.K = channels, K
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.mode = {"zeros": "constant"}.get(padding_mode, padding_mode)
        self.similarity, self.eps = similarity, eps

    def _extract(self, x, w):
        if self.padding > 0:
            x = F.pad(x, (self.padding,) * 4, mode=self.mode)
        return F.conv2d(x, w, None, self.stride, 0, self.dilation, self.C)

    def forward(self, x):
        centre = self._extract(x, self.w_ctr)
        taps = self._extract(x, self.w_tap)
        B, _, Ho, Wo = taps.shape
        taps = taps.view(B, self.C, self.K, Ho, Wo)
        y = F.cosine_similarity(centre.unsqueeze(2), taps, dim=1, eps=self.eps)
        return y if self.similarity else 1 - y


def forward_backward(layer: ConvFormCosineNFP, x: torch.Tensor, gy: torch.Tensor):
    """One reference-style training pass of the layer: forward, then autograd backward."""
    xr = x.detach().requires_grad_(True)
    y = layer(xr)
    y.backward(gy)
    return y.detach(), xr.grad
