#!/usr/bin/env python
"""Tiny fused-path workload for compute-sanitizer (racecheck / memcheck), one tool per run:

    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neighbour_feature_pooling_b200 as nfpb  # noqa: E402
from neighbour_feature_pooling_b200 import functional as NF  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
for (B, C, H, W, R, dt) in [(3, 32, 7, 7, 1, torch.float32), (2, 16, 14, 14, 1, torch.float32),
                            (2, 32, 7, 7, 1, torch.bfloat16), (2, 16, 7, 7, 2, torch.float32),
                            (300, 8, 2, 2, 1, torch.float32)]:
    layer = nfpb.NFPPooling(C, R=R, measure="cosine", padding=R).to(dev)
    x = torch.randn(B, C, H, W, device=dev).to(dt).requires_grad_(True)
    y = layer(x)
    y.backward(torch.randn_like(y))
    xa = x.detach().clone().requires_grad_(True)
    a, n = NF.nfp_gap_pair(xa, layer.config)
    (a.sum() + n.sum()).backward()
    torch.cuda.synchronize()
    print("ok", B, C, H, W, R, dt, NF.describe((B, C, H, W), dt, layer.config))
