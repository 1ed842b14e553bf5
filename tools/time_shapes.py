#!/usr/bin/env python
"""Time forward / backward of arbitrary (C,H,W) map shapes through the C ABI (whatever path they take).

    python tools/time_shapes.py B C H W [R] [dtype] ...   e.g.  python tools/time_shapes.py 64 16 112 112
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

args = sys.argv[1:]
B, C, H, W = map(int, args[:4])
R = int(args[4]) if len(args) > 4 else 1
dt = args[5] if len(args) > 5 else "fp32"
lay = args[6] if len(args) > 6 else "nchw"
dev = torch.device("cuda:0")
lb = bench.LayerBench(dev, B, C, H, W, R, dt, layout=lay)
n = 50
tf = lb.timed(lb.fwd, n, 5) / n
tb = lb.timed(lb.bwd_conservative, n, 5) / n
fb, bb = bench.algorithmic_bytes(B, C, H, W, R, lb.esz)
print(f"B={B} {C}x{H}x{W} R={R} {dt} {lay}: path {lb.path_fwd} / {lb.path_bwd}  fwd {tf * 1e6:.1f} us ({fb / tf / 1e9:.0f} GB/s)  "
      f"bwd {tb * 1e6:.1f} us ({bb / tb / 1e9:.0f} GB/s)")
