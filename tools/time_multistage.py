#!/usr/bin/env python
"""The five maps of the reference's MobileNetV3_MultiStageNFP (models/texture_pooling.py:211-268) through the C ABI:
per-map forward + backward times, their sum, and ONE CUDA-graph replay of all ten calls back to back -- what a
captured training step pays.  If the replay costs no more than the sum of the members, the maps are not launch-bound once
the step is a graph, and a single grouped launch (second half) has nothing left to win.

    python tools/time_multistage.py [B] [dtype]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dt = sys.argv[2] if len(sys.argv) > 2 else "fp32"
MAPS = [(16, 112, 112), (24, 56, 56), (40, 28, 28), (112, 14, 14), (960, 7, 7)]
dev = torch.device("cuda:0")
lbs = [bench.LayerBench(dev, B, c, h, w, 1, dt) for c, h, w in MAPS]
n = 30
tot = 0.0
for (c, h, w), lb in zip(MAPS, lbs):
    tf = lb.timed(lb.fwd, n, 5) / n
    tb = lb.timed(lb.bwd_conservative, n, 5) / n
    tot += tf + tb
    print(f"B={B} {c}x{h}x{w} {dt}: {lb.path_fwd} / {lb.path_bwd}, launches {lb.launches}: fwd {tf * 1e6:.1f} us  bwd {tb * 1e6:.1f} us")


def all_maps(i):
    for lb in lbs:
        lb.fwd(i % lb.nbuf)
    for lb in reversed(lbs):
        lb.bwd_conservative(i % lb.nbuf)


for i in range(3):
    all_maps(i)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(4):
        all_maps(i)
g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    g.replay()
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) * 1e-3 / 40
nl = sum(lb.launches for lb in lbs)
print(f"sum of the members: {tot * 1e6:.1f} us; one graph replay of all five maps fwd + bwd ({nl} launches): {t * 1e6:.1f} us per set")

# eager launches (no graph): what an uncaptured training loop pays
torch.cuda.synchronize()
e0.record()
for i in range(20):
    all_maps(i)
e1.record()
torch.cuda.synchronize()
print(f"eager (ctypes launches from Python, no graph): {e0.elapsed_time(e1) * 1e-3 / 20 * 1e6:.1f} us per set")
