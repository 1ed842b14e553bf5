#!/usr/bin/env python
"""Per-phase timeline of the fused NFP kernels (diagnostic; needs a GPU).

    python tools/phase_timing.py [--shape l4|l3] [--R 1] [--dtype fp32|bf16] [--batch 256]

Uses nfpb200_debug_phase_timing: each CTA stamps %globaltimer at its phase boundaries.  Prints, per
kernel, the distribution over CTAs of each stamp relative to the earliest CTA start.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import SHAPES, LayerBench  # noqa: E402
from neighbour_feature_pooling_b200 import _capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="l4")
ap.add_argument("--R", type=int, default=1)
ap.add_argument("--dtype", default="fp32")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--layout", default="nchw")
args = ap.parse_args()
C, H, W, _ = SHAPES[args.shape]
dev = torch.device("cuda:0")
lb = LayerBench(dev, args.batch, C, H, W, args.R, args.dtype, layout=args.layout)
lib = _capi.load()
stamps = torch.zeros(8 * 8 * 1024, dtype=torch.int64, device=dev)
names = {"fwd": ["ready", "passA", "written"], "bwd": ["ready", "passA", "coef", "passB", "drained"]}
# cluster-split kernels (one CTA per channel slice): extra stamps 5 = first sub-chunk landed, 6 = partial tables
# exchanged, 7 = table reduced, 4 = CTA done; printed as per-CTA durations since the CTA's own start
split_order = {"fwd": [(5, "landed"), (1, "passA"), (6, "xchg"), (7, "reduced"), (2, "written"), (4, "done")],
               "bwd": [(5, "landed"), (1, "passA"), (6, "xchg"), (7, "reduced"), (2, "coef"), (3, "passB"), (4, "done")]}
for which, fn in (("fwd", lb.fwd), ("bwd", lb.bwd)):
    for i in range(4):
        fn(i % lb.nbuf)
    torch.cuda.synchronize()
    for rep in range(3):
        stamps.zero_()
        lib.nfpb200_debug_phase_timing(stamps.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn((rep + 1) % lb.nbuf); e1.record()
        torch.cuda.synchronize()
        lib.nfpb200_debug_phase_timing(None)
        s = stamps.view(-1, 8).cpu()
        used = s[:, 0] > 0
        s = s[used].double()
        t0 = s[:, 0].min()
        line = f"{which} rep{rep} ctas={int(used.sum())} event_us={e0.elapsed_time(e1) * 1e3:.1f} |"
        for k, nm in enumerate(names[which]):
            col = (s[:, k] - t0) / 1e3
            line += f" {nm}: min {col.min():.1f} med {col.median():.1f} max {col.max():.1f} |"
        print(line)
        if args.layout == "nchw" and (s[:, 5] > 0).any() and not (s[:, 7] > 0).any():   # ring kernels: stamps 5 / 6
            d51 = (s[:, 5] - s[:, 1]) / 1e3
            d65 = (s[:, 6] - s[:, 5]) / 1e3
            d26 = (s[:, 2] - s[:, 6]) / 1e3
            print(f"   per CTA: pass A (warp 0) -> all warps published: med {d51.median():.2f} max {d51.max():.2f} | -> table summed: med "
                  f"{d65.median():.2f} max {d65.max():.2f} | -> {'values written' if which == 'fwd' else 'coefficients ready'}: med "
                  f"{d26.median():.2f} max {d26.max():.2f} us")
        if args.layout == "nhwc":   # token kernels: stamps 0..6 per (CTA, image), durations since the row's own start
            nm = ["landed", "stencil", "gram", "coef", "Mbuilt", "done"]
            line = "   per (CTA, image) since its start |"
            for kk in range(1, 7):
                ok = s[:, kk] > 0
                if ok.any():
                    col = (s[ok, kk] - s[ok, 0]) / 1e3
                    line += f" {nm[kk - 1]}: {col.median():.2f} ({col.min():.2f}..{col.max():.2f}) |"
            print(line)
        elif (s[:, 7] > 0).any():
            line = f"   per-CTA (since its own start; start = {((s[:, 0] - t0) / 1e3).median():.1f} med / {((s[:, 0] - t0) / 1e3).max():.1f} max us after the first CTA) |"
            for k, nm in split_order[which]:
                col = (s[:, k] - s[:, 0]) / 1e3
                col = col[s[:, k] > 0]
                line += f" {nm}: {col.median():.2f} ({col.min():.2f}..{col.max():.2f}) |"
            print(line)
            starts = ((s[:, 0] - t0) / 1e3)
            hist = torch.histc(starts.float(), bins=12, min=0, max=float(starts.max()) + 1e-3)
            print("   CTA start histogram (12 bins up to %.1f us): %s" % (float(starts.max()), [int(v) for v in hist]))
