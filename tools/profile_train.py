#!/usr/bin/env python
"""Kernel timeline of one captured ResNet18+NFP training step under DDP (what limits the multi-GPU scaling?).

    torchrun --nproc-per-node N tools/profile_train.py [--config eurosat] [--batch 256]

nsys is not installed in this image; torch.profiler (Kineto / CUPTI) sees the kernels inside CUDA-graph replays.
Rank 0 prints: the step time, the NCCL kernels (name, launches per step, time per step), how much of their time
overlaps compute kernels, and the compute-kernel time with and without DDP's all-reduce in flight.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_train  # noqa: E402
from neighbour_feature_pooling_b200 import sharding  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="eurosat")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=4)
args = ap.parse_args()
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
rank, _, world = sharding.init_from_env("nccl", torch.device("cuda", local))

prof = torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA])
state = {"n": 0}
orig_event = torch.cuda.Event


def hooked_run():
    # run_gpu times `steps` replays between two events; start the profiler right before them
    return bench_train.run_gpu(args.config, args.batch, args.steps, 5)


prof.__enter__()
out = hooked_run()
torch.cuda.synchronize()
prof.__exit__(None, None, None)
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time_total > 0]
    # keep the last `steps` graph replays: the kernels after the last long gap
    evs.sort(key=lambda e: e.time_range.start)
    step_us = out["ms_per_step"] * 1e3
    t_end = evs[-1].time_range.end
    win = [e for e in evs if e.time_range.start >= t_end - args.steps * step_us]
    nccl = [e for e in win if "nccl" in e.name.lower()]
    comp = [e for e in win if "nccl" not in e.name.lower() and "memcpy" not in e.name.lower()]

    def union(iv):
        iv = sorted(iv)
        tot, cur_s, cur_e = 0.0, None, None
        for s, e in iv:
            if cur_e is None or s > cur_e:
                if cur_e is not None:
                    tot += cur_e - cur_s
                cur_s, cur_e = s, e
            else:
                cur_e = max(cur_e, e)
        if cur_e is not None:
            tot += cur_e - cur_s
        return tot

    civ = [(e.time_range.start, e.time_range.end) for e in comp]
    niv = [(e.time_range.start, e.time_range.end) for e in nccl]
    busy_comp, busy_nccl, busy_any = union(civ), union(niv), union(civ + niv)
    print(f"config {args.config} world {world} batch/GPU {args.batch}: {out['ms_per_step']:.3f} ms/step, {out['images_per_s']:.0f} img/s")
    print(f"per step (us): compute-kernel busy {busy_comp / args.steps:.0f}, NCCL-kernel busy {busy_nccl / args.steps:.0f}, "
          f"either busy {busy_any / args.steps:.0f}, NCCL time hidden behind compute {(busy_comp + busy_nccl - busy_any) / args.steps:.0f}, "
          f"idle {step_us - busy_any / args.steps:.0f}")
    by = {}
    for e in nccl:
        d = by.setdefault(e.name[:90], [0, 0.0])
        d[0] += 1
        d[1] += e.device_time_total
    for k, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print(f"  NCCL kernel {k}: {n / args.steps:.1f} launches/step, {t / args.steps:.0f} us/step")
    byc = {}
    for e in comp:
        d = byc.setdefault(e.name[:70], [0, 0.0])
        d[0] += 1
        d[1] += e.device_time_total
    print("  top compute kernels (us/step):")
    for k, (n, t) in sorted(byc.items(), key=lambda kv: -kv[1][1])[:8]:
        print(f"    {t / args.steps:7.0f}  x{n / args.steps:.0f}  {k}")
import torch.distributed as dist
if dist.is_initialized():
    dist.destroy_process_group()
