#!/usr/bin/env python
"""Eager (un-graphed) update steps of one BASELINE config for profiling under ncu:
    ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py --updates 6
The last 2 updates (one with, one without the Polyak step) are the ones to read; earlier ones warm caches/allocator.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import benchlib as bl  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--updates", type=int, default=6)
ap.add_argument("--config", default="redq")
ap.add_argument("--buffer", type=int, default=200_000)
ap.add_argument("--full-step", action="store_true", help="finish with one actor + alpha update")
args = ap.parse_args()
import super_sac_b200 as ssb  # noqa: E402

W = bl.Workload(ssb, args.config, torch.device("cuda", 0), buffer_size=args.buffer, fill_on_device=True)
rds = None
for k in range(args.updates):
    torch.cuda.nvtx.range_push(f"update{k}")
    W.step(k)
    torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print("done")
