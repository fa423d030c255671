#!/usr/bin/env python
"""Device timeline (CUPTI via torch.profiler) of the graph-replayed, software-pipelined UTD block: every kernel with its
stream, start and duration, for a few updates in the middle of the block.  Not a bench (the profiler adds overhead per
kernel); it shows which chain of kernels bounds the update.
    python tools/kernel_timeline.py [--config redq] [--updates 6] [--first 8]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import benchlib as bl  # noqa: E402
import super_sac_b200 as ssb  # noqa: E402
from super_sac_b200 import graphed, learning_utils as lu  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="redq")
ap.add_argument("--block", type=int, default=20)
ap.add_argument("--first", type=int, default=8, help="first update of the block to print")
ap.add_argument("--updates", type=int, default=3)
ap.add_argument("--no-pipeline", action="store_true")
ap.add_argument("--e2e", action="store_true", help="the end-to-end loop (push + drop-in critic_update + Polyak, lazy logs) instead of the block")
args = ap.parse_args()
W = bl.Workload(ssb, args.config, torch.device("cuda", 0), buffer_size=200_000, fill_on_device=True)
if args.e2e:
    tr = W.host_transitions(4096, seed=1)
    graphed.enable_auto_graphs(True, lazy_logs=True, pipeline=not args.no_pipeline)
    prev = {"logs": None}

    def step(k):
        W.push(tr, k % 4096)
        logs = W.step(k)[0]
        if prev["logs"] is not None:
            float(prev["logs"]["losses/critic_overall_loss"])
        prev["logs"] = logs

    for k in range(40):
        step(k)
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for k in range(400):
        step(k)
    torch.cuda.synchronize()
    print("e2e loop: %.1f us/step" % ((time.perf_counter() - t0) * 1e6 / 400))
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for k in range(24):
            step(k)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    adam = [e for e in evs if "adam_kernel" in e.name]
    print("Adam-end to Adam-end us:", " ".join("%.1f" % (adam[i + 1].time_range.end - adam[i].time_range.end) for i in range(len(adam) - 1)))
    lo, hi = adam[9].time_range.end, adam[12].time_range.end
    for e in evs:
        if e.time_range.end < lo - 30 or e.time_range.start > hi:
            continue
        name = e.name.split("(")[0].replace("void ", "").replace("ssac::", "")
        print("%9.1f %-7s %8.1f %7.1f  %s" % (e.time_range.start - t0, getattr(e, "device_resource_id", "?"), e.time_range.start - lo, e.time_range.end - e.time_range.start, name[:70]))
    sys.exit(0)


def blk():
    import contextlib
    with (contextlib.nullcontext() if args.no_pipeline else lu.pipelined_updates()):
        for u in range(args.block):
            out = W.step(u)
    return out


g = graphed.GraphedCall(blk, warmup=1)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    g.replay()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
evs.sort(key=lambda e: e.time_range.start)
if not evs:
    print("no CUDA kernel events (CUPTI unavailable?)")
    sys.exit(0)
t0 = evs[0].time_range.start
# updates are delimited by the rng draw kernel (first kernel of every update's target side)
starts = [i for i, e in enumerate(evs) if "rng_fill" in e.name]
print("kernels in block:", len(evs), " updates seen:", len(starts), " block time: %.1f us" % (evs[-1].time_range.end - t0))
adam = [e for e in evs if "first_layer_wgrad" in e.name]   # the last reduction of an update's backward
if len(adam) > 2:
    gaps = [adam[i + 1].time_range.end - adam[i].time_range.end for i in range(len(adam) - 1)]
    print("gW1-end to gW1-end (update period) us:", " ".join("%.1f" % x for x in gaps))
lo = adam[args.first - 1].time_range.end if len(adam) > args.first else t0
hi = adam[min(args.first + args.updates, len(adam)) - 1].time_range.end
print("%-9s %-7s %8s %7s  %s" % ("start", "stream", "t-rel", "dur", "kernel"))
for e in evs:
    if e.time_range.end < lo - 30 or e.time_range.start > hi:
        continue
    name = e.name.split("(")[0].replace("void ", "").replace("ssac::", "")
    print("%9.1f %-7s %8.1f %7.1f  %s" % (e.time_range.start - t0, getattr(e, "device_resource_id", "?"), e.time_range.start - lo, e.time_range.end - e.time_range.start, name[:70]))
