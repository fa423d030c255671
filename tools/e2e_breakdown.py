#!/usr/bin/env python
"""Host-side breakdown of one end-to-end step (the `e2e` leg of bench.py): push, update call, Polyak, log read."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import benchlib as bl  # noqa: E402
import super_sac_b200 as ssb  # noqa: E402
from super_sac_b200 import graphed  # noqa: E402

W = bl.Workload(ssb, "redq", torch.device("cuda", 0), buffer_size=200_000, fill_on_device=True)
tr = W.host_transitions(4096, seed=1)
graphed.enable_auto_graphs(True, lazy_logs=True, pipeline="--pipeline" in sys.argv)
acc = {"push": 0.0, "critic_update call": 0.0, "polyak": 0.0, "read previous logs": 0.0}
N = 3000
prev = {"logs": None}


def step(k, rec):
    t0 = time.perf_counter()
    W.push(tr, k % 4096)
    t1 = time.perf_counter()
    logs = W.critic_update()[0]
    t2 = time.perf_counter()
    if k % W.cfg["target_delay"] == 0:
        W.polyak()
    t3 = time.perf_counter()
    if prev["logs"] is not None:
        float(prev["logs"]["losses/critic_overall_loss"])
    prev["logs"] = logs
    t4 = time.perf_counter()
    if rec:
        acc["push"] += t1 - t0
        acc["critic_update call"] += t2 - t1
        acc["polyak"] += t3 - t2
        acc["read previous logs"] += t4 - t3


for k in range(10):
    step(k, False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(N):
    step(k, True)
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print(f"total {1e6 * tot / N:.1f} us/step")
for k, v in acc.items():
    print(f"  {k:44s} {1e6 * v / N:7.1f} us")
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for k in range(1000):
    step(k, False)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(30)
