#!/usr/bin/env python
"""Host-side breakdown of one end-to-end step (the `e2e` leg of bench.py): push, update call, Polyak, log fetch."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from super_sac_b200 import augmentations, graphed, learning, learning_utils as lu  # noqa: E402

cfg = dict(bench.CONFIGS["redq"])
cfg["buffer"] = 200_000
agent, target, critic_opt, enc_opt, log_alphas, buf = bench.build_gpu(cfg, torch.device("cuda", 0))
B = cfg["B"]
kw = dict(buffer=buf, agent=agent, target_agent=target, critic_optimizer=critic_opt, encoder_optimizer=enc_opt,
          log_alphas=log_alphas, batch_size=B, gamma=0.99, critic_clip=None, encoder_clip=None,
          target_critic_ensemble_n=cfg["M"], weighted_bellman_temp=None, weight_type=None, pop=False,
          augmenter=augmentations.AugmentationSequence([augmentations.IdentityAug(B)]), encoder_lambda=0.0,
          random_process=None, noise_clip=None, aug_mix=0.0)
hs, ha, hr, hs1, hd = bench.synthetic_transitions(cfg, 4096, seed=1)
graphed.enable_auto_graphs(True)
acc = {"push": 0.0, "update call (graph launch + log fetch)": 0.0, "polyak": 0.0}
N = 2000


def step(k, rec):
    j = k % 4096
    t0 = time.perf_counter()
    buf.push({"obs": hs[j]}, ha[j], float(hr[j]), {"obs": hs1[j]}, bool(hd[j]))
    t1 = time.perf_counter()
    logs, _ = learning.critic_update(**kw)
    t2 = time.perf_counter()
    if k % cfg["target_delay"] == 0:
        for ac, tc in zip(agent.critics, target.critics):
            lu.soft_update(tc, ac, cfg["tau"])
    t3 = time.perf_counter()
    if rec:
        acc["push"] += t1 - t0
        acc["update call (graph launch + log fetch)"] += t2 - t1
        acc["polyak"] += t3 - t2


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
for k in range(500):
    step(k, False)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
