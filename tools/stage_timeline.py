#!/usr/bin/env python
"""Where the time of one graph-replayed REDQ-10 critic update goes: timing events recorded between the stages of
learning.critic_update INSIDE the captured graph (each event node costs ~1 us itself, so the stage times are slightly
inflated; the bench number is measured without them).
    python tools/stage_timeline.py [--config redq] [--replays 200]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="redq")
ap.add_argument("--replays", type=int, default=200)
ap.add_argument("--buffer", type=int, default=200_000)
args = ap.parse_args()
cfg = dict(bench.CONFIGS[args.config])
cfg["buffer"] = args.buffer
from super_sac_b200 import augmentations, graphed, learning, learning_utils as lu  # noqa: E402

agent, target, critic_opt, enc_opt, log_alphas, buf = bench.build_gpu(cfg, torch.device("cuda", 0))
B = cfg["B"]
kw = dict(buffer=buf, agent=agent, target_agent=target, critic_optimizer=critic_opt, encoder_optimizer=enc_opt,
          log_alphas=log_alphas, batch_size=B, gamma=0.99, critic_clip=None, encoder_clip=None,
          target_critic_ensemble_n=cfg["M"], weighted_bellman_temp=None, weight_type=None, pop=False,
          augmenter=augmentations.AugmentationSequence([augmentations.IdentityAug(B)]), encoder_lambda=0.0,
          random_process=None, noise_clip=None, aug_mix=0.0)


def step():
    lu._marks = [] if torch.cuda.is_current_stream_capturing() else None
    out = learning._critic_update_impl(**kw)
    for ac, tc in zip(agent.critics, target.critics):
        lu.soft_update(tc, ac, cfg["tau"])
    lu._mark("Polyak")
    return out


g = graphed.GraphedCall(step)
marks = lu._marks
lu._marks = None
sums = [0.0] * len(marks)
for _ in range(args.replays):
    g.replay()
    torch.cuda.synchronize()
    for i in range(1, len(marks)):
        sums[i] += marks[i - 1][1].elapsed_time(marks[i][1])
total = 0.0
for i in range(1, len(marks)):
    us = 1e3 * sums[i] / args.replays
    total += us
    print(f"{marks[i][0]:58s} {us:7.2f} us   (cumulative {total:7.2f})")
