#!/usr/bin/env python
"""Eager (un-graphed) REDQ-10 update steps for profiling under ncu:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python tools/profile_step.py --updates 6
The last 2 updates (one with, one without the Polyak step) are the ones to read; earlier ones warm caches/allocator.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--updates", type=int, default=6)
ap.add_argument("--config", default="redq")
ap.add_argument("--buffer", type=int, default=200_000)
args = ap.parse_args()
cfg = dict(bench.CONFIGS[args.config])
cfg["buffer"] = args.buffer
from super_sac_b200 import augmentations, learning, learning_utils as lu  # noqa: E402

agent, target, critic_opt, enc_opt, log_alphas, buf = bench.build_gpu(cfg, torch.device("cuda", 0))
B = cfg["B"]
kw = dict(buffer=buf, agent=agent, target_agent=target, critic_optimizer=critic_opt, encoder_optimizer=enc_opt,
          log_alphas=log_alphas, batch_size=B, gamma=0.99, critic_clip=None, encoder_clip=None,
          target_critic_ensemble_n=cfg["M"], weighted_bellman_temp=None, weight_type=None, pop=False,
          augmenter=augmentations.AugmentationSequence([augmentations.IdentityAug(B)]), encoder_lambda=0.0,
          random_process=None, noise_clip=None, aug_mix=0.0)
for k in range(args.updates):
    torch.cuda.nvtx.range_push(f"update{k}")
    logs, _ = learning.critic_update(**kw)
    if k % cfg["target_delay"] == 0:
        for ac, tc in zip(agent.critics, target.critics):
            lu.soft_update(tc, ac, cfg["tau"])
    torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print("done", logs["losses/critic_overall_loss"])
