#!/usr/bin/env python
"""SAC-Discrete step timed on the device: critic_update(discrete=True) + Polyak + online_actor_update +
alpha_update through the drop-in API, eager launches (this path is not graph-captured yet), CUDA events around K steps.
Atari-like head sizes: 64 features, H = 256, 18 actions, 2 critics, B = 256.  With --ref the UNMODIFIED reference
(baseline/_ref) runs the same step through PyTorch-CUDA on the same GPU, as a labelled comparison.

    python tools/bench_discrete.py --steps 300 --warmup 20 --ref > gpurun_out/r2_27_bench_discrete.json
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_27_discrete_launches.csv \
        python tools/bench_discrete.py --steps 2 --warmup 2
"""
import argparse
import copy
import json
import math
import os
import sys
from itertools import chain

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

S, A, H, N, M, B, NBUF = 64, 18, 256, 2, 2, 256, 100_000


def build(pkg, device, ours):
    class IdentityEncoder(pkg.nets.Encoder):
        def __init__(self):
            super().__init__()

        @property
        def embedding_dim(self):
            return S

        def forward(self, obs):
            return obs["obs"]

    torch.manual_seed(0)
    agent = pkg.Agent(act_space_size=A, encoder=IdentityEncoder(), actor_network_cls=pkg.nets.mlps.DiscreteActor,
                      critic_network_cls=pkg.nets.mlps.DiscreteCritic, discrete=True, ensemble_size=1, num_critics=N,
                      hidden_size=H, auto_rescale_targets=False)
    agent.to(device)
    target = copy.deepcopy(agent)
    target.to(device)
    rng = np.random.default_rng(0)
    buf = pkg.replay.ReplayBuffer(NBUF + 8, **(dict(device=device) if ours else {}))
    buf.load_experience({"obs": rng.standard_normal((NBUF, S)).astype(np.float32)},
                        rng.integers(0, A, size=(NBUF, 1)).astype(np.float32), rng.standard_normal(NBUF).astype(np.float32),
                        {"obs": rng.standard_normal((NBUF, S)).astype(np.float32)},
                        (rng.uniform(size=NBUF) < 0.05).astype(np.float32))
    critic_opt = torch.optim.Adam(chain(*(c.parameters() for c in agent.critics)), lr=3e-4)
    actor_opt = torch.optim.Adam(chain(*(a.parameters() for a in agent.actors)), lr=3e-4)
    enc_opt = torch.optim.Adam(agent.encoder.parameters(), lr=1e-4)
    la = torch.Tensor([math.log(0.1)]).to(device)
    la.requires_grad = True
    alpha_opt = torch.optim.Adam([la], lr=1e-4, betas=(0.5, 0.999))
    aug = pkg.augmentations.AugmentationSequence([pkg.augmentations.IdentityAug(B)])
    te = -math.log(1.0 / A) * 0.98
    L, lu = pkg.learning, pkg.learning_utils

    def step(k):
        logs, rds = L.critic_update(
            buffer=buf, agent=agent, target_agent=target, critic_optimizer=critic_opt, encoder_optimizer=enc_opt,
            log_alphas=[la], batch_size=B, gamma=0.99, critic_clip=None, encoder_clip=None, target_critic_ensemble_n=M,
            weighted_bellman_temp=None, weight_type=None, pop=False, augmenter=aug, encoder_lambda=0.0, aug_mix=0.0,
            discrete=True, random_process=None, noise_clip=None, per=False, update_priorities=False, dr3_coeff=0.0)
        if k % 2 == 0:
            for ac, tc in zip(agent.critics, target.critics):
                lu.soft_update(tc, ac, 0.005)
        alogs = L.online_actor_update(
            buffer=buf, agent=agent, pop=False, actor_optimizer=actor_opt, log_alphas=[la], batch_size=B, clip=None,
            random_process=None, noise_clip=None, augmenter=aug, aug_mix=0.0, premade_replay_dicts=rds, per=False,
            discrete=True, use_baseline=False)
        L.alpha_update(buffer=buf, agent=agent, optimizers=[alpha_opt], batch_size=B, log_alphas=[la], augmenter=aug,
                       aug_mix=0.0, target_entropy=te, premade_replay_dicts=rds, discrete=True)
        return float(logs["losses/critic_overall_loss"]), float(alogs["losses/actor_pg_loss"])

    return step


def timed(step, steps, warmup):
    for k in range(warmup):
        step(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        out = step(k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--ref", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    import super_sac_b200 as ssb

    ms, out = timed(build(ssb, dev, True), args.steps, args.warmup)
    line = {"metric": "sac_discrete_steps_per_sec", "unit": "steps/s", "value": 1e3 / ms, "ms_per_step": ms,
            "config": {"workload": f"SAC-Discrete step (critic + Polyak/2 + actor + alpha), S={S} H={H} A={A} N={N} B={B}",
                       "mode": "eager launches, logs read back every step"},
            "steps": args.steps, "warmup": args.warmup, "dtype": "f32", "last_losses": out}
    if args.ref:
        from baseline import ref_import

        if ref_import.available():
            ref = ref_import.import_reference(device="cuda")
            rms, rout = timed(build(ref, "cuda", False), max(20, args.steps // 6), 5)
            line["reference_on_this_gpu_pytorch_cuda"] = {"ms_per_step": rms, "value": 1e3 / rms, "last_losses": rout}
    print(json.dumps(line))
