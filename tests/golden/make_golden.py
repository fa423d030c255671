"""Generate the committed golden fixtures by running the UNMODIFIED reference with injected randomness.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Writes tests/golden/*.npz.  Each fixture stores the initial learnable state, the replay contents, every
injected random draw and the reference's outputs (TD targets, Bellman weights, gradients, post-step
parameters, post-Polyak targets, logged scalars) so that the oracle (CPU) and the CUDA path (GPU box,
where the reference does not exist) can be checked against the reference itself.
"""
import copy
import math
import os
import sys
from itertools import chain

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

ssac = rh.import_reference()
from super_sac import learning, learning_utils as lu, replay as rreplay, augmentations as raug, nets as rnets  # noqa: E402

torch.set_num_threads(1)


class IdentityEncoder(rnets.Encoder):
    def __init__(self, dim):
        super().__init__()
        self._dim = dim

    @property
    def embedding_dim(self):
        return self._dim

    def forward(self, obs_dict):
        return obs_dict["obs"]


class SharedEncoder(rnets.Encoder):
    """Trainable state encoder in the style of experiments/gym/train_gym.py:31-45."""

    def __init__(self, dim, hid=16):
        super().__init__()
        self.fc0 = torch.nn.Linear(dim, hid)
        self.fc1 = torch.nn.Linear(hid, dim)
        self._dim = dim

    @property
    def embedding_dim(self):
        return self._dim

    def forward(self, obs_dict):
        x = torch.relu(self.fc0(obs_dict["obs"]))
        return torch.relu(self.fc1(x))


def stack_of(modules):
    last = lambda m: m.out if hasattr(m, "out") else m.fc3
    return {
        "W1": np.stack([m.fc1.weight.detach().numpy().copy() for m in modules]),
        "b1": np.stack([m.fc1.bias.detach().numpy().copy() for m in modules]),
        "W2": np.stack([m.fc2.weight.detach().numpy().copy() for m in modules]),
        "b2": np.stack([m.fc2.bias.detach().numpy().copy() for m in modules]),
        "W3": np.stack([last(m).weight.detach().numpy().copy() for m in modules]),
        "b3": np.stack([last(m).bias.detach().numpy().copy() for m in modules]),
    }


def grads_of(modules):
    last = lambda m: m.out if hasattr(m, "out") else m.fc3
    g = lambda p: p.grad.detach().numpy().copy()
    return {
        "W1": np.stack([g(m.fc1.weight) for m in modules]),
        "b1": np.stack([g(m.fc1.bias) for m in modules]),
        "W2": np.stack([g(m.fc2.weight) for m in modules]),
        "b2": np.stack([g(m.fc2.bias) for m in modules]),
        "W3": np.stack([g(last(m).weight) for m in modules]),
        "b3": np.stack([g(last(m).bias) for m in modules]),
    }


def critic_nets(agent):
    return [net for c in agent.critics for net in c.nets]


def put(d, prefix, sub):
    for k, v in sub.items():
        d[f"{prefix}/{k}"] = np.array(v, copy=True)


def popart_state(agent):
    out = {}
    for i, p in enumerate(agent.popart):
        if p:
            out[f"{i}/mu"], out[f"{i}/nu"] = p.mu.numpy().copy(), p.nu.numpy().copy()
            out[f"{i}/w"], out[f"{i}/b"] = p.w.numpy().copy(), p.b.numpy().copy()
            out[f"{i}/t"] = np.array(p._t)
            out[f"{i}/stable"] = np.array(int(p._stable))
    return out


def run_update_case(name, cfg, out_dir=None):
    """Drives critic_update (+Polyak by the target_delay rule of main.py:409), then online_actor_update and
    alpha_update, exactly as main.py:380-414 / :491-543 call them."""
    rng = np.random.default_rng(cfg.get("seed", 0))
    torch.manual_seed(cfg.get("seed", 0))
    E, N, M = cfg["E"], cfg["N"], cfg["M"]
    S, A, H, B = cfg["S"], cfg["A"], cfg["H"], cfg["B"]
    det = cfg.get("deterministic", False)
    popart_on = cfg.get("popart", False)
    enc = SharedEncoder(S) if cfg.get("encoder") == "shared" else IdentityEncoder(S)
    agent = ssac.Agent(
        act_space_size=A, encoder=enc,
        actor_network_cls=rnets.mlps.ContinuousDeterministicActor if det else rnets.mlps.ContinuousStochasticActor,
        critic_network_cls=rnets.mlps.ContinuousCritic,
        ensemble_size=E, num_critics=N, hidden_size=H, auto_rescale_targets=popart_on,
        log_std_low=-5.0, log_std_high=2.0,
    )
    # give the biases some mass so that every gradient path is exercised from step 0
    for m in critic_nets(agent) + list(agent.actors):
        for p in m.parameters():
            if p.dim() == 1:
                p.data.add_(0.05 * torch.randn_like(p))
    if popart_on and cfg.get("popart_warm", False):
        for p in agent.popart:
            p._t = 1500
            p.mu = torch.tensor([0.3])
            p.nu = torch.tensor([1.7])
            p.w = torch.tensor([0.9])
            p.b = torch.tensor([0.1])
    target = copy.deepcopy(agent)
    # main.py:322-325 hard updates: identical at creation; perturb the target so Polyak is non-trivial
    for m in critic_nets(target):
        for p in m.parameters():
            p.data.add_(0.01 * torch.randn_like(p))

    nbuf = cfg.get("nbuf", 64)
    s = rng.standard_normal((nbuf, S)).astype(np.float32)
    a = rng.uniform(-1, 1, (nbuf, A)).astype(np.float32)
    r = rng.standard_normal((nbuf,)).astype(np.float32)
    s1 = rng.standard_normal((nbuf, S)).astype(np.float32)
    d = (rng.uniform(size=(nbuf,)) < 0.1).astype(np.float32)
    buffer = rreplay.ReplayBuffer(size=nbuf + 8)
    buffer.load_experience({"obs": s}, a, r, {"obs": s1}, d)

    out = {}
    out["cfg"] = np.array(repr(cfg))
    put(out, "buffer", dict(s=s, a=a, r=r, s1=s1, d=d))
    put(out, "init/actors", stack_of(agent.actors))
    put(out, "init/critics", stack_of(critic_nets(agent)))
    put(out, "init/target_critics", stack_of(critic_nets(target)))
    put(out, "init/popart", popart_state(agent))
    if cfg.get("encoder") == "shared":
        put(out, "init/encoder", {k: v.numpy().copy() for k, v in enc.state_dict().items()})
        put(out, "init/target_encoder", {k: v.numpy().copy() for k, v in target.encoder.state_dict().items()})

    critic_opt = torch.optim.Adam(chain(*(c.parameters() for c in agent.critics)), lr=cfg.get("critic_lr", 3e-4),
                                  weight_decay=cfg.get("critic_l2", 0.0), betas=(0.9, 0.999))
    actor_opt = torch.optim.Adam(chain(*(ac.parameters() for ac in agent.actors)), lr=cfg.get("actor_lr", 3e-4),
                                 betas=(0.9, 0.999))
    enc_opt = torch.optim.Adam(agent.encoder.parameters(), lr=cfg.get("encoder_lr", 1e-4), betas=(0.9, 0.999))
    init_alpha = max(cfg.get("init_alpha", 0.1), 1e-15)
    log_alphas, alpha_opts = [], []
    for _ in range(E):
        la = torch.Tensor([math.log(init_alpha)])
        la.requires_grad = True
        log_alphas.append(la)
        alpha_opts.append(torch.optim.Adam([la], lr=cfg.get("alpha_lr", 1e-4), betas=(0.5, 0.999)))

    sigma = cfg.get("noise_sigma")
    random_process = None
    if sigma is not None:
        random_process = lu.GaussianExplorationNoise(rh.ActionSpace(A), start_scale=sigma, final_scale=min(sigma, 0.1))
    augmenter = raug.AugmentationSequence([raug.IdentityAug(B)])
    gamma = cfg.get("gamma", 0.99)
    wt, wtemp = cfg.get("weight_type"), cfg.get("weight_temp")
    softmax_w = wt == "softmax" and E > 1

    # record intermediates without touching the reference's code
    rec = {}
    o_td, o_bw = lu.compute_td_targets, lu.compute_backup_weights

    def td_rec(*a_, **k_):
        res = o_td(*a_, **k_)
        rec.setdefault("td", []).append(res[0].detach().numpy().copy())
        rec.setdefault("a1", []).append(res[1][1].detach().numpy().copy())
        return res

    def bw_rec(*a_, **k_):
        res = o_bw(*a_, **k_)
        rec.setdefault("w", []).append(res.detach().numpy().copy() if torch.is_tensor(res) else np.array(res, dtype=np.float32))
        return res

    lu.compute_td_targets, lu.compute_backup_weights = td_rec, bw_rec
    try:
        replay_dicts = None
        for t in range(cfg["steps"]):
            idx = rng.integers(0, nbuf, size=(E, B))
            subsets = [list(rng.permutation(N)[:M]) for _ in range(E)]
            eps = rng.standard_normal((E, B, A)).astype(np.float32)
            noise = rng.standard_normal((E, B, A)).astype(np.float32)
            weps = rng.standard_normal((E, E, B, A)).astype(np.float32)
            put(out, f"step{t}/rand", dict(idx=idx, subsets=np.array(subsets), eps=eps, noise=noise, weight_eps=weps))
            # draw order inside one member iteration (learning.py:47-80): indices, [actor eps], [td3 noise],
            # subset, then for softmax weights one actor sample per ensemble member
            normal_q, randn_q = [], []
            for i in range(E):
                if not det:
                    normal_q.append(eps[i])
                if sigma is not None:
                    randn_q.append(noise[i])
                if softmax_w and not det:
                    normal_q.extend(list(weps[i]))
            rec.clear()
            with rh.injected(randint=list(idx), subsets=subsets, normal_eps=normal_q, randn=randn_q) as q:
                logs, replay_dicts = learning.critic_update(
                    buffer=buffer, agent=agent, target_agent=target, critic_optimizer=critic_opt,
                    encoder_optimizer=enc_opt, log_alphas=log_alphas, batch_size=B, gamma=gamma,
                    critic_clip=cfg.get("critic_clip"), encoder_clip=cfg.get("encoder_clip"),
                    target_critic_ensemble_n=M, weighted_bellman_temp=wtemp, weight_type=wt,
                    pop=cfg.get("pop", False), augmenter=augmenter, encoder_lambda=0.0, aug_mix=0.0,
                    discrete=False, random_process=random_process, noise_clip=cfg.get("noise_clip"),
                    per=False, update_priorities=False, dr3_coeff=cfg.get("dr3_coeff", 0.0))
                assert not q["randint"].items and not q["normal_eps"].items and not q["randn"].items
            put(out, f"step{t}/td_target", {str(i): v for i, v in enumerate(rec["td"])})
            put(out, f"step{t}/a1", {str(i): v for i, v in enumerate(rec["a1"])})
            put(out, f"step{t}/weights", {str(i): v for i, v in enumerate(rec["w"])})
            put(out, f"step{t}/critic_grads", grads_of(critic_nets(agent)))
            put(out, f"step{t}/logs", {k.replace("/", "|"): float(v) for k, v in logs.items()})
            if (t + cfg.get("step0", 0)) % cfg.get("target_delay", 1) == 0:
                for ac, tc in zip(agent.critics, target.critics):
                    lu.soft_update(tc, ac, cfg.get("tau", 0.005))
                lu.soft_update(target.encoder, agent.encoder, cfg.get("encoder_tau", 0.01))
            put(out, f"step{t}/critics", stack_of(critic_nets(agent)))
            put(out, f"step{t}/target_critics", stack_of(critic_nets(target)))
            put(out, f"step{t}/popart", popart_state(agent))
            if cfg.get("encoder") == "shared":
                put(out, f"step{t}/encoder", {k: v.numpy().copy() for k, v in enc.state_dict().items()})
                put(out, f"step{t}/target_encoder", {k: v.numpy().copy() for k, v in target.encoder.state_dict().items()})
    finally:
        lu.compute_td_targets, lu.compute_backup_weights = o_td, o_bw

    # actor + alpha update on the last critic batch (reuse_replay_dicts=True, main.py:491-543)
    eps = rng.standard_normal((E, B, A)).astype(np.float32)
    noise = rng.standard_normal((E, B, A)).astype(np.float32)
    put(out, "actor/rand", dict(eps=eps, noise=noise))
    with rh.injected(normal_eps=list(eps), randn=list(noise) if sigma is not None else []):
        # (the deterministic actor's rsample() is Normal(loc, 1e-4).rsample(): it draws too)
        alogs = learning.online_actor_update(
            buffer=buffer, agent=agent, pop=cfg.get("pop", False), actor_optimizer=actor_opt, log_alphas=log_alphas,
            batch_size=B, clip=cfg.get("actor_clip"), random_process=random_process, noise_clip=cfg.get("noise_clip"),
            augmenter=augmenter, aug_mix=0.0, premade_replay_dicts=replay_dicts, per=False, discrete=False,
            use_baseline=False)
    put(out, "actor/grads", grads_of(agent.actors))
    put(out, "actor/actors", stack_of(agent.actors))
    put(out, "actor/logs", {k.replace("/", "|"): float(v) for k, v in alogs.items()})
    if cfg.get("alpha_update", True):
        eps = rng.standard_normal((E, B, A)).astype(np.float32)
        put(out, "alpha/rand", dict(eps=eps))
        with rh.injected(normal_eps=[] if det else list(eps)):
            llogs = learning.alpha_update(
                buffer=buffer, agent=agent, optimizers=alpha_opts, batch_size=B, log_alphas=log_alphas,
                augmenter=augmenter, aug_mix=0.0, target_entropy=-float(A), premade_replay_dicts=replay_dicts,
                discrete=False)
        put(out, "alpha/log_alphas", {str(i): la.detach().numpy().copy() for i, la in enumerate(log_alphas)})
        put(out, "alpha/logs", {k.replace("/", "|"): float(v) for k, v in llogs.items()})
    path = os.path.join(out_dir or HERE, f"update_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}  ({os.path.getsize(path)/1024:.1f} KiB)")


UPDATE_CASES = {
    # C1-shaped: plain SAC, 2 critics
    "sac": dict(E=1, N=2, M=2, S=3, A=1, H=32, B=16, steps=3, target_delay=2, seed=1),
    # C2-shaped: REDQ, random subset of 2 target critics out of N
    "redq": dict(E=1, N=5, M=2, S=17, A=6, H=48, B=32, steps=2, target_delay=2, seed=2),
    # C3-shaped: SUNRISE ensemble with weighted Bellman backups + PopArt (warm, so the POP branch is live)
    "sunrise_popart": dict(E=3, N=2, M=2, S=5, A=2, H=32, B=16, steps=3, weight_type="sunrise", weight_temp=20.0,
                           popart=True, pop=True, popart_warm=True, seed=3),
    # C4-shaped learner: deterministic actor + TD3 noise + trainable encoder + clipping
    "td3_encoder": dict(E=1, N=2, M=2, S=6, A=3, H=32, B=16, steps=2, deterministic=True, noise_sigma=0.5,
                        noise_clip=0.3, encoder="shared", critic_clip=0.05, encoder_clip=40.0, actor_clip=40.0,
                        tau=0.01, encoder_tau=1.0, gamma=0.99 ** 3, init_alpha=0.0, alpha_update=False, seed=4),
    # C5-shaped extras: softmax weights + DR3 + L2 + cold PopArt
    "softmax_dr3": dict(E=2, N=2, M=1, S=4, A=2, H=32, B=16, steps=2, weight_type="softmax", weight_temp=10.0,
                        dr3_coeff=0.01, critic_clip=40.0, critic_l2=1e-3, popart=True, pop=True, seed=5),
}


def run_replay_case():
    """ReplayBuffer ring + PER trees: replay.py:10-353."""
    rng = np.random.default_rng(11)
    out = {}
    size, S, A = 50, 4, 2
    buf = rreplay.ReplayBuffer(size=size, alpha=0.6, beta=0.7)
    ops = []
    # single pushes, batched pushes with wrap-around, priority pushes
    n_ops = 0
    for t in range(30):
        s, a = rng.standard_normal(S).astype(np.float32), rng.uniform(-1, 1, A).astype(np.float32)
        r, s1, d = float(rng.standard_normal()), rng.standard_normal(S).astype(np.float32), bool(rng.uniform() < 0.2)
        buf.push({"obs": s}, a, r, {"obs": s1}, d)
        put(out, f"op{n_ops}", dict(kind="push1", s=s, a=a, r=r, s1=s1, d=d)); n_ops += 1
    for t in range(3):
        n = 12
        s, a = rng.standard_normal((n, S)).astype(np.float32), rng.uniform(-1, 1, (n, A)).astype(np.float32)
        r, s1 = rng.standard_normal((n, 1)).astype(np.float32), rng.standard_normal((n, S)).astype(np.float32)
        d = (rng.uniform(size=(n, 1)) < 0.2)
        pr = rng.uniform(0.1, 3.0, n)
        buf.push({"obs": s}, a, r, {"obs": s1}, d, priorities=pr)
        put(out, f"op{n_ops}", dict(kind="pushN", s=s, a=a, r=r, s1=s1, d=d, priorities=pr)); n_ops += 1
    st = buf._storage
    put(out, "after_push", dict(s=st.s_stack["obs"], s1=st.s1_stack["obs"], a=st.action_stack, r=st.reward_stack,
                                d=st.done_stack, next_idx=st._next_idx, filled=st._max_filled,
                                sum_tree=buf._it_sum._value, min_tree=buf._it_min._value, max_priority=buf._max_priority))
    # uniform sample with injected indices
    idx = rng.integers(0, len(buf), size=8)
    with rh.injected(randint=[idx]):
        (s_, a_, r_, s1_, d_), ridx = buf.sample_uniform(8)
    put(out, "uniform", dict(idx=idx, s=s_["obs"].numpy(), a=a_.numpy(), r=r_.numpy(), s1=s1_["obs"].numpy(), d=d_.numpy(), ridx=ridx))
    # PER sample / update rounds (duplicates in idxes are likely with B=16 over 50 slots)
    for t in range(4):
        u = rng.uniform(size=16)
        with rh.injected(np_random=[u]):
            (s_, a_, r_, s1_, d_), w, idxes = buf.sample(16)
        newp = rng.uniform(1e-3, 5.0, 16)
        buf.update_priorities(idxes, newp)
        put(out, f"per{t}", dict(u=u, idxes=idxes, weights=w.numpy(), s=s_["obs"].numpy(), a=a_.numpy(), new_priorities=newp,
                                 sum_tree=buf._it_sum._value, min_tree=buf._it_min._value, max_priority=buf._max_priority))
    out["n_ops"] = np.array(n_ops)
    path = os.path.join(HERE, "replay_per.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}  ({os.path.getsize(path)/1024:.1f} KiB)")


def run_aug_case():
    """DrQ / DrQv2 augmentation + sample_move_and_augment on a uint8 pixel buffer."""
    rng = np.random.default_rng(21)
    out = {}
    B, C, HW, nbuf, A = 8, 3, 20, 24, 2
    s = rng.integers(0, 256, (nbuf, C, HW, HW), dtype=np.uint8)
    s1 = rng.integers(0, 256, (nbuf, C, HW, HW), dtype=np.uint8)
    a = rng.uniform(-1, 1, (nbuf, A)).astype(np.float32)
    r = rng.standard_normal(nbuf).astype(np.float32)
    d = (rng.uniform(size=nbuf) < 0.1)
    buf = rreplay.ReplayBuffer(size=nbuf)
    buf.load_experience({"pixels": s}, a, r, {"pixels": s1}, d)
    put(out, "buffer", dict(s=s, s1=s1, a=a, r=r, d=d))
    idx = rng.integers(0, nbuf, B)
    # --- DrQv2 (bilinear in the reference) ---
    shift = rng.integers(0, 9, (B, 1, 1, 2))
    aug = raug.AugmentationSequence([raug.Drqv2Aug(B)])
    with rh.injected(randint=[idx, shift]):
        rd = lu.sample_move_and_augment(buf, B, aug, aug_mix=0.75, per=False)
    o, a_, r_, o1, d_ = rd["primary_batch"]
    put(out, "drqv2", dict(idx=idx, shift=shift, o=o["pixels"].numpy(), o1=o1["pixels"].numpy(),
                           ao=rd["augmented_obs"][0]["pixels"].numpy(), ao1=rd["augmented_obs"][1]["pixels"].numpy(),
                           oo=rd["original_obs"][0]["pixels"].numpy(), a=a_.numpy(), r=r_.numpy(), d=d_.numpy()))
    # --- DrQ v1 without noise (exact integer crop in the reference) ---
    w1, h1 = rng.integers(0, 8, B), rng.integers(0, 8, B)
    with rh.injected(randint=[raug_dummy for raug_dummy in (w1, h1)]):
        augobj = raug.DrqNoNoiseAug(B)  # constructor draws once
    aug = raug.AugmentationSequence([augobj])
    with rh.injected(randint=[idx, w1, h1]):
        rd = lu.sample_move_and_augment(buf, B, aug, aug_mix=1.0, per=False)
    o, _, _, o1, _ = rd["primary_batch"]
    put(out, "drqv1", dict(idx=idx, w1=w1, h1=h1, o=o["pixels"].numpy(), o1=o1["pixels"].numpy()))
    # --- DrQ v1 with N(0,1) noise ---
    with rh.injected(randint=[w1, h1]):
        augobj = raug.DrqAug(B)
    aug = raug.AugmentationSequence([augobj])
    n0 = rng.standard_normal((B, C, HW, HW)).astype(np.float32)
    n1 = rng.standard_normal((B, C, HW, HW)).astype(np.float32)
    with rh.injected(randint=[idx, w1, h1], randn_like=[n0, n1]):
        rd = lu.sample_move_and_augment(buf, B, aug, aug_mix=0.5, per=False)
    o, _, _, o1, _ = rd["primary_batch"]
    put(out, "drqv1_noise", dict(idx=idx, w1=w1, h1=h1, n0=n0, n1=n1, o=o["pixels"].numpy(), o1=o1["pixels"].numpy()))
    path = os.path.join(HERE, "aug_pixels.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}  ({os.path.getsize(path)/1024:.1f} KiB)")


def run_rad_case():
    """RadAug (augmentations.py:129-162): cv2 bilinear upscale by ``crop`` pixels + random integer crop, through
    sample_move_and_augment on uint8 frames.  Two layouts: a 9-channel frame stack (cv2's generic-channel path) and a
    3-channel image (cv2's <=4-channel path, which rounds differently)."""
    rng = np.random.default_rng(33)
    out = {}
    for tag, (B, C, HW, nbuf, crop, mix) in {"stack9": (4, 9, 36, 6, 16, 1.0), "rgb3": (4, 3, 20, 6, 5, 0.5)}.items():
        A = 2
        s = rng.integers(0, 256, (nbuf, C, HW, HW), dtype=np.uint8)
        s1 = rng.integers(0, 256, (nbuf, C, HW, HW), dtype=np.uint8)
        a = rng.uniform(-1, 1, (nbuf, A)).astype(np.float32)
        r = rng.standard_normal(nbuf).astype(np.float32)
        d = (rng.uniform(size=nbuf) < 0.1)
        buf = rreplay.ReplayBuffer(size=nbuf)
        buf.load_experience({"pixels": s}, a, r, {"pixels": s1}, d)
        idx = rng.integers(0, nbuf, B)
        h, w = rng.integers(0, crop, B), rng.integers(0, crop, B)
        with rh.injected(randint=[h, w]):
            augobj = raug.RadAug(B, crop=crop)   # constructor draws once
        aug = raug.AugmentationSequence([augobj])
        with rh.injected(randint=[idx, h, w]):
            rd = lu.sample_move_and_augment(buf, B, aug, aug_mix=mix, per=False)
        o, _, _, o1, _ = rd["primary_batch"]
        put(out, tag, dict(s=s, s1=s1, a=a, r=r, d=d, idx=idx, h=h, w=w, crop=np.int64(crop), mix=np.float64(mix),
                           o=o["pixels"].numpy(), o1=o1["pixels"].numpy()))
    path = os.path.join(HERE, "aug_rad.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}  ({os.path.getsize(path)/1024:.1f} KiB)")


def run_afbc_case():
    """Offline AFBC actor update with prioritised sampling and priority refresh (learning.py:144-219,
    learning_utils.py:241-295, adv_estimator.py:58-79, replay.py:163-190) -- BASELINE config 5's actor side."""
    rng = np.random.default_rng(31)
    torch.manual_seed(31)
    E, N, S, A, H, B, nbuf = 2, 2, 5, 2, 32, 16, 48
    agent = ssac.Agent(act_space_size=A, encoder=IdentityEncoder(S), actor_network_cls=rnets.mlps.ContinuousStochasticActor,
                       critic_network_cls=rnets.mlps.ContinuousCritic, ensemble_size=E, num_critics=N, hidden_size=H,
                       auto_rescale_targets=True, log_std_low=-5.0, log_std_high=2.0)
    for m in critic_nets(agent) + list(agent.actors):
        for p_ in m.parameters():
            if p_.dim() == 1:
                p_.data.add_(0.05 * torch.randn_like(p_))
    for pa in agent.popart:
        pa.w = torch.tensor([0.8]); pa.b = torch.tensor([-0.2])
    s = rng.standard_normal((nbuf, S)).astype(np.float32)
    a = rng.uniform(-1, 1, (nbuf, A)).astype(np.float32)
    a[::7] = np.sign(a[::7])  # some actions exactly at +-1: exercises the clamp(+-0.99) of the atanh cache miss
    r = rng.standard_normal(nbuf).astype(np.float32)
    s1 = rng.standard_normal((nbuf, S)).astype(np.float32)
    d = (rng.uniform(size=nbuf) < 0.1)
    pri = rng.uniform(0.05, 2.0, nbuf)
    buf = rreplay.ReplayBuffer(size=64, alpha=0.6, beta=1.0)
    buf.push({"obs": s}, a, r[:, None], {"obs": s1}, d[:, None], priorities=pri)
    out = {"cfg": np.array(repr(dict(E=E, N=N, S=S, A=A, H=H, B=B, nbuf=nbuf, popart=True)))}
    put(out, "buffer", dict(s=s, a=a, r=r, s1=s1, d=d, priorities=pri))
    put(out, "init/actors", stack_of(agent.actors))
    put(out, "init/critics", stack_of(critic_nets(agent)))
    put(out, "init/popart", popart_state(agent))
    actor_opt = torch.optim.Adam(chain(*(ac.parameters() for ac in agent.actors)), lr=3e-4, betas=(0.9, 0.999))
    enc_opt = torch.optim.Adam(agent.encoder.parameters(), lr=1e-4)
    augmenter = raug.AugmentationSequence([raug.IdentityAug(B)])
    for t in range(2):
        u = rng.uniform(size=(E, B))
        adv_eps = rng.standard_normal((E, 4, B, A)).astype(np.float32)
        prio_member = int(rng.integers(0, E))
        prio_eps = rng.standard_normal((4, B, A)).astype(np.float32)
        put(out, f"step{t}/rand", dict(u=u, adv_eps=adv_eps, prio_member=prio_member, prio_eps=prio_eps))
        normal_q = [e for i in range(E) for e in adv_eps[i]] + list(prio_eps)
        o_choice = rh._py_random.choice
        rh._py_random.choice = lambda seq: (list(seq)[prio_member] if isinstance(seq, range) else o_choice(seq))
        try:
            with rh.injected(np_random=list(u), normal_eps=normal_q) as q:
                logs = learning.offline_actor_update(
                    buffer=buf, agent=agent, actor_optimizer=actor_opt, encoder_optimizer=enc_opt, batch_size=B,
                    actor_clip=40.0, update_encoder=False, encoder_clip=40.0, augmenter=augmenter, actor_lambda=0.0,
                    aug_mix=0.0, premade_replay_dicts=None, per=True, discrete=False, filter_=True)
                assert not q["normal_eps"].items and not q["np_random"].items
        finally:
            rh._py_random.choice = o_choice
        put(out, f"step{t}/actor_grads", grads_of(agent.actors))
        put(out, f"step{t}/actors", stack_of(agent.actors))
        put(out, f"step{t}/logs", {k.replace("/", "|"): float(v) for k, v in logs.items()})
        put(out, f"step{t}/trees", dict(sum_tree=buf._it_sum._value, min_tree=buf._it_min._value, max_priority=buf._max_priority))
    path = os.path.join(HERE, "afbc.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}  ({os.path.getsize(path)/1024:.1f} KiB)")


def run_encoder_case():
    """BigPixelEncoder (nets/cnns.py:37-69): forward output and autograd gradients of every parameter, from the
    unmodified reference module, on two small image geometries (one odd valid size, one with 9 channels)."""
    out = {}
    for tag, (B, C, HW, O, seed) in {"rgb20": (5, 3, 20, 10, 41), "stack24": (3, 9, 24, 50, 42)}.items():
        rng = np.random.default_rng(seed)
        torch.manual_seed(seed)
        enc = rnets.cnns.BigPixelEncoder((C, HW, HW), out_dim=O)
        with torch.no_grad():   # the delta-orthogonal init leaves 8 of 9 taps and every bias at zero: fill them
            for p_ in enc.parameters():
                p_.add_(0.05 * torch.randn_like(p_))
        obs = rng.integers(0, 256, (B, C, HW, HW)).astype(np.float32)
        dout = rng.standard_normal((B, O)).astype(np.float32)
        y = enc(torch.as_tensor(obs))
        params = dict(enc.named_parameters())
        grads = torch.autograd.grad(y, list(params.values()), grad_outputs=torch.as_tensor(dout))
        put(out, tag, dict(obs=obs.astype(np.uint8), dout=dout, out=y.detach().numpy()))
        put(out, f"{tag}/params", {k: v.detach().numpy() for k, v in params.items()})
        put(out, f"{tag}/grads", {k: g.numpy() for k, g in zip(params.keys(), grads)})
    path = os.path.join(HERE, "encoder.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}  ({os.path.getsize(path)/1024:.1f} KiB)")

def _last_of(m):
    return m.out if hasattr(m, "out") else (m.act_p if hasattr(m, "act_p") else m.fc3)


def _stack(modules, grad=False):
    get = (lambda p: p.grad.detach().numpy().copy()) if grad else (lambda p: p.detach().numpy().copy())
    return {
        "W1": np.stack([get(m.fc1.weight) for m in modules]), "b1": np.stack([get(m.fc1.bias) for m in modules]),
        "W2": np.stack([get(m.fc2.weight) for m in modules]), "b2": np.stack([get(m.fc2.bias) for m in modules]),
        "W3": np.stack([get(_last_of(m).weight) for m in modules]), "b3": np.stack([get(_last_of(m).bias) for m in modules]),
    }


def run_discrete_case(name, cfg, out_dir=None):
    """SAC-Discrete: critic_update(discrete=True) + Polyak, online_actor_update(discrete=True) and
    alpha_update(discrete=True) of the unmodified reference, driven as main.py:380-414 / :491-543 drive them.  The only
    random draws on this path are the replay indices (replay.py:122) and the target-critic subset (agent.py:29)."""
    rng = np.random.default_rng(cfg.get("seed", 0))
    torch.manual_seed(cfg.get("seed", 0))
    E, N, M = cfg["E"], cfg["N"], cfg["M"]
    S, A, H, B = cfg["S"], cfg["A"], cfg["H"], cfg["B"]
    popart_on = cfg.get("popart", False)
    enc = SharedEncoder(S) if cfg.get("encoder") == "shared" else IdentityEncoder(S)
    agent = ssac.Agent(act_space_size=A, encoder=enc, actor_network_cls=rnets.mlps.DiscreteActor,
                       critic_network_cls=rnets.mlps.DiscreteCritic, discrete=True, ensemble_size=E, num_critics=N,
                       hidden_size=H, auto_rescale_targets=popart_on)
    for m in critic_nets(agent) + list(agent.actors):
        for p in m.parameters():
            if p.dim() == 1:
                p.data.add_(0.05 * torch.randn_like(p))
            elif p.shape[0] == A:
                p.data.mul_(4.0)   # spread the policies / Q rows out (the orthogonal init leaves them near-uniform)
    if popart_on and cfg.get("popart_warm", False):
        for p in agent.popart:
            p._t = 1500
            p.mu = torch.tensor([0.3])
            p.nu = torch.tensor([1.7])
            p.w = torch.tensor([0.9])
            p.b = torch.tensor([0.1])
    target = copy.deepcopy(agent)
    for m in critic_nets(target):
        for p in m.parameters():
            p.data.add_(0.01 * torch.randn_like(p))

    nbuf = cfg.get("nbuf", 64)
    s = rng.standard_normal((nbuf, S)).astype(np.float32)
    a = rng.integers(0, A, size=(nbuf, 1)).astype(np.float32)
    r = rng.standard_normal((nbuf,)).astype(np.float32)
    s1 = rng.standard_normal((nbuf, S)).astype(np.float32)
    d = (rng.uniform(size=(nbuf,)) < 0.1).astype(np.float32)
    buffer = rreplay.ReplayBuffer(size=nbuf + 8)
    buffer.load_experience({"obs": s}, a, r, {"obs": s1}, d)

    out = {"cfg": np.array(repr(cfg))}
    put(out, "buffer", dict(s=s, a=a, r=r, s1=s1, d=d))
    put(out, "init/actors", _stack(agent.actors))
    put(out, "init/critics", _stack(critic_nets(agent)))
    put(out, "init/target_critics", _stack(critic_nets(target)))
    put(out, "init/popart", popart_state(agent))
    if cfg.get("encoder") == "shared":
        put(out, "init/encoder", {k: v.numpy().copy() for k, v in enc.state_dict().items()})
        put(out, "init/target_encoder", {k: v.numpy().copy() for k, v in target.encoder.state_dict().items()})

    critic_opt = torch.optim.Adam(chain(*(c.parameters() for c in agent.critics)), lr=cfg.get("critic_lr", 3e-4),
                                  weight_decay=cfg.get("critic_l2", 0.0), betas=(0.9, 0.999))
    actor_opt = torch.optim.Adam(chain(*(ac.parameters() for ac in agent.actors)), lr=cfg.get("actor_lr", 3e-4),
                                 betas=(0.9, 0.999))
    enc_opt = torch.optim.Adam(agent.encoder.parameters(), lr=1e-4, betas=(0.9, 0.999))
    init_alpha = cfg.get("init_alpha", 0.1)
    log_alphas, alpha_opts = [], []
    for _ in range(E):
        la = torch.Tensor([math.log(init_alpha)])
        la.requires_grad = True
        log_alphas.append(la)
        alpha_opts.append(torch.optim.Adam([la], lr=cfg.get("alpha_lr", 1e-4), betas=(0.5, 0.999)))
    augmenter = raug.AugmentationSequence([raug.IdentityAug(B)])
    target_entropy = -math.log(1.0 / A) * 0.98   # main.py:240-241

    rec = {}
    o_td, o_bw = lu.compute_td_targets, lu.compute_backup_weights

    def td_rec(*a_, **k_):
        res = o_td(*a_, **k_)
        rec.setdefault("td", []).append(res[0].detach().numpy().copy())
        return res

    def bw_rec(*a_, **k_):
        res = o_bw(*a_, **k_)
        rec.setdefault("w", []).append(res.detach().numpy().copy() if torch.is_tensor(res) else np.array(res, dtype=np.float32))
        return res

    lu.compute_td_targets, lu.compute_backup_weights = td_rec, bw_rec
    try:
        replay_dicts = None
        for t in range(cfg["steps"]):
            idx = rng.integers(0, nbuf, size=(E, B))
            subsets = [list(rng.permutation(N)[:M]) for _ in range(E)]
            put(out, f"step{t}/rand", dict(idx=idx, subsets=np.array(subsets)))
            rec.clear()
            with rh.injected(randint=list(idx), subsets=subsets) as q:
                logs, replay_dicts = learning.critic_update(
                    buffer=buffer, agent=agent, target_agent=target, critic_optimizer=critic_opt,
                    encoder_optimizer=enc_opt, log_alphas=log_alphas, batch_size=B, gamma=cfg.get("gamma", 0.99),
                    critic_clip=cfg.get("critic_clip"), encoder_clip=cfg.get("encoder_clip"), target_critic_ensemble_n=M,
                    weighted_bellman_temp=cfg.get("weight_temp"), weight_type=cfg.get("weight_type"),
                    pop=cfg.get("pop", False), augmenter=augmenter, encoder_lambda=0.0, aug_mix=0.0, discrete=True,
                    random_process=None, noise_clip=None, per=False, update_priorities=False,
                    dr3_coeff=cfg.get("dr3_coeff", 0.0))
                assert not q["randint"].items and not q["subsets"].items
            put(out, f"step{t}/td_target", {str(i): v for i, v in enumerate(rec["td"])})
            put(out, f"step{t}/weights", {str(i): v for i, v in enumerate(rec["w"])})
            put(out, f"step{t}/critic_grads", _stack(critic_nets(agent), grad=True))
            put(out, f"step{t}/logs", {k.replace("/", "|"): float(v) for k, v in logs.items()})
            for ac, tc in zip(agent.critics, target.critics):
                lu.soft_update(tc, ac, cfg.get("tau", 0.005))
            put(out, f"step{t}/critics", _stack(critic_nets(agent)))
            put(out, f"step{t}/target_critics", _stack(critic_nets(target)))
            put(out, f"step{t}/popart", popart_state(agent))
            if cfg.get("encoder") == "shared":
                lu.soft_update(target.encoder, agent.encoder, cfg.get("encoder_tau", 0.01))
                put(out, f"step{t}/encoder", {k: v.numpy().copy() for k, v in enc.state_dict().items()})
                put(out, f"step{t}/target_encoder", {k: v.numpy().copy() for k, v in target.encoder.state_dict().items()})
    finally:
        lu.compute_td_targets, lu.compute_backup_weights = o_td, o_bw

    alogs = learning.online_actor_update(
        buffer=buffer, agent=agent, pop=cfg.get("pop", False), actor_optimizer=actor_opt, log_alphas=log_alphas,
        batch_size=B, clip=cfg.get("actor_clip"), random_process=None, noise_clip=None, augmenter=augmenter, aug_mix=0.0,
        premade_replay_dicts=replay_dicts, per=False, discrete=True, use_baseline=False)
    put(out, "actor/grads", _stack(agent.actors, grad=True))
    put(out, "actor/actors", _stack(agent.actors))
    put(out, "actor/logs", {k.replace("/", "|"): float(v) for k, v in alogs.items()})
    llogs = learning.alpha_update(
        buffer=buffer, agent=agent, optimizers=alpha_opts, batch_size=B, log_alphas=log_alphas, augmenter=augmenter,
        aug_mix=0.0, target_entropy=target_entropy, premade_replay_dicts=replay_dicts, discrete=True)
    put(out, "alpha/log_alphas", {str(i): la.detach().numpy().copy() for i, la in enumerate(log_alphas)})
    put(out, "alpha/logs", {k.replace("/", "|"): float(v) for k, v in llogs.items()})
    out["alpha/target_entropy"] = np.array(target_entropy)
    path = os.path.join(out_dir or HERE, f"update_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}  ({os.path.getsize(path)/1024:.1f} KiB)")


def run_discrete_afbc_case():
    """Offline (AFBC) actor update of a DISCRETE agent (learning.py:144-219 with discrete=True, learning_utils.py:241-269)
    with the indirect advantage filter (adv_estimator.py:45-56), and the priority refresh (learning_utils.py:288-295)."""
    cfg = dict(E=2, N=2, S=4, A=3, H=32, B=16, seed=21, actor_clip=40.0)
    rng = np.random.default_rng(cfg["seed"])
    torch.manual_seed(cfg["seed"])
    E, N, S, A, H, B = cfg["E"], cfg["N"], cfg["S"], cfg["A"], cfg["H"], cfg["B"]
    agent = ssac.Agent(act_space_size=A, encoder=IdentityEncoder(S), actor_network_cls=rnets.mlps.DiscreteActor,
                       critic_network_cls=rnets.mlps.DiscreteCritic, discrete=True, ensemble_size=E, num_critics=N,
                       hidden_size=H, auto_rescale_targets=True)
    for m in critic_nets(agent) + list(agent.actors):
        for p in m.parameters():
            if p.dim() == 1:
                p.data.add_(0.05 * torch.randn_like(p))
            elif p.shape[0] == A:
                p.data.mul_(4.0)
    for p in agent.popart:
        p._t = 1500
        p.mu, p.nu, p.w, p.b = torch.tensor([0.3]), torch.tensor([1.7]), torch.tensor([0.9]), torch.tensor([0.1])
    nbuf = 64
    s = rng.standard_normal((nbuf, S)).astype(np.float32)
    a = rng.integers(0, A, size=(nbuf, 1)).astype(np.float32)
    r = rng.standard_normal((nbuf,)).astype(np.float32)
    s1 = rng.standard_normal((nbuf, S)).astype(np.float32)
    d = (rng.uniform(size=(nbuf,)) < 0.1).astype(np.float32)
    buffer = rreplay.ReplayBuffer(size=nbuf + 8)
    buffer.load_experience({"obs": s}, a, r, {"obs": s1}, d)
    out = {"cfg": np.array(repr(cfg))}
    put(out, "buffer", dict(s=s, a=a, r=r, s1=s1, d=d))
    put(out, "init/actors", _stack(agent.actors))
    put(out, "init/critics", _stack(critic_nets(agent)))
    put(out, "init/popart", popart_state(agent))
    actor_opt = torch.optim.Adam(chain(*(ac.parameters() for ac in agent.actors)), lr=3e-4, betas=(0.9, 0.999))
    enc_opt = torch.optim.Adam(agent.encoder.parameters(), lr=1e-4)
    augmenter = raug.AugmentationSequence([raug.IdentityAug(B)])
    idx = rng.integers(0, nbuf, size=(E, B))
    out["rand/idx"] = idx
    for i in range(E):   # the advantage of every member on its own batch (adv_estimator.py:45-56)
        o = {"obs": torch.as_tensor(s[idx[i]])}
        with torch.no_grad():
            out[f"adv/{i}"] = agent.adv_estimator(o, torch.as_tensor(a[idx[i]]), i).numpy().copy()
    rds = []
    o_sma = lu.sample_move_and_augment

    def sma_rec(*a_, **k_):
        rd = o_sma(*a_, **k_)
        rds.append(rd)
        return rd

    lu.sample_move_and_augment = sma_rec
    try:
        with rh.injected(randint=list(idx)):
            logs = learning.offline_actor_update(
                buffer=buffer, agent=agent, actor_optimizer=actor_opt, encoder_optimizer=enc_opt, batch_size=B,
                actor_clip=cfg["actor_clip"], update_encoder=False, encoder_clip=None, augmenter=augmenter, actor_lambda=0.0,
                aug_mix=0.0, premade_replay_dicts=None, per=False, discrete=True, filter_=True)
    finally:
        lu.sample_move_and_augment = o_sma
    put(out, "actor/grads", _stack(agent.actors, grad=True))
    put(out, "actor/actors", _stack(agent.actors))
    put(out, "actor/logs", {k.replace("/", "|"): float(v) for k, v in logs.items()})
    # priority refresh on the last member's batch with ensemble member 1 (random.choice is seeded on both sides)
    got = {}
    buffer.update_priorities = lambda idxs, prios: got.update(idxs=np.asarray(idxs).copy(), prios=np.asarray(prios).copy())
    import random as _r
    _r.seed(7)
    out["priorities/member"] = np.array(_r.choice(range(E)))
    _r.seed(7)
    rds[-1]["priority_idxs"] = idx[-1]
    lu.adjust_priorities({}, rds[-1], agent, buffer)
    out["priorities/idxs"], out["priorities/values"] = got["idxs"], got["prios"].astype(np.float64)
    path = os.path.join(HERE, "discrete_afbc.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}  ({os.path.getsize(path)/1024:.1f} KiB)")


DISCRETE_CASES = {
    # a trainable encoder in front (what a pixel SAC-Discrete agent has): critic gradients flow into it, the actor's do not
    "discrete_encoder": dict(E=2, N=2, M=2, S=6, A=4, H=32, B=16, steps=2, encoder="shared", critic_clip=10.0,
                             encoder_clip=5.0, dr3_coeff=0.01, seed=13),
    # SAC-Discrete with a REDQ-style target subset (2 of 3 critics), DR3 on
    "discrete_sac": dict(E=1, N=3, M=2, S=6, A=5, H=32, B=16, steps=2, dr3_coeff=0.01, critic_clip=40.0, seed=11),
    # ensemble of 2 members: sunrise Bellman weights on gathered Q(s, a), warm PopArt, actor clipping
    "discrete_sunrise_popart": dict(E=2, N=2, M=2, S=4, A=3, H=32, B=16, steps=2, weight_type="sunrise",
                                    weight_temp=20.0, popart=True, pop=True, popart_warm=True, actor_clip=40.0, seed=12),
}


if __name__ == "__main__":
    only = set(sys.argv[1:])   # e.g. ``python make_golden.py rad`` regenerates one fixture
    for name, cfg in DISCRETE_CASES.items():
        if not only or name in only:
            run_discrete_case(name, cfg)
    if not only or "discrete_afbc" in only:
        run_discrete_afbc_case()
    for name, cfg in UPDATE_CASES.items():
        if not only or name in only:
            run_update_case(name, cfg)
    for name, fn in (("replay", run_replay_case), ("aug", run_aug_case), ("rad", run_rad_case), ("afbc", run_afbc_case),
                     ("encoder", run_encoder_case)):
        if not only or name in only:
            fn()
