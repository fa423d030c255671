"""Build super_sac_b200 objects from golden fixtures / oracle state (GPU tests, smoke, bench)."""
import copy
import math

import numpy as np
import torch

import golden_util as gu
import super_sac_b200 as ssb
from super_sac_b200 import nets

DEV = "cuda"


class IdentityEncoder(nets.Encoder):
    """experiments/gym/train_gym.py:18-28 of the reference."""

    def __init__(self, dim):
        super().__init__()
        self._dim = dim

    @property
    def embedding_dim(self):
        return self._dim

    def forward(self, obs_dict):
        return obs_dict["obs"]


def load_stack(arena, arrs):
    """arrs: dict W1..b3 (numpy / tensors, stacked [G,...]) -> arena parameters."""
    with torch.no_grad():
        for n in ("W1", "b1", "W2", "b2", "W3", "b3"):
            arena.p[n].copy_(torch.as_tensor(np.asarray(arrs[n])).to(arena.device))


def stack_of(arena):
    return {n: arena.p[n].detach().cpu().numpy().copy() for n in ("W1", "b1", "W2", "b2", "W3", "b3")}


def grads_of(arena):
    return {n: arena.g[n].detach().cpu().numpy().copy() for n in ("W1", "b1", "W2", "b2", "W3", "b3")}


def agent_from_fixture(fx, device=DEV):
    cfg = gu.cfg_of(fx)
    E, N, S, A, H = cfg["E"], cfg["N"], cfg["S"], cfg["A"], cfg["H"]
    det = cfg.get("deterministic", False)
    if cfg.get("encoder") == "shared":
        enc = gu.encoder_from(fx, "init/encoder", S)
    else:
        enc = IdentityEncoder(S)
    agent = ssb.Agent(
        act_space_size=A, encoder=enc,
        actor_network_cls=nets.mlps.ContinuousDeterministicActor if det else nets.mlps.ContinuousStochasticActor,
        critic_network_cls=nets.mlps.ContinuousCritic, ensemble_size=E, num_critics=N, hidden_size=H,
        auto_rescale_targets=cfg.get("popart", False), log_std_low=-5.0, log_std_high=2.0)
    agent.to(device)
    load_stack(agent._actor_arena, gu.sub(fx, "init/actors"))
    load_stack(agent._critic_arena, gu.sub(fx, "init/critics"))
    pst = gu.sub(fx, "init/popart")
    for i, p in enumerate(agent.popart):
        if p:
            p.mu, p.nu, p.w, p.b = pst[f"{i}/mu"], pst[f"{i}/nu"], pst[f"{i}/w"], pst[f"{i}/b"]
            p._t = int(pst[f"{i}/t"])
    target = copy.deepcopy(agent)
    target.to(device)
    load_stack(target._critic_arena, gu.sub(fx, "init/target_critics"))
    if cfg.get("encoder") == "shared":
        target.encoder.load_state_dict({k: torch.as_tensor(v) for k, v in gu.sub(fx, "init/target_encoder").items()})
    return cfg, agent, target


def buffer_from_fixture(fx, device=DEV):
    b = gu.sub(fx, "buffer")
    buf = ssb.replay.ReplayBuffer(size=len(b["a"]) + 8, device=device)
    buf.load_experience({"obs": b["s"]}, b["a"], b["r"], {"obs": b["s1"]}, b["d"])
    return buf


def optimizers(agent, cfg):
    from itertools import chain

    critic_opt = torch.optim.Adam(chain(*(c.parameters() for c in agent.critics)), lr=cfg.get("critic_lr", 3e-4),
                                  weight_decay=cfg.get("critic_l2", 0.0), betas=(0.9, 0.999))
    actor_opt = torch.optim.Adam(chain(*(a.parameters() for a in agent.actors)), lr=cfg.get("actor_lr", 3e-4),
                                 betas=(0.9, 0.999))
    enc_opt = torch.optim.Adam(agent.encoder.parameters(), lr=cfg.get("encoder_lr", 1e-4), betas=(0.9, 0.999))
    dev = agent._critic_arena.device
    init_alpha = max(cfg.get("init_alpha", 0.1), 1e-15)
    log_alphas, alpha_opts = [], []
    for _ in range(cfg["E"]):
        la = torch.Tensor([math.log(init_alpha)]).to(dev)
        la.requires_grad = True
        log_alphas.append(la)
        alpha_opts.append(torch.optim.Adam([la], lr=cfg.get("alpha_lr", 1e-4), betas=(0.5, 0.999)))
    return critic_opt, actor_opt, enc_opt, log_alphas, alpha_opts


class ActionSpace:
    def __init__(self, dim):
        self.low = -np.ones(dim, dtype=np.float32)
        self.high = np.ones(dim, dtype=np.float32)
        self.shape = (dim,)
