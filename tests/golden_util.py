"""Helpers shared by the parity tests: load a golden fixture (written by tests/golden/make_golden.py from the
unmodified reference) and rebuild the oracle state it describes."""
import ast
import math
import os

import numpy as np
import torch

from oracle import update_oracle as uo

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

UPDATE_CASES = ["sac", "redq", "sunrise_popart", "td3_encoder", "softmax_dr3"]
DISCRETE_CASES = ["discrete_sac", "discrete_sunrise_popart"]   # GPU + CPU
DISCRETE_ENCODER_CASE = "discrete_encoder"                      # trainable encoder in front: oracle + host logic (CPU)


def load(name, directory=None):
    z = np.load(os.path.join(directory or GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def cfg_of(fx):
    return ast.literal_eval(str(fx["cfg"]))


def sub(fx, prefix):
    prefix = prefix.rstrip("/") + "/"
    return {k[len(prefix):]: v for k, v in fx.items() if k.startswith(prefix)}


class SharedEncoder(torch.nn.Module):
    """Same trainable encoder as make_golden.py (a user plugin: plain PyTorch on every side)."""

    def __init__(self, dim, hid=16):
        super().__init__()
        self.have_at_least_one_param = torch.nn.Linear(1, 1)  # nets/__init__.py:24
        self.fc0 = torch.nn.Linear(dim, hid)
        self.fc1 = torch.nn.Linear(hid, dim)
        self._dim = dim

    @property
    def embedding_dim(self):
        return self._dim

    def forward(self, obs_dict):
        x = torch.relu(self.fc0(obs_dict["obs"]))
        return torch.relu(self.fc1(x))


def encoder_from(fx, prefix, S):
    enc = SharedEncoder(S)
    enc.load_state_dict({k: torch.as_tensor(v) for k, v in sub(fx, prefix).items()})
    return enc


def popart_from(fx, prefix, E):
    st = sub(fx, prefix)
    out = []
    for i in range(E):
        if f"{i}/mu" not in st:
            out.append(None)
            continue
        p = uo.PopArt()
        p.mu, p.nu = torch.as_tensor(st[f"{i}/mu"]).clone(), torch.as_tensor(st[f"{i}/nu"]).clone()
        p.w, p.b = torch.as_tensor(st[f"{i}/w"]).clone(), torch.as_tensor(st[f"{i}/b"]).clone()
        p.t, p.stable = int(st[f"{i}/t"]), bool(st[f"{i}/stable"])
        out.append(p)
    return out


def oracle_agents(fx):
    cfg = cfg_of(fx)
    E, N, S, A, H = cfg["E"], cfg["N"], cfg["S"], cfg["A"], cfg["H"]
    det = cfg.get("deterministic", False)
    agent = uo.OracleAgent(E, N, S, A, H, deterministic=det, log_std_low=-5.0, log_std_high=2.0)
    agent.actors = uo.MLPStack.from_arrays(sub(fx, "init/actors"))
    agent.critics = uo.MLPStack.from_arrays(sub(fx, "init/critics"))
    agent.popart = popart_from(fx, "init/popart", E)
    target = agent.clone()
    target.critics = uo.MLPStack.from_arrays(sub(fx, "init/target_critics"))
    if cfg.get("encoder") == "shared":
        agent.encoder = encoder_from(fx, "init/encoder", S)
        target.encoder = encoder_from(fx, "init/target_encoder", S)
    return cfg, agent, target


def discrete_oracle_agents(fx, with_target=True):
    from oracle import discrete_oracle as do

    cfg = cfg_of(fx)
    E, N, S, A, H = cfg["E"], cfg["N"], cfg["S"], cfg["A"], cfg["H"]
    agent = do.DiscreteOracleAgent(E, N, S, A, H)
    agent.actors = uo.MLPStack.from_arrays(sub(fx, "init/actors"))
    agent.critics = uo.MLPStack.from_arrays(sub(fx, "init/critics"))
    agent.popart = popart_from(fx, "init/popart", E)
    if cfg.get("encoder") == "shared":
        agent.encoder = encoder_from(fx, "init/encoder", S)
    target = agent.clone()
    if with_target:
        target.critics = uo.MLPStack.from_arrays(sub(fx, "init/target_critics"))
    if cfg.get("encoder") == "shared":
        target.encoder = encoder_from(fx, "init/target_encoder", S)
    return cfg, agent, target


def batch_from(fx, idx):
    """(o, a, r, o1, d) float tensors as learning_utils.py:184-197 produces them."""
    b = sub(fx, "buffer")
    t = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float32))
    return ({"obs": t(b["s"][idx])}, t(b["a"][idx]), t(b["r"][idx]).reshape(-1, 1), {"obs": t(b["s1"][idx])},
            t(b["d"][idx]).reshape(-1, 1))


def hp_from(cfg):
    return dict(gamma=cfg.get("gamma", 0.99), pop=cfg.get("pop", False), weight_type=cfg.get("weight_type"),
                weight_temp=cfg.get("weight_temp"), critic_clip=cfg.get("critic_clip"),
                encoder_clip=cfg.get("encoder_clip"), actor_clip=cfg.get("actor_clip"),
                dr3_coeff=cfg.get("dr3_coeff", 0.0), noise_sigma=cfg.get("noise_sigma"),
                noise_clip=cfg.get("noise_clip"))


def rands_from(fx, prefix, E):
    r = sub(fx, prefix)
    t = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float32))
    out = []
    for i in range(E):
        d = dict(eps=t(r["eps"][i]))
        if "noise" in r:
            d["noise"] = t(r["noise"][i])
        if "subsets" in r:
            d["subset"] = [int(x) for x in r["subsets"][i]]
        if "weight_eps" in r:
            d["weight_eps"] = [t(x) for x in r["weight_eps"][i]]
        out.append(d)
    return out


def log_alphas_from(cfg):
    init_alpha = max(cfg.get("init_alpha", 0.1), 1e-15)
    return [torch.tensor([math.log(init_alpha)], dtype=torch.float32) for _ in range(cfg["E"])]


def assert_close(a, b, rtol, atol, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    if not np.all(err <= tol):
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f"{what}: max violation at {i}: got {a[i]!r} want {b[i]!r} (|err|={err[i]:.3e}, tol={tol[i]:.3e}); "
                             f"max abs err {err.max():.3e}")
