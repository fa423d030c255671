"""GPU agent + CPU-oracle twin with identical parameters at arbitrary (BASELINE-sized) shapes, and the scripted
draws that keep them in lock-step.  Used by tests/test_baseline_shape_parity.py, __graft_entry__.smoke() and
bench.py's sharded-parity check.  The oracle is the checker only (oracle/__init__.py)."""
import copy
import math

import numpy as np
import torch

import cuda_util as cu
import golden_util as gu
from oracle import update_oracle as uo


def t32(x):
    return torch.as_tensor(np.asarray(x, dtype=np.float32))


def make_twins(E, N, S, A, H, det=False, popart=False, seed=0, encoder=None, device="cuda", target_jitter=0.01):
    """Returns (agent, target, o_agent, o_target): super_sac_b200 Agents on ``device`` and OracleAgents on the CPU
    holding the same numbers.  ``encoder``: an nn.Module plugin (deep-copied for every side) or None (identity)."""
    import super_sac_b200 as ssb
    from super_sac_b200 import nets

    gen = torch.Generator().manual_seed(seed)
    o_agent = uo.OracleAgent(E, N, S, A, H, deterministic=det, log_std_low=-5.0, log_std_high=2.0, popart=popart,
                             encoder=copy.deepcopy(encoder) if encoder is not None else None)
    o_agent.actors.random_init(gen)
    o_agent.critics.random_init(gen)
    o_target = o_agent.clone()
    for n in uo.PARAM_NAMES:   # a target that differs from the online nets, as it does after the first Polyak steps
        t = getattr(o_target.critics, n)
        t.add_(target_jitter * torch.randn(t.shape, generator=gen) * t.abs().mean())
    enc = copy.deepcopy(encoder) if encoder is not None else cu.IdentityEncoder(S)
    agent = ssb.Agent(act_space_size=A, encoder=enc,
                      actor_network_cls=nets.mlps.ContinuousDeterministicActor if det else nets.mlps.ContinuousStochasticActor,
                      critic_network_cls=nets.mlps.ContinuousCritic, ensemble_size=E, num_critics=N, hidden_size=H,
                      auto_rescale_targets=popart, log_std_low=-5.0, log_std_high=2.0)
    agent.to(device)
    cu.load_stack(agent._actor_arena, o_agent.actors.named())
    cu.load_stack(agent._critic_arena, o_agent.critics.named())
    target = copy.deepcopy(agent)
    target.to(device)
    cu.load_stack(target._critic_arena, o_target.critics.named())
    return agent, target, o_agent, o_target


def synthetic_state_buffer(n, S, A, seed=0):
    rng = np.random.default_rng(seed)
    return dict(s=rng.standard_normal((n, S), dtype=np.float32), a=rng.uniform(-1, 1, (n, A)).astype(np.float32),
                r=rng.standard_normal(n, dtype=np.float32), s1=rng.standard_normal((n, S), dtype=np.float32),
                d=(rng.uniform(size=n) < 0.05))


def state_batch(buf, idx):
    """primary_batch of learning_utils.py:208-214 for a host buffer dict and indices."""
    return ({"obs": t32(buf["s"][idx])}, t32(buf["a"][idx]), t32(buf["r"][idx]).reshape(-1, 1), {"obs": t32(buf["s1"][idx])},
            t32(buf["d"][idx].astype(np.float32)).reshape(-1, 1))


def oracle_optimizers(o_agent, lr_c=3e-4, lr_a=3e-4, lr_alpha=1e-4, init_alpha=0.1):
    log_alphas = [torch.tensor([math.log(max(init_alpha, 1e-15))], dtype=torch.float32) for _ in range(o_agent.E)]
    return (uo.Adam(o_agent.critics.tensors(), lr=lr_c), uo.Adam(o_agent.actors.tensors(), lr=lr_a), log_alphas,
            [uo.Adam([la], lr=lr_alpha, betas=(0.5, 0.999)) for la in log_alphas])


def cmp_stacks(arena, ostack, what, rtol=1e-4, atol_rel=1e-5, atol=0.0, grad=False, flip_lr=None):
    """Every array of an MLP arena against the oracle stack.  atol = ``atol`` + ``atol_rel`` * max|want| per array.
    ``flip_lr`` (post-Adam parameters only): Adam turns a gradient entry that is fp32 rounding noise around zero into a
    step of +-lr, so at most 1e-4 of the entries of an array may miss the tolerance, and then by no more than 2.1*lr."""
    src = arena.g if grad else arena.p
    for n in uo.PARAM_NAMES:
        want = getattr(ostack, n).numpy().astype(np.float64)
        got = src[n].detach().cpu().numpy().astype(np.float64)
        tol = atol + atol_rel * float(np.abs(want).max()) + rtol * np.abs(want)
        err = np.abs(got - want)
        bad = err > tol
        if flip_lr is not None and bad.any():
            assert bad.mean() <= 1e-4 and float(err[bad].max()) <= 2.1 * flip_lr + float(tol.max()), \
                f"{what}.{n}: {int(bad.sum())} of {bad.size} entries off, worst {float(err[bad].max()):.3e}"
            continue
        gu.assert_close(got, want, rtol, atol + atol_rel * float(np.abs(want).max()), f"{what}.{n}")


def max_err(arena, ostack, grad=False):
    """(max abs error, max |want|) over the arrays of an arena -- for drift reports."""
    src = arena.g if grad else arena.p
    e = m = 0.0
    for n in uo.PARAM_NAMES:
        want = getattr(ostack, n).numpy()
        e = max(e, float(np.abs(src[n].detach().cpu().numpy() - want).max()))
        m = max(m, float(np.abs(want).max()))
    return e, m
