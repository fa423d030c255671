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


def resync(agent, target, o_agent, o_target):
    """Copy the oracle's parameters over the GPU twins' (after a step has been compared).  Per-step parity is judged
    from identical starting points: the handful of entries that Adam moved by a different +-lr on the two sides (noise
    gradients) would otherwise perturb the next step's gradients at the 1e-3 level.  The free-running case is the
    100-step drift test."""
    for ag, oa in ((agent, o_agent), (target, o_target)):
        cu.load_stack(ag._actor_arena, oa.actors.named())
        cu.load_stack(ag._critic_arena, oa.critics.named())
        if oa.encoder is not None:
            ag.encoder.load_state_dict({k: v.to(ag._critic_arena.device) for k, v in oa.encoder.state_dict().items()})


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


def rows_ambiguous(stack, g, x, tau=2e-6):
    """Rows of ``x`` for which some hidden pre-activation of net ``g`` lies within fp32 rounding of zero:
    |z| < tau * (sum of the magnitudes of its summands)  (fp32 / 3xTF32 dot products land within ~2e-7 of that scale).

    Two correct fp32 forwards (different summation order) may put such a pre-activation on either side of the ReLU.  The
    forward value moves by ~1e-6, but in the backward a whole term appears or vanishes, and because gradients are
    random-sign sums that one term is worth ~1/sqrt(B*H) of EVERY first-layer gradient entry -- far above rtol 1e-4.
    A gradient comparison at BASELINE sizes (millions of pre-activations per update) is therefore only well posed on
    batches without such rows; the parity tests draw twice the rows they need and keep the unambiguous ones."""
    with torch.no_grad():
        W1, b1, W2, b2 = stack.W1[g], stack.b1[g], stack.W2[g], stack.b2[g]
        z1 = x @ W1.t() + b1
        s1 = x.abs() @ W1.abs().t() + b1.abs()
        h1 = torch.relu(z1)
        z2 = h1 @ W2.t() + b2
        s2 = h1 @ W2.abs().t() + b2.abs()
        return (z1.abs() < tau * s1).any(1) | (z2.abs() < tau * s2).any(1)


def keep_rows(bad, B):
    """Positions of the first B candidate rows that are not flagged."""
    good = np.flatnonzero(~np.asarray(bad))
    assert len(good) >= B, f"only {len(good)} unambiguous rows among {len(bad)} candidates"
    return good[:B]


def cmp_stacks(arena, ostack, what, rtol=1e-4, atol_rel=1e-5, atol=0.0, grad=False, noise_lr=None):
    """Every array of an MLP arena against the oracle stack.  atol = ``atol`` + ``atol_rel`` * max|want| per array.
    ``noise_lr`` (post-Adam parameters only): Adam turns a gradient entry that is fp32 rounding noise around zero into a
    step of up to +-lr, so at most 2e-4 of the entries of an array may miss the tolerance, and then by <= 2.1*noise_lr."""
    src = arena.g if grad else arena.p
    for n in uo.PARAM_NAMES:
        want = getattr(ostack, n).numpy().astype(np.float64)
        got = src[n].detach().cpu().numpy().astype(np.float64)
        base = atol + atol_rel * float(np.abs(want).max())
        if noise_lr is not None:
            err = np.abs(got - want)
            bad = err > base + rtol * np.abs(want)
            if bad.any():
                assert bad.mean() <= 2e-4 and float(err[bad].max()) <= 2.1 * noise_lr + base + rtol * float(np.abs(want).max()), \
                    f"{what}.{n}: {int(bad.sum())} of {bad.size} entries off, worst {float(err[bad].max()):.3e}"
            continue
        gu.assert_close(got, want, rtol, base, f"{what}.{n}")


def frac_within(arena, ostack, rtol, atol):
    """(fraction of parameter entries within rtol/atol of the oracle, largest absolute error) over a whole arena."""
    ok = tot = 0
    worst = 0.0
    for n in uo.PARAM_NAMES:
        want = getattr(ostack, n).numpy().astype(np.float64)
        err = np.abs(arena.p[n].detach().cpu().numpy().astype(np.float64) - want)
        ok += int((err <= atol + rtol * np.abs(want)).sum())
        tot += err.size
        worst = max(worst, float(err.max()))
    return ok / tot, worst


def max_err(arena, ostack, grad=False):
    """(max abs error, max |want|) over the arrays of an arena -- for drift reports."""
    src = arena.g if grad else arena.p
    e = m = 0.0
    for n in uo.PARAM_NAMES:
        want = getattr(ostack, n).numpy()
        e = max(e, float(np.abs(src[n].detach().cpu().numpy() - want).max()))
        m = max(m, float(np.abs(want).max()))
    return e, m
