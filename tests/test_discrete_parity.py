"""GPU parity of the SAC-Discrete path (run with -m gpu on a B200).

1. The ssac_discrete_* kernels through the C ABI against the oracle's arithmetic (oracle/discrete_oracle.py, torch-CPU
   fp32) at Atari-like sizes (A = 18 and an A > 32 row that makes the lanes stride), tolerance 1e-5 relative.
2. The drop-in learning.{critic_update, online_actor_update, alpha_update}(discrete=True) + soft_update on the golden
   vectors generated from the UNMODIFIED reference (tests/golden/update_discrete_*.npz), both GEMM implementations.
   Tolerance: north_star's fp32 rtol 1e-4, post-Adam parameters with an atol tied to the learning rate.
3. The same update sequence at a BASELINE-like network size (H = 256, B = 256, N = 2) against the oracle.
"""
import copy
import math
from itertools import chain

import numpy as np
import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _L():
    from super_sac_b200 import _lib

    return _lib.lib(), _lib.stream_ptr()


def _cmp_stack(got, want, what, rtol=RTOL, atol=1e-6):
    for n in ("W1", "b1", "W2", "b2", "W3", "b3"):
        gu.assert_close(got[n], np.asarray(want[n]), rtol, atol, f"{what}.{n}")


def _cmp_logs(logs, want, what):
    for k, v in want.items():
        k2 = k.replace("|", "/")
        if k2.startswith("gradients/"):
            assert k2 in logs
            continue
        assert k2 in logs, f"{what}: missing log key {k2}"
        gu.assert_close(float(logs[k2]), float(v), 2e-4, 2e-5, f"{what} log {k2}")


@pytest.mark.parametrize("B,A,N", [(256, 18, 2), (37, 5, 3), (64, 40, 2), (1024, 6, 10)])
def test_discrete_kernels_match_oracle(B, A, N):
    from oracle import discrete_oracle as do

    L, stream = _L()
    g = torch.Generator().manual_seed(B * 131 + A)
    logits = torch.randn(B, A, generator=g) * 3.0
    q = torch.randn(N, B, A, generator=g)
    act = torch.randint(0, A, (B,), generator=g).float()
    y = torch.randn(B, generator=g)
    w = torch.rand(B, generator=g) + 0.5
    log_alpha = torch.tensor([math.log(0.2)])
    popart = torch.tensor([0.3, 1.7, 0.9, 0.1])
    dev = "cuda"
    d = lambda t: t.to(dev).contiguous()
    lg, qd, ad, yd, wd, lad, pd = d(logits), d(q), d(act), d(y), d(w), d(log_alpha), d(popart)
    probs, logp = do.policy(logits)
    alpha = log_alpha.exp()

    # state value + mean entropy bonus
    v = torch.empty(B, device=dev)
    ent = torch.zeros(1, device=dev)
    L.discrete_value(lg.data_ptr(), qd.data_ptr(), N, B, A, lad.data_ptr(), v.data_ptr(), ent.data_ptr(), stream)
    want_v = (probs * (q.min(0).values - alpha * logp)).sum(-1)
    gu.assert_close(v.cpu().numpy(), want_v.numpy(), 1e-5, 1e-5, "discrete_value v")
    gu.assert_close(ent.cpu().numpy(), (alpha * logp).mean().reshape(1).numpy(), 1e-4, 1e-6, "discrete_value entropy log")

    # gather
    out = torch.empty(N, B, device=dev)
    L.discrete_gather_q(qd.data_ptr(), ad.data_ptr(), N, B, A, out.data_ptr(), stream)
    want_sel = q.gather(-1, act.long().reshape(1, B, 1).expand(N, B, 1)).squeeze(-1)
    assert np.array_equal(out.cpu().numpy(), want_sel.numpy())

    # critic loss seed (with PopArt and weights), E = 2
    for pop in (0, 1):
        dy = torch.full((N, B, A), 7.0, device=dev)
        loss = torch.zeros(2, device=dev)
        L.discrete_critic_loss_seed(qd.data_ptr(), N, B, A, ad.data_ptr(), yd.data_ptr(), wd.data_ptr(), None,
                                    pd.data_ptr(), pop, 2, 0, dy.data_ptr(), loss.data_ptr(), stream)
        pw, pb = (popart[2], popart[3]) if pop else (torch.tensor(1.0), torch.tensor(0.0))
        td = y[None, :] - (pw * want_sel + pb)
        want_dy = torch.zeros(N, B, A)
        want_dy.scatter_(-1, act.long().reshape(1, B, 1).expand(N, B, 1), (-2.0 * w * td * pw / (B * 2 * N)).unsqueeze(-1))
        gu.assert_close(dy.cpu().numpy(), want_dy.numpy(), 1e-5, 1e-9, f"critic seed dy (pop={pop})")
        gu.assert_close(loss.cpu().numpy(), np.array([float((w * td * td).sum() / (B * 2 * N)), float(td[N - 1].mean())]),
                        1e-4, 1e-6, f"critic seed loss (pop={pop})")

    # actor seed
    for pop in (0, 1):
        dl = torch.empty(B, A, device=dev)
        loss = torch.zeros(1, device=dev)
        L.discrete_actor_seed(lg.data_ptr(), qd.data_ptr(), N, B, A, lad.data_ptr(), pd.data_ptr(), pop, 2, dl.data_ptr(),
                              loss.data_ptr(), stream)
        vals = q.min(0).values
        if pop:
            vals = popart[2] * vals + popart[3]
        gq = vals - alpha * logp
        f = (probs * gq).sum(-1, keepdim=True)
        gu.assert_close(dl.cpu().numpy(), ((-1.0 / (2 * B)) * probs * (gq - f)).numpy(), 1e-4, 1e-8, f"actor seed (pop={pop})")
        gu.assert_close(loss.cpu().numpy(), (-f.mean() / 2).reshape(1).numpy(), 1e-4, 1e-6, f"actor loss (pop={pop})")

    # negative entropy
    ne = torch.empty(B, device=dev)
    L.discrete_neg_entropy(lg.data_ptr(), B, A, ne.data_ptr(), stream)
    gu.assert_close(ne.cpu().numpy(), (probs * logp).sum(-1).numpy(), 1e-5, 1e-6, "neg entropy")


def _discrete_agent(cfg, stacks, device="cuda"):
    import cuda_util as cu
    import super_sac_b200 as ssb
    from super_sac_b200 import nets

    E, N, S, A, H = cfg["E"], cfg["N"], cfg["S"], cfg["A"], cfg["H"]
    agent = ssb.Agent(act_space_size=A, encoder=cu.IdentityEncoder(S), actor_network_cls=nets.mlps.DiscreteActor,
                      critic_network_cls=nets.mlps.DiscreteCritic, discrete=True, ensemble_size=E, num_critics=N,
                      hidden_size=H, auto_rescale_targets=cfg.get("popart", False))
    agent.to(device)
    cu.load_stack(agent._actor_arena, stacks["actors"])
    cu.load_stack(agent._critic_arena, stacks["critics"])
    target = copy.deepcopy(agent)
    target.to(device)
    cu.load_stack(target._critic_arena, stacks["target_critics"])
    return agent, target


def _optimizers(agent, cfg):
    critic_opt = torch.optim.Adam(chain(*(c.parameters() for c in agent.critics)), lr=3e-4, betas=(0.9, 0.999))
    actor_opt = torch.optim.Adam(chain(*(a.parameters() for a in agent.actors)), lr=3e-4, betas=(0.9, 0.999))
    enc_opt = torch.optim.Adam(agent.encoder.parameters(), lr=1e-4)
    dev = agent._critic_arena.device
    log_alphas, alpha_opts = [], []
    for _ in range(cfg["E"]):
        la = torch.Tensor([math.log(cfg.get("init_alpha", 0.1))]).to(dev)
        la.requires_grad = True
        log_alphas.append(la)
        alpha_opts.append(torch.optim.Adam([la], lr=1e-4, betas=(0.5, 0.999)))
    return critic_opt, actor_opt, enc_opt, log_alphas, alpha_opts


@pytest.mark.parametrize("impl", ["tcgen05", "ffma"])
@pytest.mark.parametrize("case", gu.DISCRETE_CASES)
def test_discrete_update_matches_reference(case, impl):
    import cuda_util as cu
    import super_sac_b200 as ssb
    from super_sac_b200 import _rng, augmentations, learning, learning_utils as lu

    ssb.set_mlp_impl(impl)
    fx = gu.load("update_" + case)
    cfg = gu.cfg_of(fx)
    E, N, M, B = cfg["E"], cfg["N"], cfg["M"], cfg["B"]
    agent, target = _discrete_agent(cfg, dict(actors=gu.sub(fx, "init/actors"), critics=gu.sub(fx, "init/critics"),
                                              target_critics=gu.sub(fx, "init/target_critics")))
    pst = gu.sub(fx, "init/popart")
    for holder in (agent, target):
        for i, p in enumerate(holder.popart):
            if p:
                p.mu, p.nu, p.w, p.b = pst[f"{i}/mu"], pst[f"{i}/nu"], pst[f"{i}/w"], pst[f"{i}/b"]
                p._t = int(pst[f"{i}/t"])
    buf = cu.buffer_from_fixture(fx)
    critic_opt, actor_opt, enc_opt, log_alphas, alpha_opts = _optimizers(agent, cfg)
    augmenter = augmentations.AugmentationSequence([augmentations.IdentityAug(B)])
    rec = {}
    o_td, o_bw = lu.compute_td_targets, lu.compute_backup_weights

    def td_rec(*a_, **k_):
        res = o_td(*a_, **k_)
        rec.setdefault("td", []).append(res[0])
        return res

    def bw_rec(*a_, **k_):
        res = o_bw(*a_, **k_)
        rec.setdefault("w", []).append(res)
        return res

    lu.compute_td_targets, lu.compute_backup_weights = td_rec, bw_rec
    old_src = _rng.set_source(_rng.ScriptedSource())
    try:
        replay_dicts = None
        for t in range(cfg["steps"]):
            src = _rng.ScriptedSource()
            _rng.set_source(src)
            r = gu.sub(fx, f"step{t}/rand")
            for i in range(E):
                src.push("indices", r["idx"][i])
                src.push("subsets", r["subsets"][i].astype(np.int32))
            rec.clear()
            logs, replay_dicts = learning.critic_update(
                buffer=buf, agent=agent, target_agent=target, critic_optimizer=critic_opt, encoder_optimizer=enc_opt,
                log_alphas=log_alphas, batch_size=B, gamma=cfg.get("gamma", 0.99), critic_clip=cfg.get("critic_clip"),
                encoder_clip=None, target_critic_ensemble_n=M, weighted_bellman_temp=cfg.get("weight_temp"),
                weight_type=cfg.get("weight_type"), pop=cfg.get("pop", False), augmenter=augmenter, encoder_lambda=0.0,
                aug_mix=0.0, discrete=True, random_process=None, noise_clip=None, per=False, update_priorities=False,
                dr3_coeff=cfg.get("dr3_coeff", 0.0))
            assert src.empty(), "not every scripted draw was consumed"
            for i in range(E):
                gu.assert_close(rec["td"][i].cpu().numpy(), fx[f"step{t}/td_target/{i}"], RTOL, 1e-5, f"step{t} td_target[{i}]")
                w = rec["w"][i]
                w = w.cpu().numpy() if torch.is_tensor(w) else np.array(w, dtype=np.float32)
                gu.assert_close(w, fx[f"step{t}/weights/{i}"], RTOL, 1e-5, f"step{t} weights[{i}]")
            _cmp_stack(cu.grads_of(agent._critic_arena), gu.sub(fx, f"step{t}/critic_grads"), f"step{t} critic_grads", atol=2e-7)
            _cmp_logs(logs, gu.sub(fx, f"step{t}/logs"), f"step{t}")
            for ac, tc in zip(agent.critics, target.critics):
                lu.soft_update(tc, ac, cfg.get("tau", 0.005))
            _cmp_stack(cu.stack_of(agent._critic_arena), gu.sub(fx, f"step{t}/critics"), f"step{t} critics", atol=3e-4 * 0.05)
            _cmp_stack(cu.stack_of(target._critic_arena), gu.sub(fx, f"step{t}/target_critics"), f"step{t} target_critics",
                       atol=3e-4 * 0.05)
            want_pop = gu.sub(fx, f"step{t}/popart")
            for i, p in enumerate(agent.popart):
                if p:
                    for n in ("mu", "nu", "w", "b"):
                        gu.assert_close(getattr(p, n).cpu().numpy(), want_pop[f"{i}/{n}"], RTOL, 1e-6, f"step{t} popart[{i}].{n}")
        alogs = learning.online_actor_update(
            buffer=buf, agent=agent, pop=cfg.get("pop", False), actor_optimizer=actor_opt, log_alphas=log_alphas,
            batch_size=B, clip=cfg.get("actor_clip"), random_process=None, noise_clip=None, augmenter=augmenter, aug_mix=0.0,
            premade_replay_dicts=replay_dicts, per=False, discrete=True, use_baseline=False)
        _cmp_stack(cu.grads_of(agent._actor_arena), gu.sub(fx, "actor/grads"), "actor grads", atol=2e-7)
        _cmp_stack(cu.stack_of(agent._actor_arena), gu.sub(fx, "actor/actors"), "actors", atol=3e-4 * 0.05)
        _cmp_logs(alogs, gu.sub(fx, "actor/logs"), "actor")
        llogs = learning.alpha_update(
            buffer=buf, agent=agent, optimizers=alpha_opts, batch_size=B, log_alphas=log_alphas, augmenter=augmenter,
            aug_mix=0.0, target_entropy=float(fx["alpha/target_entropy"]), premade_replay_dicts=replay_dicts, discrete=True)
        for i, la in enumerate(log_alphas):
            gu.assert_close(la.detach().cpu().numpy(), fx[f"alpha/log_alphas/{i}"], 1e-6, 1e-7, f"log_alpha[{i}]")
        _cmp_logs(llogs, gu.sub(fx, "alpha/logs"), "alpha")
    finally:
        lu.compute_td_targets, lu.compute_backup_weights = o_td, o_bw
        _rng.set_source(old_src)
        ssb.set_mlp_impl("tcgen05")


def test_discrete_update_baseline_size_matches_oracle():
    """Atari-like SAC-Discrete step at the BASELINE network size (H = 256, B = 256, 2 critics, 18 actions, 64 features):
    critic update + Polyak, actor update, temperature update against the oracle on the same batch, then the discrete
    acting path (greedy / sampled action indices) on the device."""
    import cuda_util as cu
    from oracle import discrete_oracle as do
    from oracle import update_oracle as uo
    from super_sac_b200 import _rng, augmentations, learning, learning_utils as lu
    import super_sac_b200 as ssb

    torch.set_num_threads(4)
    cfg = dict(E=1, N=2, M=2, S=64, A=18, H=256, B=256)
    E, N, M, S, A, H, B = (cfg[k] for k in "ENMSAHB")
    gen = torch.Generator().manual_seed(5)
    oa = do.DiscreteOracleAgent(E, N, S, A, H)
    oa.actors.random_init(gen)
    oa.critics.random_init(gen)
    ot = oa.clone()
    ot.critics.random_init(gen)
    agent, target = _discrete_agent(cfg, dict(actors=oa.actors.named(), critics=oa.critics.named(),
                                              target_critics=ot.critics.named()))
    rng = np.random.default_rng(0)
    nbuf = 512
    s = rng.standard_normal((nbuf, S)).astype(np.float32)
    a = rng.integers(0, A, size=(nbuf, 1)).astype(np.float32)
    r = rng.standard_normal((nbuf,)).astype(np.float32)
    s1 = rng.standard_normal((nbuf, S)).astype(np.float32)
    d = (rng.uniform(size=(nbuf,)) < 0.1).astype(np.float32)
    buf = ssb.replay.ReplayBuffer(size=nbuf + 8, device="cuda")
    buf.load_experience({"obs": s}, a, r, {"obs": s1}, d)
    critic_opt, actor_opt, enc_opt, log_alphas, alpha_opts = _optimizers(agent, cfg)
    o_la = [torch.tensor([math.log(0.1)])]
    o_copt, o_aopt = uo.Adam(oa.critics.tensors(), lr=3e-4), uo.Adam(oa.actors.tensors(), lr=3e-4)
    o_alopt = [uo.Adam(o_la, lr=1e-4, betas=(0.5, 0.999))]
    augmenter = augmentations.AugmentationSequence([augmentations.IdentityAug(B)])
    hp = dict(gamma=0.99, critic_clip=None, dr3_coeff=0.0)
    old_src = _rng.set_source(_rng.ScriptedSource())
    try:
        replay_dicts = None
        for t in range(3):
            idx = rng.integers(0, nbuf, size=B)
            src = _rng.ScriptedSource()
            _rng.set_source(src)
            src.push("indices", idx)
            src.push("subsets", np.array([1, 0], dtype=np.int32))
            tt = lambda x: torch.as_tensor(x)
            batch = ({"obs": tt(s[idx])}, tt(a[idx]), tt(r[idx]).reshape(-1, 1), {"obs": tt(s1[idx])}, tt(d[idx]).reshape(-1, 1))
            ologs, aux = do.critic_update(oa, ot, [batch], [[1, 0]], hp, o_la, o_copt)
            logs, replay_dicts = learning.critic_update(
                buffer=buf, agent=agent, target_agent=target, critic_optimizer=critic_opt, encoder_optimizer=enc_opt,
                log_alphas=log_alphas, batch_size=B, gamma=0.99, critic_clip=None, encoder_clip=None,
                target_critic_ensemble_n=M, weighted_bellman_temp=None, weight_type=None, pop=False, augmenter=augmenter,
                encoder_lambda=0.0, aug_mix=0.0, discrete=True, random_process=None, noise_clip=None, per=False,
                update_priorities=False, dr3_coeff=0.0)
            assert src.empty()
            _cmp_stack(cu.grads_of(agent._critic_arena), aux["grads"].named(), f"step{t} critic grads", atol=1e-6)
            for k in ("losses/critic_overall_loss", "td_targets/mean_td_target_0", "td_targets/entropy_bonus_0"):
                gu.assert_close(float(logs[k]), float(ologs[k]), 2e-4, 2e-5, f"step{t} log {k}")
            uo.soft_update(ot.critics.tensors(), oa.critics.tensors(), 0.005)
            for ac, tc in zip(agent.critics, target.critics):
                lu.soft_update(tc, ac, 0.005)
            _cmp_stack(cu.stack_of(agent._critic_arena), oa.critics.named(), f"step{t} critics", atol=3e-4 * 0.05)
            _cmp_stack(cu.stack_of(target._critic_arena), ot.critics.named(), f"step{t} target critics", atol=3e-4 * 0.05)
        ologs, aux = do.online_actor_update(oa, [batch], hp, o_la, o_aopt)
        alogs = learning.online_actor_update(
            buffer=buf, agent=agent, pop=False, actor_optimizer=actor_opt, log_alphas=log_alphas, batch_size=B, clip=None,
            random_process=None, noise_clip=None, augmenter=augmenter, aug_mix=0.0, premade_replay_dicts=replay_dicts,
            per=False, discrete=True, use_baseline=False)
        _cmp_stack(cu.grads_of(agent._actor_arena), aux["grads"].named(), "actor grads", atol=1e-7)
        gu.assert_close(float(alogs["losses/actor_pg_loss"]), float(ologs["losses/actor_pg_loss"]), 2e-4, 2e-5, "actor loss")
        te = -math.log(1.0 / A) * 0.98
        ologs = do.alpha_update(oa, [batch], o_la, o_alopt, te)
        llogs = learning.alpha_update(buffer=buf, agent=agent, optimizers=alpha_opts, batch_size=B, log_alphas=log_alphas,
                                      augmenter=augmenter, aug_mix=0.0, target_entropy=te,
                                      premade_replay_dicts=replay_dicts, discrete=True)
        gu.assert_close(log_alphas[0].detach().cpu().numpy(), o_la[0].numpy(), 1e-6, 1e-7, "log_alpha")
        gu.assert_close(float(llogs["losses/alpha_loss_0"]), float(ologs["losses/alpha_loss_0"]), 2e-4, 2e-5, "alpha loss")
        # acting path: action indices, greedy = argmax of the mean policy (agent.py:204-221)
        obs = {"obs": s[:7]}
        greedy = agent.forward(obs, num_envs=7)
        probs, _ = do.policy(uo.mlp_forward(oa.actors, 0, torch.as_tensor(s[:7]))[0])
        assert greedy.shape == (7, 1)
        chosen = probs.numpy()[np.arange(7), greedy[:, 0]]
        assert np.all(chosen >= probs.max(-1).values.numpy() - 1e-5)   # argmax up to fp32 near-ties
        act = agent.sample_action(obs, num_envs=7)
        assert act.shape == (7, 1) and act.min() >= 0 and act.max() < A
    finally:
        _rng.set_source(old_src)


@pytest.mark.parametrize("B,A,N,E", [(256, 18, 2, 1), (37, 5, 3, 3), (64, 40, 2, 2)])
def test_discrete_advantage_and_bc_kernels_match_oracle(B, A, N, E):
    from oracle import discrete_oracle as do

    L, stream = _L()
    g = torch.Generator().manual_seed(B * 17 + A)
    logits = torch.randn(E, B, A, generator=g) * 2.0
    q = torch.randn(N, B, A, generator=g)
    act = torch.randint(0, A, (B,), generator=g).float()
    popart = torch.tensor([0.3, 1.7, 0.9, 0.1])
    dev = "cuda"
    lg, qd, ad, pd = (t.to(dev).contiguous() for t in (logits, q, act, popart))
    idx = act.long().unsqueeze(-1)
    for use_pop in (False, True):
        adv = torch.empty(B, device=dev)
        mask = torch.empty(B, device=dev)
        prio = torch.empty(B, dtype=torch.float64, device=dev)
        L.discrete_advantage(lg.data_ptr(), E, qd.data_ptr(), N, B, A, ad.data_ptr(), pd.data_ptr() if use_pop else None,
                             adv.data_ptr(), mask.data_ptr(), prio.data_ptr(), stream)
        pm = torch.stack([do.policy(logits[e])[0] for e in range(E)], 0).mean(0)
        mq = q.min(0).values
        if use_pop:
            mq = popart[2] * mq + popart[3]
        want = (mq.gather(-1, idx) - (pm * mq).sum(-1, keepdim=True)).squeeze(-1)
        gu.assert_close(adv.cpu().numpy(), want.numpy(), 1e-5, 2e-6, f"advantage (popart={use_pop})")
        clear = want.abs() > 1e-5   # rows whose sign is not a rounding question
        assert np.array_equal(mask.cpu().numpy()[clear.numpy()], (want >= 0).float().numpy()[clear.numpy()])
        gu.assert_close(prio.cpu().numpy(), (torch.relu(want).double() + 1e-4).numpy(), 1e-5, 2e-6, "priority")
    m = (torch.rand(B, generator=g) > 0.5).float()
    md = m.to(dev)
    probs, logp = do.policy(logits[0])
    onehot = torch.zeros(B, A).scatter_(1, idx, 1.0)
    for use_mask in (False, True):
        dl = torch.empty(B, A, device=dev)
        loss = torch.zeros(1, device=dev)
        L.discrete_bc_seed(lg.data_ptr(), ad.data_ptr(), md.data_ptr() if use_mask else None, B, A, 2, dl.data_ptr(),
                           loss.data_ptr(), stream)
        mm = m if use_mask else torch.ones(B)
        gu.assert_close(dl.cpu().numpy(), ((-mm / (B * 2)).unsqueeze(-1) * (onehot - probs)).numpy(), 1e-4, 1e-8, "bc seed")
        gu.assert_close(loss.cpu().numpy(), (-(mm * logp.gather(-1, idx).squeeze(-1)).mean()).reshape(1).numpy(), 1e-4, 1e-6,
                        "bc loss")


@pytest.mark.parametrize("impl", ["tcgen05", "ffma"])
def test_discrete_offline_update_matches_reference(impl):
    """offline_actor_update(discrete=True) with the advantage filter, agent.adv_estimator, compute_filter_stats and
    adjust_priorities (PER trees written) on the golden of the unmodified reference."""
    import random

    import cuda_util as cu
    import super_sac_b200 as ssb
    from super_sac_b200 import _rng, augmentations, learning, learning_utils as lu

    ssb.set_mlp_impl(impl)
    fx = gu.load("discrete_afbc")
    cfg = gu.cfg_of(fx)
    cfg["popart"] = True
    E, B = cfg["E"], cfg["B"]
    agent, _ = _discrete_agent(cfg, dict(actors=gu.sub(fx, "init/actors"), critics=gu.sub(fx, "init/critics"),
                                         target_critics=gu.sub(fx, "init/critics")))
    pst = gu.sub(fx, "init/popart")
    for i, p in enumerate(agent.popart):
        p.mu, p.nu, p.w, p.b = pst[f"{i}/mu"], pst[f"{i}/nu"], pst[f"{i}/w"], pst[f"{i}/b"]
        p._t = int(pst[f"{i}/t"])
    buf = cu.buffer_from_fixture(fx)
    idx = fx["rand/idx"]
    bufd = gu.sub(fx, "buffer")
    dev = agent._critic_arena.device
    old_src = _rng.set_source(_rng.ScriptedSource())
    try:
        for i in range(E):
            o = {"obs": torch.as_tensor(bufd["s"][idx[i]]).to(dev)}
            adv = agent.adv_estimator(o, torch.as_tensor(bufd["a"][idx[i]]).to(dev), i)
            gu.assert_close(adv.cpu().numpy(), fx[f"adv/{i}"], RTOL, 2e-6, f"adv[{i}]")
        _, actor_opt, enc_opt, _, _ = _optimizers(agent, cfg)
        augmenter = augmentations.AugmentationSequence([augmentations.IdentityAug(B)])
        src = _rng.ScriptedSource()
        _rng.set_source(src)
        for i in range(E):
            src.push("indices", idx[i])
        logs = learning.offline_actor_update(
            buffer=buf, agent=agent, actor_optimizer=actor_opt, encoder_optimizer=enc_opt, batch_size=B,
            actor_clip=cfg["actor_clip"], update_encoder=False, encoder_clip=None, augmenter=augmenter, actor_lambda=0.0,
            aug_mix=0.0, premade_replay_dicts=None, per=False, discrete=True, filter_=True)
        assert src.empty()
        _cmp_stack(cu.grads_of(agent._actor_arena), gu.sub(fx, "actor/grads"), "actor grads", atol=2e-7)
        _cmp_stack(cu.stack_of(agent._actor_arena), gu.sub(fx, "actor/actors"), "actors", atol=3e-4 * 0.05)
        _cmp_logs(logs, gu.sub(fx, "actor/logs"), "offline actor")
        # priority refresh on the last member's batch (random.choice seeded like the generator): values, then the trees
        src.push("indices", idx[-1])
        rd = lu.sample_move_and_augment(buffer=buf, batch_size=B, augmenter=augmenter, aug_mix=0.0, per=False)
        got = {}
        o_up = buf.update_priorities

        def rec(idxs, prios):
            got["idxs"], got["prios"] = idxs, prios
            return o_up(idxs, prios)

        buf.update_priorities = rec
        random.seed(7)
        lu.adjust_priorities({}, rd, agent, buf)
        ii = got["idxs"].cpu().numpy() if torch.is_tensor(got["idxs"]) else np.asarray(got["idxs"])
        assert np.array_equal(ii, fx["priorities/idxs"])
        pr = got["prios"].cpu().numpy() if torch.is_tensor(got["prios"]) else np.asarray(got["prios"])
        gu.assert_close(pr, fx["priorities/values"], 1e-5, 1e-6, "priorities")
        src.push("indices", idx[0])
        pct = lu.compute_filter_stats(buf, agent, augmenter, B)
        assert 0.0 <= pct <= 100.0
    finally:
        _rng.set_source(old_src)
        ssb.set_mlp_impl("tcgen05")
