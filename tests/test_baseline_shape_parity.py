"""Update-level parity at the BASELINE.json shapes ("parity gaps" 1-2): the drop-in
critic_update -> soft_update -> online_actor_update -> alpha_update (and the offline AFBC step) against the CPU
oracle (oracle/update_oracle.py, pinned to the unmodified reference by tests/test_oracle_golden.py) on injected
draws, at

  C2  REDQ      N=10, M=2, H=256, B=256, obs 17 / act 6                     (+ a 100-step drift run)
  C3  SUNRISE   E=5, N=2, H=256, B=256, weight_type="sunrise", T=20          (PopArt off / on)
  C5  AFBC      N=2, H=1024, B=1024, DR3 0.01, clips 40, PER + priority refresh, offline_actor_update
  C4  DrQv2     u8 9x84x84 frames, B=512, BigPixelEncoder, H=1024, deterministic actor + TD3 noise, Drqv2Aug

Tolerance (north_star): rtol 1e-4 for Q-values / gradients / post-step parameters, with an absolute floor of 1e-5 of
each array's largest entry for gradients (sums over B*H products whose terms cancel) and lr*0.05 for post-Adam
parameters (Adam turns a gradient that is rounding noise around zero into a +-lr step).
"""
import copy
import random as pyrandom

import numpy as np
import pytest
import torch

import golden_util as gu
import twin_util as tw
from oracle import aug_oracle as ao
from oracle import replay_oracle as ro
from oracle import update_oracle as uo

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _critic_kw(buf, agent, target, c_opt, e_opt, las, B, M, aug, **over):
    kw = dict(buffer=buf, agent=agent, target_agent=target, critic_optimizer=c_opt, encoder_optimizer=e_opt, log_alphas=las,
              batch_size=B, gamma=0.99, critic_clip=None, encoder_clip=None, target_critic_ensemble_n=M,
              weighted_bellman_temp=None, weight_type=None, pop=False, augmenter=aug, encoder_lambda=0.0,
              random_process=None, noise_clip=None, aug_mix=0.0)
    kw.update(over)
    return kw


def _cmp_logs(got, want, what, skip=("gradients/",)):
    for k, v in want.items():
        if k.startswith(skip):
            continue
        assert k in got, f"{what}: missing log key {k}"
        gu.assert_close(float(got[k]), float(v), 2e-4, 2e-5, f"{what} log {k}")


def _state_setup(E, N, S, A, H, B, nbuf, popart=False, seed=0):
    import cuda_util as cu
    import super_sac_b200 as ssb
    from super_sac_b200 import augmentations

    agent, target, o_agent, o_target = tw.make_twins(E, N, S, A, H, popart=popart, seed=seed)
    hb = tw.synthetic_state_buffer(nbuf, S, A, seed)
    buf = ssb.replay.ReplayBuffer(nbuf, device="cuda")
    buf.load_experience({"obs": hb["s"]}, hb["a"], hb["r"], {"obs": hb["s1"]}, hb["d"])
    c_opt, a_opt, e_opt, las, al_opts = cu.optimizers(agent, dict(E=E))
    o_c, o_a, o_las, o_al = tw.oracle_optimizers(o_agent)
    aug = augmentations.AugmentationSequence([augmentations.IdentityAug(B)])
    return agent, target, o_agent, o_target, hb, buf, (c_opt, a_opt, e_opt, las, al_opts), (o_c, o_a, o_las, o_al), aug


def _run_state_steps(E, N, M, S, A, H, B, steps, popart=False, pop=False, weight_type=None, temp=None, seed=0,
                     target_delay=2, check_every=1, final_atol_lr=0.05):
    """critic_update (+Polyak) x steps, then actor + alpha update on the last batch; GPU vs oracle after every
    ``check_every`` steps.  Returns the worst post-step parameter error seen (for the drift report)."""
    from super_sac_b200 import _rng, learning, learning_utils as lu

    nbuf = 5000
    agent, target, o_agent, o_target, hb, buf, opts, o_opts, aug = _state_setup(E, N, S, A, H, B, nbuf, popart, seed)
    c_opt, a_opt, e_opt, las, al_opts = opts
    o_c, o_a, o_las, o_al = o_opts
    hp = dict(gamma=0.99, pop=pop, weight_type=weight_type, weight_temp=temp)
    kw = _critic_kw(buf, agent, target, c_opt, e_opt, las, B, M, aug, pop=pop, weight_type=weight_type,
                    weighted_bellman_temp=temp)
    rng = np.random.default_rng(seed + 100)
    old = _rng.set_source(_rng.ScriptedSource())
    worst = 0.0
    try:
        rds = batches = None
        for t in range(steps):
            src = _rng.ScriptedSource()
            _rng.set_source(src)
            batches, rands = [], []
            for i in range(E):
                idx = rng.integers(0, nbuf, B)
                eps = rng.standard_normal((B, A)).astype(np.float32)
                subset = rng.permutation(N)[:M]
                src.push("indices", idx).push("subsets", subset.astype(np.int32)).push("normal", eps)
                batches.append(tw.state_batch(hb, idx))
                rands.append(dict(eps=tw.t32(eps), subset=[int(x) for x in subset]))
            logs, rds = learning.critic_update(**kw)
            assert src.empty(), "not every scripted draw was consumed"
            ologs, aux = uo.critic_update(o_agent, o_target, batches, rands, hp, o_las, o_c)
            if t % target_delay == 0:
                for ac, tc in zip(agent.critics, target.critics):
                    lu.soft_update(tc, ac, 0.005)
                uo.soft_update(o_target.critics.tensors(), o_agent.critics.tensors(), 0.005)
            if t % check_every == 0 or t == steps - 1:
                if check_every == 1:   # single-step parity: the gradients themselves
                    tw.cmp_stacks(agent._critic_arena, aux["grads"], f"step{t} critic grads", RTOL, 1e-5, grad=True)
                    _cmp_logs(logs, ologs, f"step{t}")
                tw.cmp_stacks(agent._critic_arena, o_agent.critics, f"step{t} critics", RTOL, 0.0, 3e-4 * final_atol_lr, flip_lr=3e-4)
                tw.cmp_stacks(target._critic_arena, o_target.critics, f"step{t} target critics", RTOL, 0.0, 3e-4 * final_atol_lr, flip_lr=3e-4)
                worst = max(worst, tw.max_err(agent._critic_arena, o_agent.critics)[0])
                for i, p in enumerate(agent.popart):
                    if p:
                        op = o_agent.popart[i]
                        for n in ("mu", "nu", "w", "b"):
                            gu.assert_close(getattr(p, n).cpu().numpy(), getattr(op, n).numpy(), RTOL, 1e-6, f"step{t} popart[{i}].{n}")
        # ---- actor + temperature update on the last critic batch ---------------------------------------------
        src = _rng.ScriptedSource()
        _rng.set_source(src)
        arands = []
        for i in range(E):
            eps = rng.standard_normal((B, A)).astype(np.float32)
            src.push("normal", eps)
            arands.append(dict(eps=tw.t32(eps)))
        alogs = learning.online_actor_update(buffer=buf, agent=agent, pop=pop, actor_optimizer=a_opt, log_alphas=las,
                                             batch_size=B, clip=None, random_process=None, noise_clip=None, augmenter=aug,
                                             aug_mix=0.0, premade_replay_dicts=rds)
        assert src.empty()
        oalogs, aaux = uo.online_actor_update(o_agent, batches, arands, hp, o_las, o_a)
        tw.cmp_stacks(agent._actor_arena, aaux["grads"], "actor grads", RTOL, 1e-5, grad=True)
        tw.cmp_stacks(agent._actor_arena, o_agent.actors, "actors", RTOL, 0.0, 3e-4 * 0.05, flip_lr=3e-4)
        _cmp_logs(alogs, oalogs, "actor")
        src = _rng.ScriptedSource()
        _rng.set_source(src)
        lrands = []
        for i in range(E):
            eps = rng.standard_normal((B, A)).astype(np.float32)
            src.push("normal", eps)
            lrands.append(dict(eps=tw.t32(eps)))
        llogs = learning.alpha_update(buffer=buf, agent=agent, optimizers=al_opts, batch_size=B, log_alphas=las, augmenter=aug,
                                      aug_mix=0.0, target_entropy=-float(A), premade_replay_dicts=rds, discrete=False)
        ollogs = uo.alpha_update(o_agent, batches, lrands, o_las, o_al, -float(A))
        for i, la in enumerate(las):
            gu.assert_close(la.detach().cpu().numpy(), o_las[i].numpy(), 1e-5, 1e-6, f"log_alpha[{i}]")
        _cmp_logs(llogs, ollogs, "alpha")
    finally:
        _rng.set_source(old)
    return worst


@pytest.mark.parametrize("impl", ["tcgen05", "ffma"])
def test_c2_redq_update_matches_oracle(impl):
    """BASELINE configs[1]: REDQ N=10, M=2, 2x256, B=256 -- three updates, then actor + alpha."""
    import super_sac_b200 as ssb

    ssb.set_mlp_impl(impl)
    try:
        _run_state_steps(E=1, N=10, M=2, S=17, A=6, H=256, B=256, steps=3)
    finally:
        ssb.set_mlp_impl("tcgen05")


def test_c2_redq_100_step_drift():
    """100 consecutive REDQ-10 updates (Polyak every 2nd): Adam amplifies early gradient differences, so
    this is where 3xTF32 truncation or a reordered backward would show.  Checked every 10 steps at rtol 1e-4 +
    lr*0.5 absolute (parameters move by up to 100*lr = 3e-2 over the run)."""
    worst = _run_state_steps(E=1, N=10, M=2, S=17, A=6, H=256, B=256, steps=100, check_every=10, final_atol_lr=0.5)
    print(f"[drift] worst |param - oracle| over 100 REDQ-10 updates: {worst:.3e} (lr = 3e-4)")
    assert worst < 2.1 * 3e-4 + 3e-4 * 0.5 + 1e-4 * 3.0   # the per-array gate (incl. the sign-flip allowance) is in cmp_stacks


@pytest.mark.parametrize("popart", [False, True])
def test_c3_sunrise_update_matches_oracle(popart):
    """BASELINE configs[2]: SUNRISE, 5 members x 2 critics, weighted Bellman backups T=20 (member lanes run side by
    side); with PopArt on the TD target goes through td_target_kernel (ART/POP statistics on the device)."""
    _run_state_steps(E=5, N=2, M=2, S=17, A=6, H=256, B=256, steps=2, popart=popart, pop=popart, weight_type="sunrise",
                     temp=20.0, target_delay=1)


def test_c5_offline_afbc_step_matches_oracle():
    """BASELINE configs[4] shapes: N=2, 3x1024 (23-1024-1024-1), B=1024; critic_update with DR3 0.01 + global-norm clip
    40 + priority refresh, Polyak, then offline_actor_update (PER sampling, advantage filter, clip 40, priority
    refresh).  The PER trees are compared after every refresh; sampled indices must coincide with the oracle's on
    >= 99 % of the rows (priorities come from fp32 advantages that agree to ~1e-6, so a prefix-sum boundary can move)."""
    import cuda_util as cu
    import super_sac_b200 as ssb
    from super_sac_b200 import _rng, augmentations, learning, learning_utils as lu

    E, N, M, S, A, H, B, nbuf = 1, 2, 2, 17, 6, 1024, 1024, 20000
    agent, target, o_agent, o_target = tw.make_twins(E, N, S, A, H, seed=5)
    hb = tw.synthetic_state_buffer(nbuf, S, A, 5)
    buf = ssb.replay.ReplayBuffer(nbuf, alpha=0.6, beta=1.0, device="cuda")
    buf.load_experience({"obs": hb["s"]}, hb["a"], hb["r"], {"obs": hb["s1"]}, hb["d"])
    obuf = ro.ReplayOracle(nbuf, alpha=0.6, beta=1.0)
    obuf.load_experience({"obs": hb["s"]}, hb["a"], hb["r"], {"obs": hb["s1"]}, hb["d"])
    c_opt, a_opt, e_opt, las, _ = cu.optimizers(agent, dict(E=E, init_alpha=1e-15))
    o_c, o_a, o_las, _ = tw.oracle_optimizers(o_agent, init_alpha=1e-15)
    aug = augmentations.AugmentationSequence([augmentations.IdentityAug(B)])
    hp = dict(gamma=0.99, critic_clip=40.0, dr3_coeff=0.01, actor_clip=40.0, filter=True)
    kw = _critic_kw(buf, agent, target, c_opt, e_opt, las, B, M, aug, critic_clip=40.0, encoder_clip=40.0,
                    update_priorities=True, dr3_coeff=0.01)
    rng = np.random.default_rng(77)
    nrm = lambda *shape: rng.standard_normal(shape).astype(np.float32)
    sampled = []
    o_per = buf.sample_indices_per

    def rec_per(bs):
        idx, w = o_per(bs)
        sampled.append(idx)
        return idx, w

    buf.sample_indices_per = rec_per
    old = _rng.set_source(_rng.ScriptedSource())
    try:
        for step in range(2):
            # ---- critic update: uniform batch, DR3, clip, priority refresh on the sampled rows --------------------
            src = _rng.ScriptedSource()
            _rng.set_source(src)
            idx = rng.integers(0, nbuf, B)
            eps, subset = nrm(B, A), rng.permutation(N)[:M]
            prio_eps = [nrm(B, A) for _ in range(4)]
            src.push("indices", idx).push("subsets", subset.astype(np.int32)).push("normal", eps)
            for e in prio_eps:
                src.push("normal", e)
            logs, _ = learning.critic_update(**kw)
            assert src.empty()
            batch = tw.state_batch(hb, idx)
            ologs, aux = uo.critic_update(o_agent, o_target, [batch], [dict(eps=tw.t32(eps), subset=[int(x) for x in subset])],
                                          hp, o_las, o_c)
            tw.cmp_stacks(agent._critic_arena, aux["grads"], f"step{step} critic grads (clipped)", RTOL, 1e-5, grad=True)
            tw.cmp_stacks(agent._critic_arena, o_agent.critics, f"step{step} critics", RTOL, 0.0, 3e-4 * 0.05, flip_lr=3e-4)
            _cmp_logs(logs, ologs, f"critic step{step}")
            adv = uo.advantage(o_agent, 0, batch[0], batch[1], [tw.t32(e) for e in prio_eps])
            obuf.update_priorities(idx, (torch.relu(adv) + 1e-4).squeeze(1).numpy())   # fp32, as learning_utils.py:288-295
            gu.assert_close(buf._it_sum.cpu().numpy(), obuf.it_sum.value, 1e-4, 1e-8, f"step{step} sum tree after critic refresh")
            lu.soft_update(target.critics[0], agent.critics[0], 0.005)
            uo.soft_update(o_target.critics.tensors(), o_agent.critics.tensors(), 0.005)
            tw.cmp_stacks(target._critic_arena, o_target.critics, f"step{step} target critics", RTOL, 0.0, 3e-4 * 0.05, flip_lr=3e-4)
            # ---- offline actor update: PER batch, advantage filter, priority refresh -----------------------------
            src = _rng.ScriptedSource()
            _rng.set_source(src)
            u = rng.random(B)
            adv_eps, prio_eps = [nrm(B, A) for _ in range(4)], [nrm(B, A) for _ in range(4)]
            src.push("uniform01", u)
            for e in adv_eps + prio_eps:
                src.push("normal", e)
            sampled.clear()
            alogs = learning.offline_actor_update(
                buffer=buf, agent=agent, actor_optimizer=a_opt, encoder_optimizer=e_opt, batch_size=B, actor_clip=40.0,
                update_encoder=False, encoder_clip=40.0, augmenter=aug, actor_lambda=0.0, aug_mix=0.0,
                premade_replay_dicts=None, per=True, discrete=False, filter_=True)
            assert src.empty()
            got_idx = sampled[0].cpu().numpy()
            _, w, want_idx = obuf.sample(u)
            agree = float(np.mean(got_idx == want_idx))
            assert agree >= 0.99, f"step{step}: PER indices agree on {agree:.4f} of the rows"
            batch = tw.state_batch(hb, got_idx)
            oalogs, aaux = uo.offline_actor_update(o_agent, [batch], [dict(adv_eps=[tw.t32(e) for e in adv_eps])], hp, o_a)
            tw.cmp_stacks(agent._actor_arena, aaux["grads"], f"step{step} actor grads (clipped)", RTOL, 1e-5, grad=True)
            tw.cmp_stacks(agent._actor_arena, o_agent.actors, f"step{step} actors", RTOL, 0.0, 3e-4 * 0.05, flip_lr=3e-4)
            _cmp_logs(alogs, oalogs, f"afbc step{step}")
            adv = uo.advantage(o_agent, 0, batch[0], batch[1], [tw.t32(e) for e in prio_eps])
            obuf.update_priorities(got_idx, (torch.relu(adv) + 1e-4).squeeze(1).numpy())
            gu.assert_close(buf._it_sum.cpu().numpy(), obuf.it_sum.value, 1e-4, 1e-8, f"step{step} sum tree after actor refresh")
    finally:
        buf.sample_indices_per = o_per
        _rng.set_source(old)


class _PixEnc(torch.nn.Module):
    """experiments/dmc/train_dmc_from_pixels.py:15-27: the DMC encoder plugin around BigPixelEncoder."""

    def __init__(self, inner):
        super().__init__()
        self.have_at_least_one_param = torch.nn.Linear(1, 1)
        self.net = inner

    @property
    def embedding_dim(self):
        return self.net.embedding_dim

    def forward(self, obs):
        return self.net(obs["pixels"])

    def forward_rolling(self, obs):
        return self.forward(obs)


def test_c4_drqv2_pixel_update_matches_oracle():
    """BASELINE configs[3]: uint8 9x84x84 frames in the device ring, B=512, Drqv2Aug(pad 4) fused into the gather,
    BigPixelEncoder (50-d), 2 critics 56-1024-1024-1, deterministic actor + TD3 target noise (sigma 0.6, clip 0.3),
    gamma 0.99^3, critic tau 0.01, encoder tau 1.0.  The encoder is differentiated by autograd on both sides (cuDNN
    with TF32 off on the GPU side, ATen-CPU in the oracle), so this checks the whole pixel path: gather + shift + cast
    bit-exact against oracle/aug_oracle.py, then gradients / parameters at rtol 1e-4."""
    import cuda_util as cu
    import super_sac_b200 as ssb
    from super_sac_b200 import _rng, augmentations, learning, learning_utils as lu, nets

    torch.manual_seed(4)
    C, HW, A, H, B, nbuf, N = 9, 84, 6, 1024, 512, 600, 2
    inner = nets.cnns.BigPixelEncoder((C, HW, HW), 50)
    enc = _PixEnc(inner)
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    agent, target, o_agent, o_target = tw.make_twins(1, N, 50, A, H, det=True, seed=4, encoder=enc)
    rng = np.random.default_rng(4)
    s = rng.integers(0, 256, (nbuf, C, HW, HW), dtype=np.uint8)
    s1 = rng.integers(0, 256, (nbuf, C, HW, HW), dtype=np.uint8)
    a = rng.uniform(-1, 1, (nbuf, A)).astype(np.float32)
    r = rng.standard_normal(nbuf).astype(np.float32)
    d = rng.uniform(size=nbuf) < 0.05
    buf = ssb.replay.ReplayBuffer(nbuf, device="cuda")
    buf.load_experience({"pixels": s}, a, r, {"pixels": s1}, d)
    from itertools import chain

    c_opt = torch.optim.Adam(chain(*(c.parameters() for c in agent.critics)), lr=1e-4)
    e_opt = torch.optim.Adam(agent.encoder.parameters(), lr=1e-4)
    o_c = uo.Adam(o_agent.critics.tensors(), lr=1e-4)
    o_e = torch.optim.Adam(o_agent.encoder.parameters(), lr=1e-4)
    las = [torch.tensor([-30.0], device="cuda", requires_grad=True)]
    noise_proc = lu.GaussianExplorationNoise(cu.ActionSpace(A), start_scale=0.6, final_scale=0.1)
    aug = augmentations.AugmentationSequence([augmentations.Drqv2Aug(B)])
    gamma = 0.99**3
    kw = _critic_kw(buf, agent, target, c_opt, e_opt, las, B, 2, aug, gamma=gamma, random_process=noise_proc, noise_clip=0.3,
                    aug_mix=1.0)
    old = _rng.set_source(_rng.ScriptedSource())
    try:
        for step in range(2):
            idx = rng.integers(0, nbuf, B)
            shift = rng.integers(0, 9, (B, 2))
            noise = rng.standard_normal((B, A)).astype(np.float32)
            subset = rng.permutation(N)[:2]
            src = _rng.ScriptedSource()
            _rng.set_source(src)
            src.push("indices", idx).push("shifts", shift.astype(np.int32)).push("normal", noise).push("subsets", subset.astype(np.int32))
            logs, rds = learning.critic_update(**kw)
            assert src.empty()
            o = {"pixels": tw.t32(ao.drq_v2_crop(s[idx], shift))}
            o1 = {"pixels": tw.t32(ao.drq_v2_crop(s1[idx], shift))}
            assert torch.equal(rds[0]["primary_batch"][0]["pixels"].cpu(), o["pixels"]), "gather + shift + cast must be bit-exact"
            assert torch.equal(rds[0]["primary_batch"][3]["pixels"].cpu(), o1["pixels"])
            lu.soft_update(target.critics[0], agent.critics[0], 0.01)
            lu.soft_update(target.encoder, agent.encoder, 1.0)
            batch = (o, tw.t32(a[idx]), tw.t32(r[idx]).reshape(-1, 1), o1, tw.t32(d[idx].astype(np.float32)).reshape(-1, 1))
            hp = dict(gamma=gamma, noise_sigma=0.6, noise_clip=0.3)
            ologs, aux = uo.critic_update(o_agent, o_target, [batch], [dict(eps=None, noise=tw.t32(noise), subset=[int(x) for x in subset])],
                                          hp, [torch.tensor([-30.0])], o_c, o_e)
            uo.soft_update(o_target.critics.tensors(), o_agent.critics.tensors(), 0.01)
            uo.soft_update([p.data for p in o_target.encoder.parameters()], [p.data for p in o_agent.encoder.parameters()], 1.0)
            tw.cmp_stacks(agent._critic_arena, aux["grads"], f"step{step} critic grads", 2e-4, 2e-5, grad=True)
            tw.cmp_stacks(agent._critic_arena, o_agent.critics, f"step{step} critics", RTOL, 0.0, 1e-4 * 0.05, flip_lr=1e-4)
            tw.cmp_stacks(target._critic_arena, o_target.critics, f"step{step} target critics", RTOL, 0.0, 1e-4 * 0.05, flip_lr=1e-4)
            for (k, p), (_, q) in zip(agent.encoder.named_parameters(), o_agent.encoder.named_parameters()):
                if p.grad is None or q.grad is None:
                    assert p.grad is None and q.grad is None, k
                    continue
                want = q.grad.numpy()
                gu.assert_close(p.grad.cpu().numpy(), want, 2e-4, 2e-5 * float(np.abs(want).max()), f"step{step} encoder grad {k}")
                err = (p.detach().cpu() - q.detach()).abs()
                bad = err > 2e-4 * q.detach().abs() + 1e-4 * 0.05
                assert bad.float().mean() <= 1e-4 and float(err.max()) <= 2.2e-4, f"step{step} encoder {k}: {int(bad.sum())} entries off"
            gu.assert_close(logs["losses/critic_overall_loss"], ologs["losses/critic_overall_loss"], 2e-4, 1e-6, "loss")
        # ---- actor update (deterministic actor, TD3 noise on the policy action) ----------------------------------
        a_opt = torch.optim.Adam(chain(*(m.parameters() for m in agent.actors)), lr=1e-4)
        o_a = uo.Adam(o_agent.actors.tensors(), lr=1e-4)
        eps, nz = rng.standard_normal((B, A)).astype(np.float32), rng.standard_normal((B, A)).astype(np.float32)
        src = _rng.ScriptedSource()
        _rng.set_source(src)
        src.push("normal", eps).push("normal", nz)
        learning.online_actor_update(buffer=buf, agent=agent, pop=False, actor_optimizer=a_opt, log_alphas=las, batch_size=B,
                                     clip=None, random_process=noise_proc, noise_clip=0.3, augmenter=aug, aug_mix=1.0,
                                     premade_replay_dicts=rds)
        assert src.empty()
        _, aaux = uo.online_actor_update(o_agent, [batch], [dict(eps=tw.t32(eps), noise=tw.t32(nz))], hp, [torch.tensor([-30.0])], o_a)
        tw.cmp_stacks(agent._actor_arena, aaux["grads"], "actor grads", 2e-4, 2e-5, grad=True)
        tw.cmp_stacks(agent._actor_arena, o_agent.actors, "actors", RTOL, 0.0, 1e-4 * 0.05, flip_lr=1e-4)
    finally:
        _rng.set_source(old)
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
