"""Update-level parity at the BASELINE.json shapes ("parity gaps" 1-2): the drop-in
critic_update -> soft_update -> online_actor_update -> alpha_update (and the offline AFBC step) against the CPU
oracle (oracle/update_oracle.py, pinned to the unmodified reference by tests/test_oracle_golden.py) on injected
draws, at

  C2  REDQ      N=10, M=2, H=256, B=256, obs 17 / act 6                     (+ a 100-step drift run)
  C3  SUNRISE   E=5, N=2, H=256, B=256, weight_type="sunrise", T=20          (PopArt off / on)
  C5  AFBC      N=2, H=1024, B=1024, DR3 0.01, clips 40, PER + priority refresh, offline_actor_update
  C4  DrQv2     u8 9x84x84 frames, B=512, BigPixelEncoder, H=1024, deterministic actor + TD3 noise, Drqv2Aug

Tolerance (north_star): rtol 1e-4 for TD targets / losses / gradients / post-step parameters, with an absolute floor
of 1e-5 of each array's largest entry for gradients (sums over B*H products whose terms cancel) and lr*0.05 for
post-Adam parameters (Adam turns a gradient entry that is rounding noise around zero into a step of up to +-lr:
at most 2e-4 of an array's entries may miss the tolerance, by <= 2.1 lr).

Well-posedness at these sizes.  One update evaluates millions of ReLU pre-activations; a pre-activation inside fp32
rounding of zero may land on either side of the ReLU in two correct fp32 implementations (any two BLAS libraries
differ like that), and in the backward that decision moves every first-layer gradient entry by ~1/sqrt(B*H) -- far
above 1e-4.  So each batch is drawn from more candidate rows than it needs and the rows with such a pre-activation
(|z| < 2e-6 of the magnitude of its summands, evaluated by the oracle) are left out: what is compared is then a
smooth function of the inputs and rtol 1e-4 is meaningful (twin_util.rows_ambiguous).
"""
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


def _stack_norm(stack):
    """sqrt(sum g^2) over every array of an oracle gradient stack (learning_utils.py:95-106 get_grad_norm)."""
    return float(sum(float((getattr(stack, n).double() ** 2).sum()) for n in uo.PARAM_NAMES)) ** 0.5


def _cmp_sum_tree(buf, obuf, what, delta_adv=3e-5):
    """PER sum tree after a priority refresh.  Leaves are (relu(A) + 1e-4)^0.6 with A = Q(s,a) - mean Q(s,a'~pi) an fp32
    difference of O(1) values: its absolute error delta_adv (a few 1e-6 per Q value) is amplified by
    d p / d A = 0.6 p^(-2/3) near the 1e-4 floor, so leaves are held to rtol 1e-4 + 0.6 p^(-2/3) delta_adv and the root
    (the total mass the sampler draws against) to rtol 1e-4."""
    got, want = buf._it_sum.cpu().numpy(), obuf.it_sum.value
    cap = obuf.it_sum.capacity
    gl, wl = got[cap:], want[cap:]
    tol = 1e-4 * np.abs(wl) + 0.6 * np.maximum(wl, 1e-12) ** (-2.0 / 3.0) * delta_adv
    bad = np.abs(gl - wl) > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} leaves off, worst {float(np.abs(gl - wl)[bad].max()):.3e}"
    gu.assert_close(got[1], want[1], 1e-4, 0.0, f"{what} (root)")


def _nets_ambiguous(stack, nets, x):
    bad = torch.zeros(x.shape[0], dtype=torch.bool)
    for g in nets:
        bad |= tw.rows_ambiguous(stack, g, x)
    return bad


def _run_state_steps(E, N, M, S, A, H, B, steps, popart=False, pop=False, weight_type=None, temp=None, seed=0,
                     target_delay=2, check_every=1):
    """critic_update (+Polyak) x steps, then an actor and a temperature update (each on a batch of its own);
    GPU vs oracle after every step (check_every = 1: per-entry parity incl. the gradients) or every ``check_every``
    steps (the drift run).  Returns the worst post-step parameter error seen."""
    import cuda_util as cu
    import super_sac_b200 as ssb
    from super_sac_b200 import _rng, augmentations, learning, learning_utils as lu

    nbuf = 5000
    agent, target, o_agent, o_target = tw.make_twins(E, N, S, A, H, popart=popart, seed=seed)
    hb = tw.synthetic_state_buffer(nbuf, S, A, seed)
    buf = ssb.replay.ReplayBuffer(nbuf, device="cuda")
    buf.load_experience({"obs": hb["s"]}, hb["a"], hb["r"], {"obs": hb["s1"]}, hb["d"])
    c_opt, a_opt, e_opt, las, al_opts = cu.optimizers(agent, dict(E=E))
    o_c, o_a, o_las, o_al = tw.oracle_optimizers(o_agent)
    aug = augmentations.AugmentationSequence([augmentations.IdentityAug(B)])
    hp = dict(gamma=0.99, pop=pop, weight_type=weight_type, weight_temp=temp)
    kw = _critic_kw(buf, agent, target, c_opt, e_opt, las, B, M, aug, pop=pop, weight_type=weight_type,
                    weighted_bellman_temp=temp)
    rng = np.random.default_rng(seed + 100)
    old = _rng.set_source(_rng.ScriptedSource())
    worst, dropped, strict = 0.0, 0, check_every == 1
    try:
        for t in range(steps):
            src = _rng.ScriptedSource()
            _rng.set_source(src)
            batches, rands = [], []
            for i in range(E):
                cand = rng.integers(0, nbuf, 2 * B)
                eps = rng.standard_normal((2 * B, A)).astype(np.float32)
                cb = tw.state_batch(hb, cand)
                bad = _nets_ambiguous(o_agent.critics, range(i * N, (i + 1) * N), torch.cat((cb[0]["obs"], cb[1]), dim=-1))
                keep = tw.keep_rows(bad, B)
                dropped += int(bad[: keep[-1] + 1].sum())
                idx, eps = cand[keep], eps[keep]
                subset = rng.permutation(N)[:M]
                src.push("indices", idx).push("subsets", subset.astype(np.int32)).push("normal", eps)
                batches.append(tw.state_batch(hb, idx))
                rands.append(dict(eps=tw.t32(eps), subset=[int(x) for x in subset]))
            logs, _ = learning.critic_update(**kw)
            assert src.empty(), "not every scripted draw was consumed"
            ologs, aux = uo.critic_update(o_agent, o_target, batches, rands, hp, o_las, o_c)
            if t % target_delay == 0:
                for ac, tc in zip(agent.critics, target.critics):
                    lu.soft_update(tc, ac, 0.005)
                uo.soft_update(o_target.critics.tensors(), o_agent.critics.tensors(), 0.005)
            if strict:   # single-step parity: the gradients themselves, then the post-Adam / post-Polyak parameters
                tw.cmp_stacks(agent._critic_arena, aux["grads"], f"step{t} critic grads", RTOL, 1e-5, grad=True)
                _cmp_logs(logs, ologs, f"step{t}")
                if E == 1:   # the logged gradient norm (learning.py:134): with one member the "random" member is member 0
                    gu.assert_close(float(logs["gradients/critic_random_grad"]), _stack_norm(aux["grads"]), RTOL, 1e-7,
                                    f"step{t} log gradients/critic_random_grad")
                tw.cmp_stacks(agent._critic_arena, o_agent.critics, f"step{t} critics", RTOL, 0.0, 3e-4 * 0.05, noise_lr=3e-4)
                tw.cmp_stacks(target._critic_arena, o_target.critics, f"step{t} target critics", RTOL, 0.0, 3e-4 * 0.05, noise_lr=3e-4)
                tw.resync(agent, target, o_agent, o_target)
            elif t % check_every == 0 or t == steps - 1:
                # drift run: >= 99.99 % of all entries within rtol 1e-4 + lr/10, none further than lr
                for ar, st, nm in ((agent._critic_arena, o_agent.critics, "critics"), (target._critic_arena, o_target.critics, "targets")):
                    frac, w = tw.frac_within(ar, st, RTOL, 3e-4 * 0.1)
                    assert frac >= 0.9999 and w <= 3e-4, f"step{t} {nm}: {frac:.5f} of the entries within tolerance, worst {w:.3e}"
                    worst = max(worst, w)
            for i, p in enumerate(agent.popart):
                if p:
                    op = o_agent.popart[i]
                    for n in ("mu", "nu", "w", "b"):
                        gu.assert_close(getattr(p, n).cpu().numpy(), getattr(op, n).numpy(), RTOL, 1e-6, f"step{t} popart[{i}].{n}")
        # ---- actor update on a batch of its own (premade_replay_dicts=None), rows unambiguous for the actor on s and
        # for the critics on (s, pi(s)) ---------------------------------------------------------------------------
        src = _rng.ScriptedSource()
        _rng.set_source(src)
        abatches, arands = [], []
        for i in range(E):
            cand = rng.integers(0, nbuf, 2 * B)
            eps = rng.standard_normal((2 * B, A)).astype(np.float32)
            cb = tw.state_batch(hb, cand)
            s = cb[0]["obs"]
            with torch.no_grad():
                a_pi, _, _ = uo.actor_sample(o_agent, i, s, tw.t32(eps))
            bad = _nets_ambiguous(o_agent.actors, [i], s) | _nets_ambiguous(o_agent.critics, range(i * N, (i + 1) * N), torch.cat((s, a_pi), dim=-1))
            keep = tw.keep_rows(bad, B)
            src.push("indices", cand[keep]).push("normal", eps[keep])
            abatches.append(tw.state_batch(hb, cand[keep]))
            arands.append(dict(eps=tw.t32(eps[keep])))
        alogs = learning.online_actor_update(buffer=buf, agent=agent, pop=pop, actor_optimizer=a_opt, log_alphas=las,
                                             batch_size=B, clip=None, random_process=None, noise_clip=None, augmenter=aug,
                                             aug_mix=0.0, premade_replay_dicts=None)
        assert src.empty()
        oalogs, aaux = uo.online_actor_update(o_agent, abatches, arands, hp, o_las, o_a)
        tw.cmp_stacks(agent._actor_arena, aaux["grads"], "actor grads", RTOL, 1e-5, grad=True)
        if E == 1:   # learning.py:417-419
            gu.assert_close(float(alogs["gradients/random_actor_online_grad"]), _stack_norm(aaux["grads"]), RTOL, 1e-7,
                            "log gradients/random_actor_online_grad")
        tw.cmp_stacks(agent._actor_arena, o_agent.actors, "actors", RTOL, 0.0, 3e-4 * 0.05, noise_lr=3e-4)
        _cmp_logs(alogs, oalogs, "actor")
        # ---- temperature update (no backward through the networks) ------------------------------------------------
        src = _rng.ScriptedSource()
        _rng.set_source(src)
        lbatches, lrands = [], []
        for i in range(E):
            idx = rng.integers(0, nbuf, B)
            eps = rng.standard_normal((B, A)).astype(np.float32)
            src.push("indices", idx).push("normal", eps)
            lbatches.append(tw.state_batch(hb, idx))
            lrands.append(dict(eps=tw.t32(eps)))
        llogs = learning.alpha_update(buffer=buf, agent=agent, optimizers=al_opts, batch_size=B, log_alphas=las, augmenter=aug,
                                      aug_mix=0.0, target_entropy=-float(A), premade_replay_dicts=None, discrete=False)
        assert src.empty()
        ollogs = uo.alpha_update(o_agent, lbatches, lrands, o_las, o_al, -float(A))
        for i, la in enumerate(las):
            gu.assert_close(la.detach().cpu().numpy(), o_las[i].numpy(), 1e-5, 1e-6, f"log_alpha[{i}]")
        _cmp_logs(llogs, ollogs, "alpha")
        print(f"[parity] E={E} N={N} H={H} B={B}: {dropped} candidate rows left out over {steps} critic updates")
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
    this is where 3xTF32 truncation or a reordered backward would show.  Checked every 10 steps: >= 99.99 % of all
    parameters within rtol 1e-4 + lr/10 of the oracle's and none further than lr (they move by up to 100*lr = 3e-2;
    measured on B200: worst |difference| 1.9e-5 after 100 updates)."""
    worst = _run_state_steps(E=1, N=10, M=2, S=17, A=6, H=256, B=256, steps=100, check_every=10)
    print(f"[drift] worst |param - oracle| over 100 REDQ-10 updates: {worst:.3e} (lr = 3e-4)")


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
            cand, eps = rng.integers(0, nbuf, 3 * B), nrm(3 * B, A)
            cb = tw.state_batch(hb, cand)
            with torch.no_grad():   # both forwards of the DR3 update: (s, a) and (s1, a1 ~ pi(s1))
                a1, _, _ = uo.actor_sample(o_agent, 0, cb[3]["obs"], tw.t32(eps))
            bad = (_nets_ambiguous(o_agent.critics, range(N), torch.cat((cb[0]["obs"], cb[1]), dim=-1))
                   | _nets_ambiguous(o_agent.critics, range(N), torch.cat((cb[3]["obs"], a1), dim=-1)))
            keep = tw.keep_rows(bad, B)
            idx, eps = cand[keep], eps[keep]
            subset = rng.permutation(N)[:M]
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
            tw.cmp_stacks(agent._critic_arena, o_agent.critics, f"step{step} critics", RTOL, 0.0, 3e-4 * 0.05, noise_lr=3e-4)
            _cmp_logs(logs, ologs, f"critic step{step}")
            adv = uo.advantage(o_agent, 0, batch[0], batch[1], [tw.t32(e) for e in prio_eps])
            obuf.update_priorities(idx, (torch.relu(adv) + 1e-4).squeeze(1).numpy())   # fp32, as learning_utils.py:288-295
            _cmp_sum_tree(buf, obuf, f"step{step} sum tree after critic refresh")
            lu.soft_update(target.critics[0], agent.critics[0], 0.005)
            uo.soft_update(o_target.critics.tensors(), o_agent.critics.tensors(), 0.005)
            tw.cmp_stacks(target._critic_arena, o_target.critics, f"step{step} target critics", RTOL, 0.0, 3e-4 * 0.05, noise_lr=3e-4)
            tw.resync(agent, target, o_agent, o_target)
            # ---- offline actor update: PER batch, advantage filter, priority refresh -----------------------------
            src = _rng.ScriptedSource()
            _rng.set_source(src)
            u_c = rng.random(3 * B)
            (s_c, *_), _, idx_c = obuf.sample(u_c)
            keep = tw.keep_rows(_nets_ambiguous(o_agent.actors, [0], tw.t32(s_c["obs"])), B)
            u = u_c[keep]
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
            tw.cmp_stacks(agent._actor_arena, o_agent.actors, f"step{step} actors", RTOL, 0.0, 3e-4 * 0.05, noise_lr=3e-4)
            _cmp_logs(alogs, oalogs, f"afbc step{step}")
            adv = uo.advantage(o_agent, 0, batch[0], batch[1], [tw.t32(e) for e in prio_eps])
            obuf.update_priorities(got_idx, (torch.relu(adv) + 1e-4).squeeze(1).numpy())
            _cmp_sum_tree(buf, obuf, f"step{step} sum tree after actor refresh")
            tw.resync(agent, target, o_agent, o_target)
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
    gamma 0.99^3, critic tau 0.01, encoder tau 1.0.  The encoder is a PyTorch plugin differentiated by autograd on both
    sides (cuDNN with TF32 off here, ATen-CPU in the oracle).  Checked: gather + shift + cast bit-exact against
    oracle/aug_oracle.py; critic / actor gradients, post-step parameters and the gradient handed to the plugin
    (dL/ds_rep) at rtol 1e-4 (2e-4 where the two encoders' forward values enter); the plugin's own parameter gradients
    at 2 % relative L2 (two conv libraries disagree on some of the ~1e8 ReLU decisions inside the conv stack)."""
    from itertools import chain

    import cuda_util as cu
    import super_sac_b200 as ssb
    from super_sac_b200 import _rng, augmentations, learning, learning_utils as lu, nets

    torch.manual_seed(4)
    C, HW, A, H, B, nbuf, N = 9, 84, 6, 1024, 512, 900, 2
    enc = _PixEnc(nets.cnns.BigPixelEncoder((C, HW, HW), 50))
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
    c_opt = torch.optim.Adam(chain(*(c.parameters() for c in agent.critics)), lr=1e-4)
    e_opt = torch.optim.Adam(agent.encoder.parameters(), lr=1e-4)
    o_c = uo.Adam(o_agent.critics.tensors(), lr=1e-4)
    o_e = torch.optim.Adam(o_agent.encoder.parameters(), lr=1e-4)
    las = [torch.tensor([-30.0], device="cuda", requires_grad=True)]
    noise_proc = lu.GaussianExplorationNoise(cu.ActionSpace(A), start_scale=0.6, final_scale=0.1)
    aug = augmentations.AugmentationSequence([augmentations.Drqv2Aug(B)])
    gamma = 0.99**3
    hp = dict(gamma=gamma, noise_sigma=0.6, noise_clip=0.3)
    kw = _critic_kw(buf, agent, target, c_opt, e_opt, las, B, 2, aug, gamma=gamma, random_process=noise_proc, noise_clip=0.3,
                    aug_mix=1.0)
    NC = B + B // 2   # candidate rows per batch
    old = _rng.set_source(_rng.ScriptedSource())
    handed = []   # what the CUDA path hands to autograd: (s_rep, dL/ds_rep)
    o_backward = torch.autograd.backward

    def rec_backward(tensors, grad_tensors=None, *a_, **k_):
        handed.append([g.detach().clone() for g in grad_tensors])
        return o_backward(tensors, grad_tensors, *a_, **k_)

    try:
        for step in range(2):
            cand, shift = rng.integers(0, nbuf, NC), rng.integers(0, 9, (NC, 2))
            with torch.no_grad():
                rep = o_agent.encode({"pixels": tw.t32(ao.drq_v2_crop(s[cand], shift))})
            keep = tw.keep_rows(_nets_ambiguous(o_agent.critics, range(N), torch.cat((rep, tw.t32(a[cand])), dim=-1)), B)
            idx, shift = cand[keep], shift[keep]
            noise = rng.standard_normal((B, A)).astype(np.float32)
            subset = rng.permutation(N)[:2]
            src = _rng.ScriptedSource()
            _rng.set_source(src)
            src.push("indices", idx).push("shifts", shift.astype(np.int32)).push("normal", noise).push("subsets", subset.astype(np.int32))
            handed.clear()
            torch.autograd.backward = rec_backward
            try:
                logs, rds = learning.critic_update(**kw)
            finally:
                torch.autograd.backward = o_backward
            assert src.empty()
            o = {"pixels": tw.t32(ao.drq_v2_crop(s[idx], shift))}
            o1 = {"pixels": tw.t32(ao.drq_v2_crop(s1[idx], shift))}
            assert torch.equal(rds[0]["primary_batch"][0]["pixels"].cpu(), o["pixels"]), "gather + shift + cast must be bit-exact"
            assert torch.equal(rds[0]["primary_batch"][3]["pixels"].cpu(), o1["pixels"])
            lu.soft_update(target.critics[0], agent.critics[0], 0.01)
            lu.soft_update(target.encoder, agent.encoder, 1.0)
            batch = (o, tw.t32(a[idx]), tw.t32(r[idx]).reshape(-1, 1), o1, tw.t32(d[idx].astype(np.float32)).reshape(-1, 1))
            ologs, aux = uo.critic_update(o_agent, o_target, [batch], [dict(eps=None, noise=tw.t32(noise), subset=[int(x) for x in subset])],
                                          hp, [torch.tensor([-30.0])], o_c, o_e)
            uo.soft_update(o_target.critics.tensors(), o_agent.critics.tensors(), 0.01)
            uo.soft_update([p.data for p in o_target.encoder.parameters()], [p.data for p in o_agent.encoder.parameters()], 1.0)
            tw.cmp_stacks(agent._critic_arena, aux["grads"], f"step{step} critic grads", 2e-4, 2e-5, grad=True)
            tw.cmp_stacks(agent._critic_arena, o_agent.critics, f"step{step} critics", RTOL, 0.0, 1e-4 * 0.05, noise_lr=1e-4)
            tw.cmp_stacks(target._critic_arena, o_target.critics, f"step{step} target critics", RTOL, 0.0, 1e-4 * 0.05, noise_lr=1e-4)
            # the gradient the CUDA path hands to the encoder plugin (dL/ds_rep, summed over both critics): strict
            want = aux["s_rep_grad"][0].numpy()
            gu.assert_close(handed[0][0].cpu().numpy(), want, 2e-4, 2e-5 * float(np.abs(want).max()), f"step{step} dL/ds_rep")
            # the plugin's own backward (cuDNN here, ATen-CPU in the oracle; ~100 M ReLU decisions inside the conv stack, so
            # single entries differ by per cent, see the module docstring): per-tensor relative L2 error, loose step check
            for (k, p), (_, q) in zip(agent.encoder.named_parameters(), o_agent.encoder.named_parameters()):
                if p.grad is None or q.grad is None:
                    assert p.grad is None and q.grad is None, k
                    continue
                rel = float((p.grad.cpu() - q.grad).norm() / q.grad.norm().clamp_min(1e-30))
                assert rel <= 2e-2, f"step{step} encoder grad {k}: relative L2 error {rel:.3e}"
                assert float((p.detach().cpu() - q.detach()).abs().max()) <= 2.2e-4 * (step + 1), f"step{step} encoder {k}"
            gu.assert_close(logs["losses/critic_overall_loss"], ologs["losses/critic_overall_loss"], 2e-4, 1e-6, "loss")
            tw.resync(agent, target, o_agent, o_target)
        # ---- actor update (deterministic actor, TD3 noise on the policy action), on a batch of its own ------------------
        a_opt = torch.optim.Adam(chain(*(m.parameters() for m in agent.actors)), lr=1e-4)
        o_a = uo.Adam(o_agent.actors.tensors(), lr=1e-4)
        cand, shift = rng.integers(0, nbuf, NC), rng.integers(0, 9, (NC, 2))
        eps, nz = rng.standard_normal((NC, A)).astype(np.float32), rng.standard_normal((NC, A)).astype(np.float32)
        with torch.no_grad():
            rep = o_agent.encode({"pixels": tw.t32(ao.drq_v2_crop(s[cand], shift))})
            a_pi, _, _ = uo.actor_sample(o_agent, 0, rep, None)
            a_pi = uo.gaussian_noise_clamp(a_pi + 1e-4 * tw.t32(eps), tw.t32(nz), 0.6, 0.3, -1.0, 1.0)
        keep = tw.keep_rows(_nets_ambiguous(o_agent.actors, [0], rep) | _nets_ambiguous(o_agent.critics, range(N), torch.cat((rep, a_pi), dim=-1)), B)
        idx, shift, eps, nz = cand[keep], shift[keep], eps[keep], nz[keep]
        src = _rng.ScriptedSource()
        _rng.set_source(src)
        src.push("indices", idx).push("shifts", shift.astype(np.int32)).push("normal", eps).push("normal", nz)
        learning.online_actor_update(buffer=buf, agent=agent, pop=False, actor_optimizer=a_opt, log_alphas=las, batch_size=B,
                                     clip=None, random_process=noise_proc, noise_clip=0.3, augmenter=aug, aug_mix=1.0,
                                     premade_replay_dicts=None)
        assert src.empty()
        o = {"pixels": tw.t32(ao.drq_v2_crop(s[idx], shift))}
        batch = (o, tw.t32(a[idx]), None, None, None)
        _, aaux = uo.online_actor_update(o_agent, [batch], [dict(eps=tw.t32(eps), noise=tw.t32(nz))], hp, [torch.tensor([-30.0])], o_a)
        tw.cmp_stacks(agent._actor_arena, aaux["grads"], "actor grads", 2e-4, 2e-5, grad=True)
        tw.cmp_stacks(agent._actor_arena, o_agent.actors, "actors", RTOL, 0.0, 1e-4 * 0.05, noise_lr=1e-4)
    finally:
        _rng.set_source(old)
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
