"""GPU parity of the drop-in update functions against the reference's golden vectors (run with -m gpu on a B200).

The same fixtures that pin the oracle (tests/test_oracle_golden.py) are replayed through
super_sac_b200.learning.{critic_update, online_actor_update, alpha_update} + learning_utils.soft_update with the
reference's indices / subsets / eps / noise injected.  Tolerance: north_star's fp32 rtol 1e-4 (plus small atol for
near-zero entries; post-Adam parameters get an atol tied to the learning rate); Polyak is bit-exact
given equal inputs, so target nets are compared at the same tolerance as the online nets they track.
"""
import numpy as np
import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _cmp_stack(got, want, what, rtol=RTOL, atol=1e-6):
    for n in ("W1", "b1", "W2", "b2", "W3", "b3"):
        gu.assert_close(got[n], want[n], rtol, atol, f"{what}.{n}")


def _cmp_logs(logs, want, what):
    for k, v in want.items():
        k2 = k.replace("|", "/")
        if k2.startswith("gradients/"):
            continue
        assert k2 in logs, f"{what}: missing log key {k2}"
        gu.assert_close(float(logs[k2]), float(v), 2e-4, 2e-5, f"{what} log {k2}")


@pytest.mark.parametrize("impl", ["tcgen05", "ffma"])
@pytest.mark.parametrize("case", gu.UPDATE_CASES)
def test_update_matches_reference(case, impl):
    import cuda_util as cu
    import super_sac_b200 as ssb
    from super_sac_b200 import _rng, augmentations, learning, learning_utils as lu

    ssb.set_mlp_impl(impl)
    fx = gu.load("update_" + case)
    cfg, agent, target = cu.agent_from_fixture(fx)
    E, N, M, B, A = cfg["E"], cfg["N"], cfg["M"], cfg["B"], cfg["A"]
    det = cfg.get("deterministic", False)
    buf = cu.buffer_from_fixture(fx)
    critic_opt, actor_opt, enc_opt, log_alphas, alpha_opts = cu.optimizers(agent, cfg)
    sigma = cfg.get("noise_sigma")
    random_process = None
    if sigma is not None:
        random_process = lu.GaussianExplorationNoise(cu.ActionSpace(A), start_scale=sigma, final_scale=min(sigma, 0.1))
    augmenter = augmentations.AugmentationSequence([augmentations.IdentityAug(B)])
    softmax_w = cfg.get("weight_type") == "softmax" and E > 1

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
                if not det:
                    src.push("normal", r["eps"][i])
                if sigma is not None:
                    src.push("normal", r["noise"][i])
                if softmax_w and not det:
                    for j in range(E):
                        src.push("normal", r["weight_eps"][i][j])
            rec.clear()
            logs, replay_dicts = learning.critic_update(
                buffer=buf, agent=agent, target_agent=target, critic_optimizer=critic_opt, encoder_optimizer=enc_opt,
                log_alphas=log_alphas, batch_size=B, gamma=cfg.get("gamma", 0.99), critic_clip=cfg.get("critic_clip"),
                encoder_clip=cfg.get("encoder_clip"), target_critic_ensemble_n=M,
                weighted_bellman_temp=cfg.get("weight_temp"), weight_type=cfg.get("weight_type"), pop=cfg.get("pop", False),
                augmenter=augmenter, encoder_lambda=0.0, aug_mix=0.0, discrete=False, random_process=random_process,
                noise_clip=cfg.get("noise_clip"), per=False, update_priorities=False, dr3_coeff=cfg.get("dr3_coeff", 0.0))
            assert src.empty(), "not every scripted draw was consumed"
            for i in range(E):
                gu.assert_close(rec["td"][i].cpu().numpy(), fx[f"step{t}/td_target/{i}"], RTOL, 1e-5, f"step{t} td_target[{i}]")
                w = rec["w"][i]
                w = w.cpu().numpy() if torch.is_tensor(w) else np.array(w, dtype=np.float32)
                gu.assert_close(w, fx[f"step{t}/weights/{i}"], RTOL, 1e-5, f"step{t} weights[{i}]")
            _cmp_stack(cu.grads_of(agent._critic_arena), gu.sub(fx, f"step{t}/critic_grads"), f"step{t} critic_grads",
                       rtol=RTOL, atol=2e-7)
            _cmp_logs(logs, gu.sub(fx, f"step{t}/logs"), f"step{t}")
            assert "gradients/critic_random_grad" in logs and "gradients/encoder_criticloss_grad_norm" in logs
            if (t + cfg.get("step0", 0)) % cfg.get("target_delay", 1) == 0:
                for ac, tc in zip(agent.critics, target.critics):
                    lu.soft_update(tc, ac, cfg.get("tau", 0.005))
                lu.soft_update(target.encoder, agent.encoder, cfg.get("encoder_tau", 0.01))
            lr = cfg.get("critic_lr", 3e-4)
            _cmp_stack(cu.stack_of(agent._critic_arena), gu.sub(fx, f"step{t}/critics"), f"step{t} critics", atol=lr * 0.05)
            _cmp_stack(cu.stack_of(target._critic_arena), gu.sub(fx, f"step{t}/target_critics"), f"step{t} target_critics",
                       atol=lr * 0.05)
            want_pop = gu.sub(fx, f"step{t}/popart")
            for i, p in enumerate(agent.popart):
                if not p:
                    continue
                for n in ("mu", "nu", "w", "b"):
                    gu.assert_close(getattr(p, n).cpu().numpy(), want_pop[f"{i}/{n}"], RTOL, 1e-6, f"step{t} popart[{i}].{n}")
                assert int(p._stable) == int(want_pop[f"{i}/stable"])
            if cfg.get("encoder") == "shared":
                want = gu.sub(fx, f"step{t}/encoder")
                for k, v in agent.encoder.state_dict().items():
                    gu.assert_close(v.cpu().numpy(), want[k], RTOL, 1e-4 * 0.05, f"step{t} encoder.{k}")
                want = gu.sub(fx, f"step{t}/target_encoder")
                for k, v in target.encoder.state_dict().items():
                    gu.assert_close(v.cpu().numpy(), want[k], RTOL, 1e-4 * 0.05, f"step{t} target_encoder.{k}")

        # ---- actor update on the last critic batch (reuse_replay_dicts) ----
        src = _rng.ScriptedSource()
        _rng.set_source(src)
        r = gu.sub(fx, "actor/rand")
        for i in range(E):
            src.push("normal", r["eps"][i])
            if sigma is not None:
                src.push("normal", r["noise"][i])
        alogs = learning.online_actor_update(
            buffer=buf, agent=agent, pop=cfg.get("pop", False), actor_optimizer=actor_opt, log_alphas=log_alphas,
            batch_size=B, clip=cfg.get("actor_clip"), random_process=random_process, noise_clip=cfg.get("noise_clip"),
            augmenter=augmenter, aug_mix=0.0, premade_replay_dicts=replay_dicts, per=False, discrete=False,
            use_baseline=False)
        assert src.empty()
        _cmp_stack(cu.grads_of(agent._actor_arena), gu.sub(fx, "actor/grads"), "actor grads", rtol=RTOL, atol=2e-7)
        _cmp_stack(cu.stack_of(agent._actor_arena), gu.sub(fx, "actor/actors"), "actors", atol=3e-4 * 0.05)
        _cmp_logs(alogs, gu.sub(fx, "actor/logs"), "actor")
        if cfg.get("alpha_update", True):
            src = _rng.ScriptedSource()
            _rng.set_source(src)
            r = gu.sub(fx, "alpha/rand")
            if not det:
                for i in range(E):
                    src.push("normal", r["eps"][i])
            llogs = learning.alpha_update(
                buffer=buf, agent=agent, optimizers=alpha_opts, batch_size=B, log_alphas=log_alphas, augmenter=augmenter,
                aug_mix=0.0, target_entropy=-float(A), premade_replay_dicts=replay_dicts, discrete=False)
            for i, la in enumerate(log_alphas):
                gu.assert_close(la.detach().cpu().numpy(), fx[f"alpha/log_alphas/{i}"], 1e-5, 1e-6, f"log_alpha[{i}]")
            _cmp_logs(llogs, gu.sub(fx, "alpha/logs"), "alpha")
    finally:
        lu.compute_td_targets, lu.compute_backup_weights = o_td, o_bw
        _rng.set_source(old_src)
        ssb.set_mlp_impl("tcgen05")
