"""Pin the CPU oracle against golden vectors produced by the unmodified reference (CPU only, no GPU).

Tolerances: both sides are torch-CPU fp32 with the same GEMM library, but the oracle's explicit backward and
the reference's autograd sum in different orders, so 'equal' means a few fp32 ulps, not bit-equal.
Replay / segment-tree / augmentation results are integer / byte / float64-tree work: bit-exact.
"""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import aug_oracle as ao
from oracle import discrete_oracle as do
from oracle import replay_oracle as ro
from oracle import update_oracle as uo

RTOL, ATOL = 2e-5, 2e-6


def _cmp_stack(stack, want, what, rtol=RTOL, atol=ATOL):
    for n in uo.PARAM_NAMES:
        gu.assert_close(getattr(stack, n).numpy(), want[n], rtol, atol, f"{what}.{n}")


def _cmp_logs(logs, want, what):
    for k, v in want.items():
        k2 = k.replace("|", "/")
        if k2.startswith("gradients/"):
            continue  # random.choice(...) picks: not part of the arithmetic
        assert k2 in logs, f"{what}: missing log key {k2}"
        gu.assert_close(logs[k2], float(v), 1e-4, 1e-5, f"{what} log {k2}")


@pytest.mark.parametrize("case", gu.UPDATE_CASES)
def test_update_oracle_matches_reference(case):
    check_update_oracle(gu.load("update_" + case))


def check_update_oracle(fx):
    torch.set_num_threads(1)
    cfg, agent, target = gu.oracle_agents(fx)
    E = cfg["E"]
    hp = gu.hp_from(cfg)
    log_alphas = gu.log_alphas_from(cfg)
    critic_opt = uo.Adam(agent.critics.tensors(), lr=cfg.get("critic_lr", 3e-4), weight_decay=cfg.get("critic_l2", 0.0))
    actor_opt = uo.Adam(agent.actors.tensors(), lr=cfg.get("actor_lr", 3e-4))
    enc_opt = None
    if agent.encoder is not None:
        enc_opt = torch.optim.Adam(agent.encoder.parameters(), lr=cfg.get("encoder_lr", 1e-4))
    alpha_opts = [uo.Adam([la], lr=cfg.get("alpha_lr", 1e-4), betas=(0.5, 0.999)) for la in log_alphas]

    batches = None
    for t in range(cfg["steps"]):
        rands = gu.rands_from(fx, f"step{t}/rand", E)
        idx = fx[f"step{t}/rand/idx"]
        batches = [gu.batch_from(fx, idx[i]) for i in range(E)]
        logs, aux = uo.critic_update(agent, target, batches, rands, hp, log_alphas, critic_opt, enc_opt)
        for i in range(E):
            gu.assert_close(aux["td_target"][i].numpy(), fx[f"step{t}/td_target/{i}"], RTOL, ATOL, f"step{t} td_target[{i}]")
            w = aux["weights"][i]
            w = w.numpy() if torch.is_tensor(w) else np.array(w, dtype=np.float32)
            gu.assert_close(w, fx[f"step{t}/weights/{i}"], RTOL, ATOL, f"step{t} weights[{i}]")
        _cmp_stack(aux["grads"], gu.sub(fx, f"step{t}/critic_grads"), f"step{t} critic_grads", rtol=1e-4, atol=1e-7)
        _cmp_logs(logs, gu.sub(fx, f"step{t}/logs"), f"step{t}")
        if (t + cfg.get("step0", 0)) % cfg.get("target_delay", 1) == 0:
            uo.soft_update(target.critics.tensors(), agent.critics.tensors(), cfg.get("tau", 0.005))
            if agent.encoder is not None:
                uo.soft_update([p.data for p in target.encoder.parameters()],
                               [p.data for p in agent.encoder.parameters()], cfg.get("encoder_tau", 0.01))
        # post-Adam parameters: atol tied to the learning rate (Adam normalises the step to ~lr)
        _cmp_stack(agent.critics, gu.sub(fx, f"step{t}/critics"), f"step{t} critics", rtol=1e-5, atol=3e-4 * 2e-2)
        _cmp_stack(target.critics, gu.sub(fx, f"step{t}/target_critics"), f"step{t} target_critics", rtol=1e-5, atol=1e-6)
        want_pop = gu.sub(fx, f"step{t}/popart")
        for i, p in enumerate(agent.popart):
            if p is None:
                continue
            for n in ("mu", "nu", "w", "b"):
                gu.assert_close(getattr(p, n).numpy(), want_pop[f"{i}/{n}"], 1e-5, 1e-7, f"step{t} popart[{i}].{n}")
            assert int(p.stable) == int(want_pop[f"{i}/stable"])
        if agent.encoder is not None:
            want = gu.sub(fx, f"step{t}/encoder")
            for k, v in agent.encoder.state_dict().items():
                gu.assert_close(v.numpy(), want[k], 1e-5, 1e-4 * 2e-2, f"step{t} encoder.{k}")

    rands = gu.rands_from(fx, "actor/rand", E)
    hp_a = dict(hp)
    logs, aux = uo.online_actor_update(agent, batches, rands, hp_a, log_alphas, actor_opt)
    _cmp_stack(aux["grads"], gu.sub(fx, "actor/grads"), "actor grads", rtol=1e-4, atol=1e-7)
    _cmp_stack(agent.actors, gu.sub(fx, "actor/actors"), "actors", rtol=1e-5, atol=3e-4 * 2e-2)
    _cmp_logs(logs, gu.sub(fx, "actor/logs"), "actor")
    if cfg.get("alpha_update", True):
        rands = gu.rands_from(fx, "alpha/rand", E)
        logs = uo.alpha_update(agent, batches, rands, log_alphas, alpha_opts, target_entropy=-float(cfg["A"]))
        for i, la in enumerate(log_alphas):
            gu.assert_close(la.numpy(), fx[f"alpha/log_alphas/{i}"], 1e-6, 1e-7, f"log_alpha[{i}]")
        _cmp_logs(logs, gu.sub(fx, "alpha/logs"), "alpha")


@pytest.mark.parametrize("case", gu.DISCRETE_CASES + [gu.DISCRETE_ENCODER_CASE])
def test_discrete_oracle_matches_reference(case):
    """SAC-Discrete: the restated discrete branches against the unmodified reference's outputs."""
    check_discrete_oracle(gu.load("update_" + case))


def check_discrete_oracle(fx):
    torch.set_num_threads(1)
    cfg, agent, target = gu.discrete_oracle_agents(fx)
    E = cfg["E"]
    hp = gu.hp_from(cfg)
    log_alphas = gu.log_alphas_from(cfg)
    critic_opt = uo.Adam(agent.critics.tensors(), lr=cfg.get("critic_lr", 3e-4))
    actor_opt = uo.Adam(agent.actors.tensors(), lr=cfg.get("actor_lr", 3e-4))
    alpha_opts = [uo.Adam([la], lr=cfg.get("alpha_lr", 1e-4), betas=(0.5, 0.999)) for la in log_alphas]
    enc_opt = torch.optim.Adam(agent.encoder.parameters(), lr=1e-4) if agent.encoder is not None else None
    batches = None
    for t in range(cfg["steps"]):
        idx, subsets = fx[f"step{t}/rand/idx"], fx[f"step{t}/rand/subsets"]
        batches = [gu.batch_from(fx, idx[i]) for i in range(E)]
        logs, aux = do.critic_update(agent, target, batches, [[int(x) for x in subsets[i]] for i in range(E)], hp,
                                     log_alphas, critic_opt, enc_opt)
        for i in range(E):
            gu.assert_close(aux["td_target"][i].numpy(), fx[f"step{t}/td_target/{i}"], RTOL, ATOL, f"step{t} td_target[{i}]")
            w = aux["weights"][i]
            w = w.numpy() if torch.is_tensor(w) else np.array(w, dtype=np.float32)
            gu.assert_close(w, fx[f"step{t}/weights/{i}"], RTOL, ATOL, f"step{t} weights[{i}]")
        _cmp_stack(aux["grads"], gu.sub(fx, f"step{t}/critic_grads"), f"step{t} critic_grads", rtol=1e-4, atol=1e-7)
        _cmp_logs(logs, gu.sub(fx, f"step{t}/logs"), f"step{t}")
        uo.soft_update(target.critics.tensors(), agent.critics.tensors(), cfg.get("tau", 0.005))
        _cmp_stack(agent.critics, gu.sub(fx, f"step{t}/critics"), f"step{t} critics", rtol=1e-5, atol=3e-4 * 2e-2)
        _cmp_stack(target.critics, gu.sub(fx, f"step{t}/target_critics"), f"step{t} target_critics", rtol=1e-5, atol=1e-6)
        if agent.encoder is not None:
            uo.soft_update([p.data for p in target.encoder.parameters()], [p.data for p in agent.encoder.parameters()],
                           cfg.get("encoder_tau", 0.01))
            for which, enc in (("encoder", agent.encoder), ("target_encoder", target.encoder)):
                want = gu.sub(fx, f"step{t}/{which}")
                for k, v in enc.state_dict().items():
                    gu.assert_close(v.numpy(), want[k], 1e-5, 1e-4 * 2e-2, f"step{t} {which}.{k}")
        want_pop = gu.sub(fx, f"step{t}/popart")
        for i, p in enumerate(agent.popart):
            if p is None:
                continue
            for n in ("mu", "nu", "w", "b"):
                gu.assert_close(getattr(p, n).numpy(), want_pop[f"{i}/{n}"], 1e-5, 1e-7, f"step{t} popart[{i}].{n}")
    logs, aux = do.online_actor_update(agent, batches, hp, log_alphas, actor_opt)
    _cmp_stack(aux["grads"], gu.sub(fx, "actor/grads"), "actor grads", rtol=1e-4, atol=1e-7)
    _cmp_stack(agent.actors, gu.sub(fx, "actor/actors"), "actors", rtol=1e-5, atol=3e-4 * 2e-2)
    _cmp_logs(logs, gu.sub(fx, "actor/logs"), "actor")
    logs = do.alpha_update(agent, batches, log_alphas, alpha_opts, float(fx["alpha/target_entropy"]))
    for i, la in enumerate(log_alphas):
        gu.assert_close(la.numpy(), fx[f"alpha/log_alphas/{i}"], 1e-6, 1e-7, f"log_alpha[{i}]")
    _cmp_logs(logs, gu.sub(fx, "alpha/logs"), "alpha")


def test_discrete_afbc_oracle_matches_reference():
    """Offline (AFBC) actor update of a discrete agent + indirect advantage + priority refresh against the reference."""
    torch.set_num_threads(1)
    fx = gu.load("discrete_afbc")
    cfg, agent, _ = gu.discrete_oracle_agents(fx, with_target=False)
    E = cfg["E"]
    idx = fx["rand/idx"]
    batches = [gu.batch_from(fx, idx[i]) for i in range(E)]
    for i in range(E):
        adv = do.advantage(agent, i, batches[i][0]["obs"], batches[i][1])
        gu.assert_close(adv.numpy(), fx[f"adv/{i}"], RTOL, ATOL, f"adv[{i}]")
    actor_opt = uo.Adam(agent.actors.tensors(), lr=3e-4)
    logs, aux = do.offline_actor_update(agent, batches, dict(actor_clip=cfg["actor_clip"]), actor_opt)
    _cmp_stack(aux["grads"], gu.sub(fx, "actor/grads"), "actor grads", rtol=1e-4, atol=1e-7)
    _cmp_stack(agent.actors, gu.sub(fx, "actor/actors"), "actors", rtol=1e-5, atol=3e-4 * 2e-2)
    _cmp_logs(logs, gu.sub(fx, "actor/logs"), "offline actor")
    pr = do.priorities(agent, int(fx["priorities/member"]), batches[-1])
    gu.assert_close(pr.numpy(), fx["priorities/values"], 1e-5, 1e-6, "priorities")


def test_replay_oracle_matches_reference():
    fx = gu.load("replay_per")
    buf = ro.ReplayOracle(50, alpha=0.6, beta=0.7)
    for k in range(int(fx["n_ops"])):
        op = gu.sub(fx, f"op{k}")
        if str(op["kind"]) == "push1":
            buf.push({"obs": op["s"]}, op["a"], float(op["r"]), {"obs": op["s1"]}, bool(op["d"]))
        else:
            buf.push({"obs": op["s"]}, op["a"], op["r"], {"obs": op["s1"]}, op["d"], priorities=op["priorities"])
    ap = gu.sub(fx, "after_push")
    st = buf.storage
    assert np.array_equal(st.s["obs"], ap["s"]) and np.array_equal(st.s1["obs"], ap["s1"])
    assert np.array_equal(st.action, ap["a"]) and np.array_equal(st.reward, ap["r"]) and np.array_equal(st.done, ap["d"])
    assert st.next_idx == int(ap["next_idx"]) and st.filled == int(ap["filled"])
    assert np.array_equal(buf.it_sum.value, ap["sum_tree"]) and np.array_equal(buf.it_min.value, ap["min_tree"])
    u = gu.sub(fx, "uniform")
    (s, a, r, s1, d), idx = buf.sample_uniform(u["idx"])
    for got, want in ((s["obs"], u["s"]), (a, u["a"]), (r, u["r"]), (s1["obs"], u["s1"]), (d, u["d"]), (idx, u["ridx"])):
        assert np.array_equal(got, want)
    for t in range(4):
        p = gu.sub(fx, f"per{t}")
        (s, a, r, s1, d), w, idxes = buf.sample(p["u"])
        assert np.array_equal(idxes, p["idxes"])
        assert np.array_equal(w, p["weights"])  # float64, same operation order: bit-exact
        assert np.array_equal(s["obs"], p["s"]) and np.array_equal(a, p["a"])
        buf.update_priorities(idxes, p["new_priorities"])
        assert np.array_equal(buf.it_sum.value, p["sum_tree"]) and np.array_equal(buf.it_min.value, p["min_tree"])
        assert buf.max_priority == float(p["max_priority"])


def test_aug_oracle_matches_reference():
    fx = gu.load("aug_pixels")
    b = gu.sub(fx, "buffer")
    # DrQ v1 without noise: exact integer crop in the reference -> bit-exact
    g = gu.sub(fx, "drqv1")
    idx = g["idx"]
    assert np.array_equal(ao.drq_v1_crop(b["s"][idx], g["w1"], g["h1"]), g["o"])
    assert np.array_equal(ao.drq_v1_crop(b["s1"][idx], g["w1"], g["h1"]), g["o1"])
    # DrQ v1 with injected N(0,1) noise, aug_mix = 0.5
    g = gu.sub(fx, "drqv1_noise")
    for key, src, noise in (("o", b["s"], g["n0"]), ("o1", b["s1"], g["n1"])):
        aug = ao.drq_v1_crop(src[idx], g["w1"], g["h1"], noise=noise)
        got = ao.mix(src[idx].astype(np.float32), aug, 0.5)
        assert np.array_equal(got, g[key])
    # DrQv2: the reference's bilinear grid_sample sits <= 4e-3 (0..255 scale) off the integer crop it encodes
    g = gu.sub(fx, "drqv2")
    aug = ao.drq_v2_crop(b["s"][idx], g["shift"])
    assert np.abs(aug - g["ao"]).max() <= 4e-3
    got = ao.mix(b["s"][idx].astype(np.float32), aug, 0.75)
    assert np.abs(got - g["o"]).max() <= 4e-3
    k = int(len(idx) * 0.75)
    assert np.array_equal(got[k:], g["o"][k:])  # un-augmented rows are a pure uint8 -> float cast
    assert np.array_equal(b["s"][idx].astype(np.float32), g["oo"])
    aug1 = ao.drq_v2_crop(b["s1"][idx], g["shift"])
    assert np.abs(aug1 - g["ao1"]).max() <= 4e-3


def test_rad_oracle_matches_reference():
    """RadAug (cv2 bilinear upscale + integer crop, augmentations.py:129-162) through sample_move_and_augment:
    bit-exact on the 9-channel frame stack, <= 1e-3 (0..255 scale) on the 3-channel image (cv2's <=4-channel path)."""
    fx = gu.load("aug_rad")
    for tag in ("stack9", "rgb3"):
        g = gu.sub(fx, tag)
        idx, crop, mixv = g["idx"], int(g["crop"]), float(g["mix"])
        for key, src in (("o", g["s"]), ("o1", g["s1"])):
            got = ao.mix(src[idx].astype(np.float32), ao.rad_crop(src[idx], g["h"], g["w"], crop), mixv)
            if src.shape[1] > 4:
                assert np.array_equal(got, g[key]), f"{tag}/{key}"
            else:
                assert np.abs(got - g[key]).max() <= 1e-3, f"{tag}/{key}"
            k = int(len(idx) * mixv)
            assert np.array_equal(got[k:], g[key][k:])   # un-augmented rows: a pure uint8 -> float cast


def _afbc_setup(fx):
    cfg = gu.cfg_of(fx)
    E, N, S, A, H = cfg["E"], cfg["N"], cfg["S"], cfg["A"], cfg["H"]
    agent = uo.OracleAgent(E, N, S, A, H, log_std_low=-5.0, log_std_high=2.0)
    agent.actors = uo.MLPStack.from_arrays(gu.sub(fx, "init/actors"))
    agent.critics = uo.MLPStack.from_arrays(gu.sub(fx, "init/critics"))
    agent.popart = gu.popart_from(fx, "init/popart", E)
    b = gu.sub(fx, "buffer")
    buf = ro.ReplayOracle(64, alpha=0.6, beta=1.0)
    buf.push({"obs": b["s"]}, b["a"], b["r"][:, None], {"obs": b["s1"]}, b["d"][:, None], priorities=b["priorities"])
    return cfg, agent, buf


def test_afbc_oracle_matches_reference():
    """offline_actor_update (advantage-filtered BC) + PER sampling + priority refresh."""
    torch.set_num_threads(1)
    fx = gu.load("afbc")
    cfg, agent, buf = _afbc_setup(fx)
    E, B = cfg["E"], cfg["B"]
    opt = uo.Adam(agent.actors.tensors(), lr=3e-4)
    t32 = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float32))
    for step in range(2):
        r = gu.sub(fx, f"step{step}/rand")
        batches, rands, last_idx = [], [], None
        for i in range(E):
            (s, a, rew, s1, d), w, idxes = buf.sample(r["u"][i])
            batches.append(({"obs": t32(s["obs"])}, t32(a), t32(rew), {"obs": t32(s1["obs"])}, t32(d)))
            rands.append(dict(adv_eps=[t32(e) for e in r["adv_eps"][i]]))
            last_idx = idxes
        logs, aux = uo.offline_actor_update(agent, batches, rands, dict(actor_clip=40.0, filter=True), opt)
        want = gu.sub(fx, f"step{step}/actor_grads")
        for n in uo.PARAM_NAMES:
            gu.assert_close(getattr(aux["grads"], n).numpy(), want[n], 1e-4, 1e-7, f"step{step} actor grad {n}")
        want = gu.sub(fx, f"step{step}/actors")
        for n in uo.PARAM_NAMES:
            gu.assert_close(getattr(agent.actors, n).numpy(), want[n], 1e-5, 3e-4 * 2e-2, f"step{step} actors {n}")
        _cmp_logs(logs, gu.sub(fx, f"step{step}/logs"), f"afbc step{step}")
        o, a = batches[-1][0], batches[-1][1]
        adv = uo.advantage(agent, int(r["prio_member"]), o, a, [t32(e) for e in r["prio_eps"]])
        newp = (torch.relu(adv) + 1e-4).squeeze(1).numpy()
        buf.update_priorities(last_idx, newp)
        tr = gu.sub(fx, f"step{step}/trees")
        gu.assert_close(buf.it_sum.value, tr["sum_tree"], 1e-5, 1e-9, "sum tree")  # priorities come from fp32 advantages
        assert np.array_equal(np.isinf(buf.it_min.value), np.isinf(tr["min_tree"]))
        assert abs(buf.max_priority - float(tr["max_priority"])) <= 1e-5 * float(tr["max_priority"])


@pytest.mark.parametrize("tag", ["rgb20", "stack24"])
def test_encoder_oracle_matches_reference(tag):
    """oracle/encoder_oracle.py (explicit tap-by-tap forward / backward) == the unmodified BigPixelEncoder + autograd."""
    from oracle import encoder_oracle as eo

    torch.set_num_threads(1)
    fx = gu.load("encoder")
    params = gu.sub(fx, f"{tag}/params")
    want = gu.sub(fx, f"{tag}/grads")
    out, cache = eo.forward(params, fx[f"{tag}/obs"].astype(np.float32))
    gu.assert_close(out.numpy(), fx[f"{tag}/out"], 2e-5, 2e-6, f"{tag} out")
    g = eo.backward(cache, fx[f"{tag}/dout"])
    for n in eo.PARAM_NAMES:
        gu.assert_close(g[n].numpy(), want[n], 1e-4, 1e-5, f"{tag} grad {n}")
