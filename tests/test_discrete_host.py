"""Host logic of the SAC-Discrete updates (super_sac_b200/discrete.py) on CPU: the product's Python runs UNCHANGED on CPU
tensors against a numpy emulation of the C-ABI entry points it calls (tests/emulated_abi.py reads the same raw pointers),
and the results are compared with the golden vectors of the unmodified reference.  What this pins: argument order,
strides, the loss / gradient normalisation constants, PopArt and log plumbing, Adam wiring.  What it does NOT pin: the
CUDA kernels (tests/test_discrete_parity.py, -m gpu).  The product itself still refuses CPU tensors (test_host_logic).
"""
import numpy as np
import pytest
import torch

import golden_util as gu
from emulated_abi import EmulatedLib

RTOL = 1e-4


class IdentityEncoder(torch.nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.have_at_least_one_param = torch.nn.Linear(1, 1)
        self._dim = dim

    @property
    def embedding_dim(self):
        return self._dim

    def forward(self, obs_dict):
        return obs_dict["obs"]


def _load_stack(arena, arrs):
    with torch.no_grad():
        for n in ("W1", "b1", "W2", "b2", "W3", "b3"):
            arena.p[n].copy_(torch.as_tensor(arrs[n]))


def _cmp_stack(views, want, what, rtol=RTOL, atol=1e-6):
    for n in ("W1", "b1", "W2", "b2", "W3", "b3"):
        gu.assert_close(views[n].numpy(), want[n], rtol, atol, f"{what}.{n}")


def _cmp_logs(logs, want, what):
    for k, v in want.items():
        k2 = k.replace("|", "/")
        if k2.startswith("gradients/"):
            assert k2 in logs
            continue
        assert k2 in logs, f"{what}: missing log key {k2}"
        gu.assert_close(float(logs[k2]), float(v), 2e-4, 2e-5, f"{what} log {k2}")


@pytest.fixture
def emulated(monkeypatch):
    from super_sac_b200 import _lib, _logs, _ops

    emu = EmulatedLib()
    monkeypatch.setattr(_lib, "_lib", emu)
    monkeypatch.setattr(_lib, "stream_ptr", lambda: 0)
    monkeypatch.setattr(_ops, "check_cuda", lambda *t: None)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)

    def fetch(self, keep=False, synced=False):   # DeviceLogs.fetch without the pinned device->host copy
        host = self._buf[: self._n].tolist()
        for key, slot, transform in self._pending:
            if isinstance(transform, _logs._Multi):
                self[key] = transform.fn(*[host[s_] for s_ in slot])
                continue
            val = sum(host[s_] for s_ in slot) if isinstance(slot, (list, tuple)) else host[slot]
            self[key] = transform(val) if transform is not None else val
        self._pending = []
        return self

    monkeypatch.setattr(_logs.DeviceLogs, "fetch", fetch)
    return emu


@pytest.mark.parametrize("case", gu.DISCRETE_CASES + [gu.DISCRETE_ENCODER_CASE])
def test_discrete_updates_host_logic(case, emulated, monkeypatch):
    import copy
    import math
    from itertools import chain

    import super_sac_b200 as ssb
    from super_sac_b200 import _rng, augmentations, learning, learning_utils as lu, nets

    fx = gu.load("update_" + case)
    cfg = gu.cfg_of(fx)
    E, N, M, S, A, H, B = cfg["E"], cfg["N"], cfg["M"], cfg["S"], cfg["A"], cfg["H"], cfg["B"]
    shared = cfg.get("encoder") == "shared"   # a trainable (user, PyTorch) encoder in front: critic gradients train it
    agent = ssb.Agent(act_space_size=A, encoder=gu.encoder_from(fx, "init/encoder", S) if shared else IdentityEncoder(S),
                      actor_network_cls=nets.mlps.DiscreteActor,
                      critic_network_cls=nets.mlps.DiscreteCritic, discrete=True, ensemble_size=E, num_critics=N,
                      hidden_size=H, auto_rescale_targets=cfg.get("popart", False))
    assert agent._critic_arena.O == A and agent._critic_arena.D == S and agent._actor_arena.O == A
    _load_stack(agent._actor_arena, gu.sub(fx, "init/actors"))
    _load_stack(agent._critic_arena, gu.sub(fx, "init/critics"))
    pst = gu.sub(fx, "init/popart")
    for i, p in enumerate(agent.popart):
        if p:
            p.mu, p.nu, p.w, p.b = pst[f"{i}/mu"], pst[f"{i}/nu"], pst[f"{i}/w"], pst[f"{i}/b"]
            p._t = int(pst[f"{i}/t"])
    target = copy.deepcopy(agent)
    assert target.discrete and target._critic_arena.O == A
    _load_stack(target._critic_arena, gu.sub(fx, "init/target_critics"))
    if shared:
        target.encoder.load_state_dict({k: torch.as_tensor(v) for k, v in gu.sub(fx, "init/target_encoder").items()})
    # the module views see the arena (state_dict keys as the reference's: fc1 / fc2 / act_p, fc1 / fc2 / out)
    assert set(agent.actors[0].state_dict()) == {"fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "act_p.weight", "act_p.bias"}
    assert agent.actors[0].act_p.weight.data_ptr() == agent._actor_arena.p["W3"][0].data_ptr()

    critic_opt = torch.optim.Adam(chain(*(c.parameters() for c in agent.critics)), lr=3e-4, betas=(0.9, 0.999))
    actor_opt = torch.optim.Adam(chain(*(a.parameters() for a in agent.actors)), lr=3e-4, betas=(0.9, 0.999))
    enc_opt = torch.optim.Adam(agent.encoder.parameters(), lr=1e-4)
    log_alphas, alpha_opts = [], []
    for _ in range(E):
        la = torch.Tensor([math.log(cfg.get("init_alpha", 0.1))])
        la.requires_grad = True
        log_alphas.append(la)
        alpha_opts.append(torch.optim.Adam([la], lr=1e-4, betas=(0.5, 0.999)))
    augmenter = augmentations.AugmentationSequence([augmentations.IdentityAug(B)])

    # the replay gather is a kernel of the continuous path (checked on the GPU): hand the golden batch over in the packed
    # [s | a] layout sample_move_and_augment produces for state observations (learning_utils.py, product)
    bufd = gu.sub(fx, "buffer")
    queue = []

    def fake_sample(buffer, batch_size, augmenter, aug_mix, per=True, _idx=None):
        idx = torch.empty(batch_size, dtype=torch.int64)
        _rng.source().indices(idx, len(bufd["a"]))
        ix = idx.numpy()
        XA = torch.as_tensor(np.concatenate([bufd["s"][ix], bufd["a"][ix]], 1).astype(np.float32))
        X1 = torch.as_tensor(np.concatenate([bufd["s1"][ix], np.zeros((batch_size, 1), np.float32)], 1))
        rd = lu.ReplayDict()
        rd["primary_batch"] = ({"obs": XA[:, :S]}, XA[:, S:], torch.as_tensor(bufd["r"][ix]).reshape(-1, 1),
                               {"obs": X1[:, :S]}, torch.as_tensor(bufd["d"][ix]).reshape(-1, 1))
        rd["priority_idxs"], rd["imp_weights"] = idx, torch.ones(1)
        queue.append(rd)
        return rd

    monkeypatch.setattr(lu, "sample_move_and_augment", fake_sample)
    rec = {}
    o_td, o_bw = lu.compute_td_targets, lu.compute_backup_weights
    monkeypatch.setattr(lu, "compute_td_targets", lambda *a, **k: rec.setdefault("td", []).append(o_td(*a, **k)) or rec["td"][-1])
    monkeypatch.setattr(lu, "compute_backup_weights", lambda *a, **k: rec.setdefault("w", []).append(o_bw(*a, **k)) or rec["w"][-1])
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
                buffer=None, agent=agent, target_agent=target, critic_optimizer=critic_opt, encoder_optimizer=enc_opt,
                log_alphas=log_alphas, batch_size=B, gamma=cfg.get("gamma", 0.99), critic_clip=cfg.get("critic_clip"),
                encoder_clip=cfg.get("encoder_clip"), target_critic_ensemble_n=M, weighted_bellman_temp=cfg.get("weight_temp"),
                weight_type=cfg.get("weight_type"), pop=cfg.get("pop", False), augmenter=augmenter, encoder_lambda=0.0,
                aug_mix=0.0, discrete=True, random_process=None, noise_clip=None, per=False, update_priorities=False,
                dr3_coeff=cfg.get("dr3_coeff", 0.0))
            assert src.empty(), "not every scripted draw was consumed"
            for i in range(E):
                gu.assert_close(rec["td"][i][0].numpy(), fx[f"step{t}/td_target/{i}"], RTOL, 1e-5, f"step{t} td_target[{i}]")
                w = rec["w"][i]
                w = w.numpy() if torch.is_tensor(w) else np.array(w, dtype=np.float32)
                gu.assert_close(w, fx[f"step{t}/weights/{i}"], RTOL, 1e-5, f"step{t} weights[{i}]")
            _cmp_stack(agent._critic_arena.g, gu.sub(fx, f"step{t}/critic_grads"), f"step{t} critic_grads", atol=2e-7)
            _cmp_logs(logs, gu.sub(fx, f"step{t}/logs"), f"step{t}")
            with torch.no_grad():   # Polyak is a kernel of the continuous path: applied here in torch (learning_utils.py:160-162)
                tau = cfg.get("tau", 0.005)
                tf, sf = target._critic_arena.flat, agent._critic_arena.flat
                tf.copy_(tf * (1.0 - tau) + sf * tau)
            _cmp_stack(agent._critic_arena.p, gu.sub(fx, f"step{t}/critics"), f"step{t} critics", atol=3e-4 * 0.05)
            if shared:
                with torch.no_grad():
                    for tp, sp in zip(target.encoder.parameters(), agent.encoder.parameters()):
                        tp.copy_(tp * (1.0 - 0.01) + sp * 0.01)
                for which, enc in (("encoder", agent.encoder), ("target_encoder", target.encoder)):
                    want = gu.sub(fx, f"step{t}/{which}")
                    for k, v in enc.state_dict().items():
                        gu.assert_close(v.numpy(), want[k], RTOL, 1e-4 * 0.05, f"step{t} {which}.{k}")
            want_pop = gu.sub(fx, f"step{t}/popart")
            for i, p in enumerate(agent.popart):
                if p:
                    for n in ("mu", "nu", "w", "b"):
                        gu.assert_close(getattr(p, n).numpy(), want_pop[f"{i}/{n}"], RTOL, 1e-6, f"step{t} popart[{i}].{n}")
        alogs = learning.online_actor_update(
            buffer=None, agent=agent, pop=cfg.get("pop", False), actor_optimizer=actor_opt, log_alphas=log_alphas,
            batch_size=B, clip=cfg.get("actor_clip"), random_process=None, noise_clip=None, augmenter=augmenter, aug_mix=0.0,
            premade_replay_dicts=replay_dicts, per=False, discrete=True, use_baseline=False)
        _cmp_stack(agent._actor_arena.g, gu.sub(fx, "actor/grads"), "actor grads", atol=2e-7)
        _cmp_stack(agent._actor_arena.p, gu.sub(fx, "actor/actors"), "actors", atol=3e-4 * 0.05)
        _cmp_logs(alogs, gu.sub(fx, "actor/logs"), "actor")
        llogs = learning.alpha_update(
            buffer=None, agent=agent, optimizers=alpha_opts, batch_size=B, log_alphas=log_alphas, augmenter=augmenter,
            aug_mix=0.0, target_entropy=float(fx["alpha/target_entropy"]), premade_replay_dicts=replay_dicts, discrete=True)
        for i, la in enumerate(log_alphas):
            gu.assert_close(la.detach().numpy(), fx[f"alpha/log_alphas/{i}"], 1e-6, 1e-7, f"log_alpha[{i}]")
        _cmp_logs(llogs, gu.sub(fx, "alpha/logs"), "alpha")
        for name in ("discrete_value", "discrete_critic_loss_seed", "discrete_actor_seed", "discrete_neg_entropy"):
            assert name in emulated.calls
    finally:
        _rng.set_source(old_src)


def test_discrete_offline_update_host_logic(emulated, monkeypatch):
    """offline_actor_update(discrete=True) with the indirect advantage filter, the AdvantageEstimator call surface and
    adjust_priorities, on CPU through the emulated C ABI, against the golden of the unmodified reference."""
    import random
    from itertools import chain

    import super_sac_b200 as ssb
    from super_sac_b200 import _rng, augmentations, learning, learning_utils as lu, nets

    fx = gu.load("discrete_afbc")
    cfg = gu.cfg_of(fx)
    E, N, S, A, H, B = cfg["E"], cfg["N"], cfg["S"], cfg["A"], cfg["H"], cfg["B"]
    agent = ssb.Agent(act_space_size=A, encoder=IdentityEncoder(S), actor_network_cls=nets.mlps.DiscreteActor,
                      critic_network_cls=nets.mlps.DiscreteCritic, discrete=True, ensemble_size=E, num_critics=N,
                      hidden_size=H, auto_rescale_targets=True)
    _load_stack(agent._actor_arena, gu.sub(fx, "init/actors"))
    _load_stack(agent._critic_arena, gu.sub(fx, "init/critics"))
    pst = gu.sub(fx, "init/popart")
    for i, p in enumerate(agent.popart):
        p.mu, p.nu, p.w, p.b = pst[f"{i}/mu"], pst[f"{i}/nu"], pst[f"{i}/w"], pst[f"{i}/b"]
        p._t = int(pst[f"{i}/t"])
    bufd, idx = gu.sub(fx, "buffer"), fx["rand/idx"]

    def make_rd(ix):
        XA = torch.as_tensor(np.concatenate([bufd["s"][ix], bufd["a"][ix]], 1).astype(np.float32))
        rd = lu.ReplayDict()
        rd["primary_batch"] = ({"obs": XA[:, :S]}, XA[:, S:], None, None, None)
        rd["priority_idxs"], rd["imp_weights"] = torch.as_tensor(ix), torch.ones(1)
        return rd

    for i in range(E):   # agent.adv_estimator(o, a, i) -> [B, 1] (adv_estimator.py:81-89 call surface)
        rd = make_rd(idx[i])
        adv = agent.adv_estimator(rd["primary_batch"][0], rd["primary_batch"][1], i)
        gu.assert_close(adv.numpy(), fx[f"adv/{i}"], RTOL, 1e-6, f"adv[{i}]")

    def fake_sample(buffer, batch_size, augmenter, aug_mix, per=True, _idx=None):
        ix = torch.empty(batch_size, dtype=torch.int64)
        _rng.source().indices(ix, len(bufd["a"]))
        return make_rd(ix.numpy())

    monkeypatch.setattr(lu, "sample_move_and_augment", fake_sample)
    actor_opt = torch.optim.Adam(chain(*(a.parameters() for a in agent.actors)), lr=3e-4, betas=(0.9, 0.999))
    enc_opt = torch.optim.Adam(agent.encoder.parameters(), lr=1e-4)
    augmenter = augmentations.AugmentationSequence([augmentations.IdentityAug(B)])
    src = _rng.ScriptedSource()
    old_src = _rng.set_source(src)
    try:
        for i in range(E):
            src.push("indices", idx[i])
        logs = learning.offline_actor_update(
            buffer=None, agent=agent, actor_optimizer=actor_opt, encoder_optimizer=enc_opt, batch_size=B,
            actor_clip=cfg["actor_clip"], update_encoder=False, encoder_clip=None, augmenter=augmenter, actor_lambda=0.0,
            aug_mix=0.0, premade_replay_dicts=None, per=False, discrete=True, filter_=True)
        assert src.empty()
    finally:
        _rng.set_source(old_src)
    _cmp_stack(agent._actor_arena.g, gu.sub(fx, "actor/grads"), "actor grads", atol=2e-7)
    _cmp_stack(agent._actor_arena.p, gu.sub(fx, "actor/actors"), "actors", atol=3e-4 * 0.05)
    _cmp_logs(logs, gu.sub(fx, "actor/logs"), "offline actor")

    class Buf:
        def update_priorities(self, idxs, prios):
            self.idxs, self.prios = idxs, prios

    buf = Buf()
    random.seed(7)
    lu.adjust_priorities({}, make_rd(idx[-1]), agent, buf)
    assert np.array_equal(np.asarray(buf.idxs), fx["priorities/idxs"])
    gu.assert_close(buf.prios.numpy(), fx["priorities/values"], 1e-5, 1e-6, "priorities")
    assert buf.prios.dtype == torch.float64
    assert "discrete_advantage" in emulated.calls and "discrete_bc_seed" in emulated.calls


def test_discrete_acting_path_shapes(emulated):
    """Agent.forward / sample_action of a discrete agent (agent.py:204-221, :262-320) return action indices."""
    import super_sac_b200 as ssb
    from super_sac_b200 import nets

    agent = ssb.Agent(act_space_size=4, encoder=IdentityEncoder(3), actor_network_cls=nets.mlps.DiscreteActor,
                      critic_network_cls=nets.mlps.DiscreteCritic, discrete=True, ensemble_size=2, num_critics=2,
                      hidden_size=16, ucb_bonus=0.5)
    obs = {"obs": np.zeros((5, 3), np.float32)}
    s = torch.zeros(5, 3)
    with torch.no_grad():
        greedy = agent._discrete_forward(s)
        act, dist = agent._discrete_sample(s, 5)
    assert greedy.shape == (5, 1) and act.shape == (5, 1) and dist.probs.shape == (5, 4)
    assert 0 <= int(act.min()) and int(act.max()) < 4
    agent.ucb_bonus = 0.0
    with torch.no_grad():
        act, dist = agent._discrete_sample(s, 5)
    assert act.shape == (5, 1)
    # the public call surface with numpy observations (agent.py:204-221, :262-327): indices come back as numpy arrays
    g = agent.forward(obs, num_envs=5)
    assert g.shape == (5, 1) and np.array_equal(g, greedy.numpy())
    assert np.array_equal(agent.discrete_forward(obs, num_envs=5), g)
    one = agent.forward({"obs": np.zeros(3, np.float32)})          # num_envs = 1: the env axis is squeezed away
    assert one.shape == (1,)
    a, dist = agent.sample_action(obs, num_envs=5, return_dist=True)
    assert a.shape == (5, 1) and a.dtype.kind == "i" and dist.probs.shape == (5, 4)
    assert agent.sample_action({"obs": np.zeros(3, np.float32)}).shape == (1,)


def test_discrete_unsupported_options_fail_loudly(emulated):
    """What the discrete path does not implement raises instead of silently computing something else."""
    import super_sac_b200 as ssb
    from super_sac_b200 import discrete, learning, nets

    cont = ssb.Agent(2, IdentityEncoder(3), nets.mlps.ContinuousStochasticActor, nets.mlps.ContinuousCritic,
                     ensemble_size=1, num_critics=2, hidden_size=16)
    with pytest.raises(ValueError, match="discrete=True"):
        learning.alpha_update(buffer=None, agent=cont, optimizers=[], batch_size=4, log_alphas=[], augmenter=None,
                              aug_mix=0.0, target_entropy=0.0, premade_replay_dicts=None, discrete=True)
    agent = ssb.Agent(act_space_size=4, encoder=IdentityEncoder(3), actor_network_cls=nets.mlps.DiscreteActor,
                      critic_network_cls=nets.mlps.DiscreteCritic, discrete=True, ensemble_size=2, num_critics=2,
                      hidden_size=16)
    with pytest.raises(NotImplementedError, match="sunrise"):
        discrete.compute_backup_weights({}, {"primary_batch": (None,) * 5}, agent, agent, "softmax", 10.0, 4)
    with pytest.raises(NotImplementedError, match="invariance"):
        discrete.critic_update(None, agent, agent, None, None, [], 4, 0.99, None, None, 2, None, None, False, None, 0.5, 0.0,
                               False, False, 0.0)
    with pytest.raises(NotImplementedError, match="invariance"):
        discrete.offline_actor_update(None, agent, None, None, 4, None, False, None, None, 0.5, 0.0, None, False, True)
    # a discrete Agent needs S -> H -> H -> A networks (the continuous critic takes [s | a])
    with pytest.raises(NotImplementedError):
        ssb.Agent(4, IdentityEncoder(3), nets.mlps.DiscreteActor, nets.mlps.ContinuousCritic, discrete=True)
