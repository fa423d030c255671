"""Drop-in proof: the reference's OWN training loop -- the unmodified ``super_sac.main.super_sac``
of baseline/_ref (main.py:285-546) -- driven for a few hundred steps on a stub environment with ``Agent``,
``ReplayBuffer``, ``learning``, ``learning_utils`` and ``augmentations`` swapped for this package and
``enable_auto_graphs()`` on.  Checks: it runs, every logged scalar is finite, and the update functions returned exactly
the log keys the reference's own functions return on the same configuration (run on the CPU with the reference's
classes)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from baseline import ref_import  # noqa: E402

pytestmark = pytest.mark.gpu


class _Box:
    def __init__(self, dim):
        self.low, self.high, self.shape = -np.ones(dim, np.float32), np.ones(dim, np.float32), (dim,)
        self._rng = np.random.default_rng(0)

    def sample(self):
        return self._rng.uniform(-1, 1, self.shape).astype(np.float32)


class _StubEnv:
    """A 17-d / 6-d linear system with a quadratic reward; dict observations like the reference's wrappers produce."""

    def __init__(self, seed, horizon=50):
        self.rng = np.random.default_rng(seed)
        self.action_space = _Box(6)
        self.horizon = horizon
        self.A = self.rng.standard_normal((17, 17)).astype(np.float32) * 0.1
        self.Bm = self.rng.standard_normal((17, 6)).astype(np.float32) * 0.3

    def reset(self):
        self.t = 0
        self.x = self.rng.standard_normal(17).astype(np.float32)
        return {"obs": self.x.copy()}, {}

    def step(self, a):
        self.t += 1
        self.x = np.tanh(self.A @ self.x + self.Bm @ np.asarray(a, np.float32) + 0.05 * self.rng.standard_normal(17)).astype(np.float32)
        rew = float(-np.square(self.x).mean())
        return {"obs": self.x.copy()}, rew, False, self.t >= self.horizon, {}


def _run(pkg, main, device, steps, graphs):
    import copy  # noqa: F401

    class IdentityEncoder(pkg.nets.Encoder):
        def __init__(self):
            super().__init__()

        @property
        def embedding_dim(self):
            return 17

        def forward(self, obs):
            return obs["obs"]

    torch.manual_seed(0)
    agent = pkg.Agent(act_space_size=6, encoder=IdentityEncoder(), actor_network_cls=pkg.nets.mlps.ContinuousStochasticActor,
                      critic_network_cls=pkg.nets.mlps.ContinuousCritic, ensemble_size=1, num_critics=2, hidden_size=64,
                      auto_rescale_targets=True, log_std_low=-5.0, log_std_high=2.0)
    ours = pkg.__name__ == "super_sac_b200"
    buffer = pkg.replay.ReplayBuffer(10_000, **(dict(device=device) if ours else {}))
    env = _StubEnv(1)
    pkg.learning_utils.warmup_buffer(buffer, env, 300, 50, 1, 0.99)
    seen = {}
    L = main.learning
    wrapped = {}
    for fn in ("critic_update", "online_actor_update", "alpha_update"):
        orig = getattr(L, fn)

        def make(orig=orig, fn=fn):
            def f(*a, **k):
                out = orig(*a, **k)
                logs = out[0] if isinstance(out, tuple) else out
                seen.setdefault(fn, set()).update(logs.keys())
                for key, v in logs.items():
                    assert np.isfinite(float(v)), f"{fn}: {key} = {v}"
                return out
            return f

        wrapped[fn] = orig
        setattr(L, fn, make())
    try:
        main.super_sac(agent, buffer, env, _StubEnv(2), num_steps_offline=0, num_steps_online=steps, batch_size=64,
                       critic_updates_per_step=2, use_afbc_update_online=False, use_pg_update_online=True, pop=True,
                       weight_type=None, eval_interval=10**9, evaluation_method=lambda *a, **k: {"eval/mean_return": 0.0},
                       log_to_disk=False, save_to_disk=False, verbosity=0,
                       max_episode_steps=50, target_delay=2)
    finally:
        for fn, orig in wrapped.items():
            setattr(L, fn, orig)
    return agent, buffer, seen


def test_reference_training_loop_runs_on_the_drop_in():
    if not ref_import.available():
        pytest.skip("baseline/_ref (the unmodified reference) did not travel")
    import super_sac_b200 as ssb
    from super_sac_b200 import graphed

    ref = ref_import.import_reference(device="cpu")
    main = ref.main
    # (1) the reference with its own classes, on the CPU: the log keys to expect
    _, _, want = _run(ref, main, "cpu", steps=12, graphs=False)
    # (2) the same loop with this package swapped in
    saved = (main.learning, main.lu, main.augmentations, main.device)
    main.learning, main.lu, main.augmentations, main.device = ssb.learning, ssb.learning_utils, ssb.augmentations, torch.device("cuda")
    graphed.enable_auto_graphs(True)
    try:
        agent, buffer, got = _run(ssb, main, torch.device("cuda"), steps=200, graphs=True)
    finally:
        graphed.enable_auto_graphs(False)
        main.learning, main.lu, main.augmentations, main.device = saved
    for fn in want:
        assert got[fn] == want[fn], f"{fn}: log keys differ: {sorted(got[fn] ^ want[fn])}"
    assert len(buffer) == 300 + 199   # env interaction starts at step 1 (main.py:327)
    assert buffer.total_sample_calls > 0
    for p in list(agent.critics[0].parameters()) + list(agent.actors[0].parameters()):
        assert torch.isfinite(p).all()
    # the caller-built torch.optim.Adam objects of main.py:188-227 were recognised: their state views moved
    assert int(agent._critic_arena.flat.isfinite().all())


class _StubPixelEnv:
    """uint8 frame stacks [3, 20, 20] under the key "obs" (what the reference's pixel wrappers hand out), 2-d actions."""

    def __init__(self, seed, horizon=25):
        self.rng = np.random.default_rng(seed)
        self.action_space = _Box(2)
        self.horizon = horizon

    def _frame(self):
        base = (np.add.outer(np.arange(20), np.arange(20)) * 3 + 40 * self.phase) % 256
        img = np.stack([(base + 37 * c) % 256 for c in range(3)]).astype(np.int64)
        return {"obs": ((img + self.rng.integers(0, 8, img.shape)) % 256).astype(np.uint8)}

    def reset(self):
        self.t, self.phase = 0, 0.0
        return self._frame(), {}

    def step(self, a):
        self.t += 1
        self.phase += float(np.asarray(a, np.float32).sum())
        return self._frame(), float(np.cos(self.phase)), False, self.t >= self.horizon, {}


def test_reference_training_loop_runs_on_the_drop_in_with_the_native_pixel_encoder():
    """The same unmodified loop in its DrQv2 configuration (experiments/dmc/drqv2.gin): pixel observations, Drqv2Aug,
    BigPixelEncoder (here the native tcgen05 encoder with its fused optimiser step), deterministic actor with the
    exploration process, 3-step returns, encoder tau 1.0.  Finite logs, the reference's log keys, a trained encoder."""
    if not ref_import.available():
        pytest.skip("baseline/_ref (the unmodified reference) did not travel")
    import super_sac_b200 as ssb

    ref = ref_import.import_reference(device="cpu")
    main = ref.main

    def run(pkg, device, steps):
        class PixelEncoder(pkg.nets.Encoder):   # experiments/dmc/train_dmc_from_pixels.py:15-27
            def __init__(self):
                super().__init__()
                self._enc = pkg.nets.cnns.BigPixelEncoder((3, 20, 20), 12)

            @property
            def embedding_dim(self):
                return 12

            def forward(self, obs_dict):
                return self._enc(obs_dict["obs"])

        torch.manual_seed(0)
        agent = pkg.Agent(act_space_size=2, encoder=PixelEncoder(), actor_network_cls=pkg.nets.mlps.ContinuousDeterministicActor,
                          critic_network_cls=pkg.nets.mlps.ContinuousCritic, ensemble_size=1, num_critics=2, hidden_size=64,
                          auto_rescale_targets=False)
        ours = pkg.__name__ == "super_sac_b200"
        buffer = pkg.replay.ReplayBuffer(2_000, **(dict(device=device) if ours else {}))
        env = _StubPixelEnv(1)
        pkg.learning_utils.warmup_buffer(buffer, env, 120, 25, 3, 0.99)
        enc0 = {k: v.detach().clone() for k, v in agent.encoder._enc.state_dict().items()}
        seen = {}
        L = main.learning
        wrapped = {}
        for fn in ("critic_update", "online_actor_update"):
            orig = getattr(L, fn)

            def make(orig=orig, fn=fn):
                def f(*a, **k):
                    out = orig(*a, **k)
                    logs = out[0] if isinstance(out, tuple) else out
                    seen.setdefault(fn, set()).update(logs.keys())
                    for key, v in logs.items():
                        assert np.isfinite(float(v)), f"{fn}: {key} = {v}"
                    return out
                return f

            wrapped[fn] = orig
            setattr(L, fn, make())
        try:
            main.super_sac(agent, buffer, env, _StubPixelEnv(2), num_steps_offline=0, num_steps_online=steps, batch_size=32,
                           critic_updates_per_step=1, use_afbc_update_online=False, use_pg_update_online=True, pop=False,
                           weight_type=None, init_alpha=0, alpha_lr=0, use_exploration_process=True,
                           exploration_param_init=1.0, exploration_param_final=0.1, exploration_param_anneal=500,
                           exploration_update_clip=0.3, n_step=3, gamma=0.99, mlp_tau=0.01, encoder_tau=1.0, target_delay=1,
                           actor_clip=None, critic_clip=None, encoder_clip=None,
                           augmenter=pkg.augmentations.AugmentationSequence([pkg.augmentations.Drqv2Aug(32)]), aug_mix=1.0,
                           eval_interval=10**9, evaluation_method=lambda *a, **k: {"eval/mean_return": 0.0},
                           log_to_disk=False, save_to_disk=False, verbosity=0, max_episode_steps=25)
        finally:
            for fn, orig in wrapped.items():
                setattr(L, fn, orig)
        return agent, enc0, seen

    _, _, want = run(ref, "cpu", steps=8)
    saved = (main.learning, main.lu, main.augmentations, main.device)
    main.learning, main.lu, main.augmentations, main.device = ssb.learning, ssb.learning_utils, ssb.augmentations, torch.device("cuda")
    from super_sac_b200 import graphed

    graphed.enable_auto_graphs(True)   # the native encoder and the device-side noise scale make this configuration capturable
    try:
        agent, enc0, got = run(ssb, torch.device("cuda"), steps=60)
        captured = sorted(k[0] for k, e in graphed._auto["cache"].items() if any(sl.graph is not None for sl in e.slots))
        assert "critic" in captured, captured
    finally:
        graphed.enable_auto_graphs(False)
        main.learning, main.lu, main.augmentations, main.device = saved
    for fn in want:
        assert got[fn] == want[fn], f"{fn}: log keys differ: {sorted(got[fn] ^ want[fn])}"
    net = agent.encoder._enc
    assert net.__dict__.get("_flat") is not None, "the native encoder path did not run"
    moved = 0.0
    for k, v in net.state_dict().items():
        assert torch.isfinite(v).all(), k
        moved = max(moved, float((v.cpu() - enc0[k].cpu()).abs().max()))
    assert moved > 0.0, "the encoder was never updated"


class _Discrete:
    def __init__(self, n):
        self.n, self.shape = n, ()
        self._rng = np.random.default_rng(0)

    def sample(self):
        return int(self._rng.integers(0, self.n))


class _StubDiscreteEnv:
    """An 8-d state driven by one of 4 discrete pushes; dict observations, gym's Discrete action space surface."""

    def __init__(self, seed, horizon=40):
        self.rng = np.random.default_rng(seed)
        self.action_space = _Discrete(4)
        self.horizon = horizon
        self.A = self.rng.standard_normal((8, 8)).astype(np.float32) * 0.2
        self.push = self.rng.standard_normal((4, 8)).astype(np.float32) * 0.5
        self.actions_seen = []

    def reset(self):
        self.t = 0
        self.x = self.rng.standard_normal(8).astype(np.float32)
        return {"obs": self.x.copy()}, {}

    def step(self, a):
        a = int(np.asarray(a).reshape(-1)[0])
        assert 0 <= a < 4, a
        self.actions_seen.append(a)
        self.t += 1
        self.x = np.tanh(self.A @ self.x + self.push[a] + 0.05 * self.rng.standard_normal(8)).astype(np.float32)
        return {"obs": self.x.copy()}, float(-np.square(self.x).mean()), False, self.t >= self.horizon, {}


def test_reference_training_loop_runs_on_the_drop_in_with_discrete_actions():
    """the same unmodified loop with a discrete agent (DiscreteActor / DiscreteCritic, SAC-Discrete updates,
    target entropy from action_space.n, main.py:240-241): finite logs, the reference's log keys, valid action indices."""
    if not ref_import.available():
        pytest.skip("baseline/_ref (the unmodified reference) did not travel")
    import super_sac_b200 as ssb

    ref = ref_import.import_reference(device="cpu")
    main = ref.main

    def run(pkg, device, steps):
        class IdentityEncoder(pkg.nets.Encoder):
            def __init__(self):
                super().__init__()

            @property
            def embedding_dim(self):
                return 8

            def forward(self, obs):
                return obs["obs"]

        torch.manual_seed(0)
        agent = pkg.Agent(act_space_size=4, encoder=IdentityEncoder(), actor_network_cls=pkg.nets.mlps.DiscreteActor,
                          critic_network_cls=pkg.nets.mlps.DiscreteCritic, discrete=True, ensemble_size=1, num_critics=2,
                          hidden_size=64, auto_rescale_targets=True)
        ours = pkg.__name__ == "super_sac_b200"
        buffer = pkg.replay.ReplayBuffer(5_000, **(dict(device=device) if ours else {}))
        env = _StubDiscreteEnv(1)
        pkg.learning_utils.warmup_buffer(buffer, env, 200, 40, 1, 0.99)
        seen = {}
        L = main.learning
        wrapped = {}
        for fn in ("critic_update", "online_actor_update", "alpha_update"):
            orig = getattr(L, fn)

            def make(orig=orig, fn=fn):
                def f(*a, **k):
                    assert k.get("discrete", False) is True, f"{fn} was not called with discrete=True"
                    out = orig(*a, **k)
                    logs = out[0] if isinstance(out, tuple) else out
                    seen.setdefault(fn, set()).update(logs.keys())
                    for key, v in logs.items():
                        assert np.isfinite(float(v)), f"{fn}: {key} = {v}"
                    return out
                return f

            wrapped[fn] = orig
            setattr(L, fn, make())
        try:
            main.super_sac(agent, buffer, env, _StubDiscreteEnv(2), num_steps_offline=0, num_steps_online=steps,
                           batch_size=64, critic_updates_per_step=1, use_afbc_update_online=False,
                           use_pg_update_online=True, pop=True, weight_type=None, eval_interval=10**9,
                           evaluation_method=lambda *a, **k: {"eval/mean_return": 0.0}, log_to_disk=False,
                           save_to_disk=False, verbosity=0, max_episode_steps=40, target_delay=2)
        finally:
            for fn, orig in wrapped.items():
                setattr(L, fn, orig)
        return agent, buffer, env, seen

    _, _, _, want = run(ref, "cpu", steps=10)
    saved = (main.learning, main.lu, main.augmentations, main.device)
    main.learning, main.lu, main.augmentations, main.device = ssb.learning, ssb.learning_utils, ssb.augmentations, torch.device("cuda")
    try:
        agent, buffer, env, got = run(ssb, torch.device("cuda"), steps=120)
    finally:
        main.learning, main.lu, main.augmentations, main.device = saved
    assert set(want) == {"critic_update", "online_actor_update", "alpha_update"}
    for fn in want:
        assert got[fn] == want[fn], f"{fn}: log keys differ: {sorted(got[fn] ^ want[fn])}"
    assert len(buffer) == 200 + 119
    assert len(set(env.actions_seen[200:])) > 1, "the policy never explored"
    for p in list(agent.critics[0].parameters()) + list(agent.actors[0].parameters()):
        assert torch.isfinite(p).all()
