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
