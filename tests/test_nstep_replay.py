"""On-the-fly n-step transitions and the frame-deduplicated ring (super_sac_b200/nstep_replay.py) against the
reference's host-side construction: main.py:353-365 builds n-step transitions with a deque and pushes them into the classic
ring.  Fed with the same stream of one-step transitions, NStepReplayBuffer must hand out exactly those transitions (bit for
bit, including the accumulated return), and an update drawn from it must equal the update drawn from a classic buffer
that holds the reference's n-step tuples."""
from collections import deque

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def reference_nstep_tuples(steps, n_step, gamma):
    """The transitions the reference's training loop pushes (main.py:340-365): a deque of the last n one-step transitions,
    cleared at every reset; once it is full its oldest entry is popped and the discounted rewards of the rest are added in
    a Python loop.  ``steps``: (state, action, reward, next_state, terminated, episode_over)."""
    out = []
    dq = deque([], maxlen=n_step)
    done = True
    for state, action, reward, next_state, terminated, over in steps:
        if done:
            dq.clear()
        dq.append((state, action, reward, next_state, terminated))
        done = over
        if len(dq) == dq.maxlen:
            s, a, r, s1, d = dq.popleft()
            for i, trans in enumerate(dq):
                *_, r_i, s1, d = trans
                r += (gamma ** (i + 1)) * r_i
            out.append((s, a, r, s1, d))
    return out


def make_stream(kind, n_steps, rng, reward_f32, k=3):
    """Episodes of random length (some shorter than the n-step window), terminated or merely truncated at their end."""
    steps = []
    while len(steps) < n_steps:
        T = int(rng.integers(1, 14))
        if kind == "flat":
            obs = [rng.standard_normal(5).astype(np.float32) for _ in range(T + 1)]
            stack = lambda t: {"obs": obs[t]}   # noqa: E731
        else:
            frames = [rng.integers(0, 256, (2, 12, 12), dtype=np.uint8) for _ in range(T + 1)]
            prop = [rng.standard_normal(3).astype(np.float32) for _ in range(T + 1)]

            def stack(t):   # frame stacking wrapper: the first frame is repeated at reset
                return {"pixels": np.concatenate([frames[max(t - j, 0)] for j in range(k - 1, -1, -1)], 0), "proprio": prop[t]}
        terminated_at_end = bool(rng.uniform() < 0.5)
        for t in range(T):
            r = rng.standard_normal()
            r = np.float32(r) if reward_f32 else float(r)
            last = t == T - 1
            steps.append((stack(t), rng.uniform(-1, 1, 2).astype(np.float32), r, stack(t + 1), last and terminated_at_end, last))
    return steps


def as_np(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


@pytest.mark.parametrize("reward_f32", [False, True], ids=["float64-rewards", "float32-rewards"])
@pytest.mark.parametrize("kind,n_step,size,fcap", [("flat", 3, 4096, None), ("flat", 1, 4096, None), ("pixels", 3, 4096, None),
                                                  ("pixels", 5, 61, 70), ("flat", 3, 40, 44)])
def test_nstep_buffer_hands_out_the_reference_transitions(kind, n_step, size, fcap, reward_f32):
    import super_sac_b200 as ssb

    rng = np.random.default_rng(n_step * 100 + size)
    gamma = 0.97
    steps = make_stream(kind, 400, rng, reward_f32)
    want = reference_nstep_tuples(steps, n_step, gamma)
    buf = ssb.replay.NStepReplayBuffer(size, n_step=n_step, gamma=gamma, frame_stack=3 if kind == "pixels" else 1, frame_capacity=fcap,
                                       device=DEV)
    for s, a, r, s1, term, over in steps:
        buf.push(s, a, r, s1, term, terminate_traj=over)
    n = len(buf)
    assert 0 < n <= len(want)
    if size >= 4096:
        assert n == len(want)          # nothing evicted: the same set of transitions, in the same order
    want = want[len(want) - n:]        # a wrapped ring holds the most recent ones
    (s, a, r, s1, d) = buf.get_all_transitions()
    torch.cuda.synchronize()
    for key in s:
        assert np.array_equal(as_np(s[key]), np.stack([w[0][key] for w in want])), f"s[{key}]"
        assert np.array_equal(as_np(s1[key]), np.stack([w[3][key] for w in want])), f"s1[{key}]"
    assert np.array_equal(as_np(a), np.stack([w[1] for w in want]))
    assert np.array_equal(as_np(r).reshape(-1), np.asarray([np.float32(w[2]) for w in want])), "n-step return"
    assert np.array_equal(as_np(d).reshape(-1).astype(bool), np.asarray([bool(w[4]) for w in want]))
    if kind == "pixels":   # every frame once (+ k per episode start) instead of 2k copies per transition
        assert buf.bytes_per_transition() < 0.25 * (2 * 6 * 12 * 12)


def test_nstep_buffer_rejects_out_of_order_transitions():
    import super_sac_b200 as ssb

    buf = ssb.replay.NStepReplayBuffer(64, n_step=2, device=DEV)
    o = lambda v: {"obs": np.full(4, v, np.float32)}   # noqa: E731
    buf.push(o(0), np.zeros(1, np.float32), 0.0, o(1), False)
    with pytest.raises(ValueError):
        buf.push(o(5), np.zeros(1, np.float32), 0.0, o(6), False)


@pytest.mark.parametrize("kind", ["flat", "pixels"])
def test_update_batches_from_nstep_buffer_equal_classic_buffer(kind):
    """sample_move_and_augment (scripted positions and shifts) over the one-step ring == over a classic ReplayBuffer loaded
    with the reference's n-step tuples -- so every update entry point sees identical batches."""
    import super_sac_b200 as ssb
    from super_sac_b200 import _rng, augmentations, learning_utils as lu

    rng = np.random.default_rng(7)
    n_step, gamma, B = 3, 0.99, 64
    steps = make_stream(kind, 300, rng, reward_f32=False)
    want = reference_nstep_tuples(steps, n_step, gamma)
    nb = ssb.replay.NStepReplayBuffer(4096, n_step=n_step, gamma=gamma, frame_stack=3 if kind == "pixels" else 1, device=DEV)
    for s, a, r, s1, term, over in steps:
        nb.push(s, a, r, s1, term, terminate_traj=over)
    cb = ssb.replay.ReplayBuffer(4096, device=DEV)
    keys = list(want[0][0].keys())
    cb.load_experience({k: np.stack([w[0][k] for w in want]) for k in keys}, np.stack([w[1] for w in want]),
                       np.asarray([w[2] for w in want], dtype=np.float32), {k: np.stack([w[3][k] for w in want]) for k in keys},
                       np.asarray([w[4] for w in want]))
    assert len(nb) == len(cb) == len(want)
    idx = rng.integers(0, len(want), B)
    if kind == "pixels":
        augm = augmentations.AugmentationSequence([augmentations.Drqv2Aug(B)], keys=["pixels"])
        shift = rng.integers(0, 9, (B, 2)).astype(np.int32)
    else:
        augm = augmentations.AugmentationSequence([augmentations.IdentityAug(B)])
        shift = None
    outs = []
    for buf in (cb, nb):
        src = _rng.ScriptedSource()
        old = _rng.set_source(src)
        try:
            src.push("indices", idx)
            if shift is not None:
                src.push("shifts", shift)
            rd = lu.sample_move_and_augment(buffer=buf, batch_size=B, augmenter=augm, aug_mix=1.0 if kind == "pixels" else 0.0, per=False)
        finally:
            _rng.set_source(old)
        outs.append(rd["primary_batch"])
    torch.cuda.synchronize()
    (o0, a0, r0, p0, d0), (o1, a1, r1, p1, d1) = outs
    for key in keys:
        assert torch.equal(o0[key], o1[key]) and torch.equal(p0[key], p1[key]), key
    assert torch.equal(a0, a1) and torch.equal(r0, r1) and torch.equal(d0, d1)


def test_critic_update_runs_on_nstep_buffer_and_matches_classic():
    """The whole critic update (graph-replayed as well) on the one-step ring == on the classic buffer with the reference's
    n-step tuples: same Philox seed, same positions, identical parameters afterwards."""
    import copy
    from itertools import chain

    import cuda_util as cu
    import super_sac_b200 as ssb
    from super_sac_b200 import augmentations, graphed, learning, learning_utils as lu, nets

    rng = np.random.default_rng(3)
    n_step, gamma, B = 3, 0.99, 64
    steps = make_stream("flat", 600, rng, reward_f32=False)
    want = reference_nstep_tuples(steps, n_step, gamma)

    def run(which, auto):
        ssb.manual_seed(21)
        torch.manual_seed(21)
        agent = ssb.Agent(act_space_size=2, encoder=cu.IdentityEncoder(5), actor_network_cls=nets.mlps.ContinuousStochasticActor,
                          critic_network_cls=nets.mlps.ContinuousCritic, ensemble_size=1, num_critics=4, hidden_size=64,
                          auto_rescale_targets=False, log_std_low=-5.0, log_std_high=2.0)
        agent.to(DEV)
        target = copy.deepcopy(agent)
        if which == "nstep":
            buf = ssb.replay.NStepReplayBuffer(4096, n_step=n_step, gamma=gamma, device=DEV)
            for s, a, r, s1, term, over in steps:
                buf.push(s, a, r, s1, term, terminate_traj=over)
        else:
            buf = ssb.replay.ReplayBuffer(4096, device=DEV)
            buf.load_experience({"obs": np.stack([w[0]["obs"] for w in want])}, np.stack([w[1] for w in want]),
                                np.asarray([w[2] for w in want], dtype=np.float32), {"obs": np.stack([w[3]["obs"] for w in want])},
                                np.asarray([w[4] for w in want]))
        c_opt = torch.optim.Adam(chain(*(c.parameters() for c in agent.critics)), lr=3e-4)
        e_opt = torch.optim.Adam(agent.encoder.parameters(), lr=1e-4)
        la = [torch.tensor([-2.3], device=DEV, requires_grad=True)]
        aug = augmentations.AugmentationSequence([augmentations.IdentityAug(B)])
        graphed.enable_auto_graphs(auto)
        try:
            for k in range(6):
                logs, _ = learning.critic_update(
                    buffer=buf, agent=agent, target_agent=target, critic_optimizer=c_opt, encoder_optimizer=e_opt, log_alphas=la,
                    batch_size=B, gamma=gamma ** n_step, critic_clip=None, encoder_clip=None, target_critic_ensemble_n=2,
                    weighted_bellman_temp=None, weight_type=None, pop=False, augmenter=aug, encoder_lambda=0.0,
                    random_process=None, noise_clip=None, aug_mix=0.0)
                if k % 2 == 0:
                    lu.soft_update(target.critics[0], agent.critics[0], 0.005)
            float(logs["losses/critic_overall_loss"])
        finally:
            graphed.enable_auto_graphs(False)
        torch.cuda.synchronize()
        return agent._critic_arena.flat.clone(), target._critic_arena.flat.clone()

    c0, t0 = run("classic", False)
    for auto in (False, True):
        c1, t1 = run("nstep", auto)
        assert torch.equal(c0, c1) and torch.equal(t0, t1), f"auto graphs {auto}"
