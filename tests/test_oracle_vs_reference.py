"""Live cross-check of the oracle against the UNMODIFIED reference on FRESH seeds and shapes (CPU; runs only where
/root/reference exists, i.e. in the build container -- the committed fixtures in tests/golden/ are what travels).

The golden generator (tests/golden/make_golden.py) drives the imported reference with injected randomness; here it writes
new fixtures into a temporary directory with seeds / sizes that no committed fixture uses, and the same checks as
tests/test_oracle_golden.py run on them.  This guards the oracle against having been fitted to the committed cases.
"""
import os
import sys

import pytest

import golden_util as gu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_harness as rh  # noqa: E402

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="the reference tree is not present on this machine")

FRESH_UPDATE = {
    "redq_fresh": dict(E=1, N=4, M=3, S=9, A=4, H=40, B=24, steps=2, target_delay=1, seed=101),
    "sunrise_fresh": dict(E=2, N=2, M=2, S=7, A=3, H=24, B=12, steps=2, weight_type="sunrise", weight_temp=15.0,
                          popart=True, pop=True, popart_warm=True, seed=102),
}
FRESH_DISCRETE = {
    "discrete_fresh": dict(E=1, N=2, M=1, S=10, A=7, H=24, B=20, steps=3, seed=103),
    "discrete_ens_fresh": dict(E=3, N=2, M=2, S=5, A=4, H=16, B=12, steps=2, weight_type="sunrise", weight_temp=10.0,
                               popart=True, pop=True, popart_warm=True, dr3_coeff=0.02, actor_clip=1.0, seed=104),
}


@pytest.mark.parametrize("name", sorted(FRESH_UPDATE))
def test_update_oracle_on_fresh_reference_runs(name, tmp_path):
    import make_golden as mg
    from test_oracle_golden import check_update_oracle

    mg.run_update_case(name, FRESH_UPDATE[name], out_dir=str(tmp_path))
    check_update_oracle(gu.load("update_" + name, directory=str(tmp_path)))


@pytest.mark.parametrize("name", sorted(FRESH_DISCRETE))
def test_discrete_oracle_on_fresh_reference_runs(name, tmp_path):
    import make_golden as mg
    from test_oracle_golden import check_discrete_oracle

    mg.run_discrete_case(name, FRESH_DISCRETE[name], out_dir=str(tmp_path))
    check_discrete_oracle(gu.load("update_" + name, directory=str(tmp_path)))


def test_discrete_checkpoints_interchange_with_the_reference(tmp_path):
    """Agent.save / load of a discrete agent use the reference's file names and state_dict keys (agent.py:172-202): a
    checkpoint written by the unmodified reference loads into this package's Agent (the values land in the flat arenas)
    and one written by this package loads back into the reference."""
    import numpy as np
    import torch

    import super_sac_b200 as ssb

    ref = rh.import_reference()

    def build(pkg):
        class Enc(pkg.nets.Encoder):
            def __init__(self):
                super().__init__()

            @property
            def embedding_dim(self):
                return 6

            def forward(self, obs):
                return obs["obs"]

        return pkg.Agent(act_space_size=5, encoder=Enc(), actor_network_cls=pkg.nets.mlps.DiscreteActor,
                         critic_network_cls=pkg.nets.mlps.DiscreteCritic, discrete=True, ensemble_size=2, num_critics=2,
                         hidden_size=16, auto_rescale_targets=False)

    torch.manual_seed(3)
    theirs, ours = build(ref), build(ssb)
    d1, d2 = tmp_path / "ref", tmp_path / "ours"
    d1.mkdir(), d2.mkdir()
    theirs.save(str(d1))
    ours.load(str(d1))
    for i in range(2):
        want = theirs.actors[i].state_dict()
        got = ours.actors[i].state_dict()
        assert set(want) == set(got)
        for k in want:
            assert torch.equal(want[k], got[k]), k
        for k, v in theirs.critics[i].state_dict().items():
            assert torch.equal(v, ours.critics[i].state_dict()[k]), k
    # the loaded values live in the arenas the kernels read (net g = member * N + k)
    assert torch.equal(ours._critic_arena.p["W3"][3], theirs.critics[1].nets[1].out.weight)
    assert torch.equal(ours._actor_arena.p["W3"][1], theirs.actors[1].act_p.weight)
    with torch.no_grad():
        ours._critic_arena.flat.mul_(0.5)
    ours.save(str(d2))
    theirs.load(str(d2))
    assert torch.equal(theirs.critics[0].nets[0].fc1.weight, ours._critic_arena.p["W1"][0])
    assert np.isfinite(theirs.critics[0].nets[0].fc1.weight.detach().numpy()).all()


def test_epsilon_greedy_process_follows_the_reference():
    """learning_utils.EpsilonGreedyExplorationNoise (the discrete agents' exploration process, main.py:246-253): the same
    python / numpy random streams give the same actions and the same epsilon schedule as the reference's class."""
    import random

    import numpy as np

    from super_sac_b200 import learning_utils as lu

    ref = rh.import_reference()

    class Space:
        n = 6

    outs = []
    for cls in (ref.learning_utils.EpsilonGreedyExplorationNoise, lu.EpsilonGreedyExplorationNoise):
        random.seed(5)
        np.random.seed(5)
        proc = cls(Space(), eps_start=0.9, eps_final=0.05, steps_annealed=40)
        seq = []
        for t in range(60):
            a = proc.sample(np.array([t % 6], dtype=np.int64), update_schedule=(t % 3 != 0))
            seq.append((int(a[0]), proc.current_scale))
        outs.append(seq)
    assert outs[0] == outs[1]
