"""CPU-only checks of the host side: the C-ABI library loads and exports every declared symbol, the flat arena keeps
the reference's parameter surface (names, state_dict, deepcopy, optimizer construction), and the product refuses to
run without a CUDA device (no CPU fallback)."""
import copy
import ctypes
import os

import numpy as np
import pytest
import torch

import super_sac_b200 as ssb
from super_sac_b200 import _arena, _lib, nets


class IdentityEncoder(nets.Encoder):
    def __init__(self, dim):
        super().__init__()
        self._dim = dim

    @property
    def embedding_dim(self):
        return self._dim

    def forward(self, obs_dict):
        return obs_dict["obs"]


def make_agent(E=2, N=3, S=17, A=6, H=32, det=False, popart=True):
    return ssb.Agent(A, IdentityEncoder(S), nets.mlps.ContinuousDeterministicActor if det else nets.mlps.ContinuousStochasticActor,
                     nets.mlps.ContinuousCritic, ensemble_size=E, num_critics=N, hidden_size=H, auto_rescale_targets=popart)


def test_library_exports_every_declared_symbol():
    protos = _lib.parse_header()
    assert len(protos) >= 30
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(cdll, name), f"{name} declared in include/ssac_b200.h but not exported"
    lib = _lib.lib()
    assert lib.version() >= 100
    assert lib.mlp_backward_ws(10, 256, 256) == 2 * 10 * 256 * 256


def test_bad_arguments_return_error_codes_not_crashes():
    lib = _lib.lib()
    with pytest.raises(_lib.SsacError, match="null pointer"):
        lib.polyak(None, None, 10, 0.005, None)
    with pytest.raises(_lib.SsacError, match="power of two"):
        lib.tree_set(1, 1, 12, 1, 1, 4, None)
    with pytest.raises(_lib.SsacError, match="M <= N"):
        lib.rng_fill(1, None, 0, 0, None, None, 0, 1, 1, 4, 9, None, 0, 0, None, 0, None)


def test_arena_keeps_reference_parameter_surface():
    agent = make_agent()
    ca = agent._critic_arena
    # names / shapes of the reference's state_dict (agent.py:13-18, nets/mlps.py:113-129)
    sd = agent.critics[1].state_dict()
    assert list(sd.keys()) == [f"nets.{k}.{l}.{p}" for k in range(3) for l in ("fc1", "fc2", "out") for p in ("weight", "bias")]
    assert sd["nets.0.fc1.weight"].shape == (32, 23) and sd["nets.2.out.weight"].shape == (1, 32)
    assert list(agent.actors[0].state_dict().keys()) == ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc3.weight", "fc3.bias"]
    # every parameter is a view into the flat arena, net-major, member i at nets [i*N, (i+1)*N)
    p = agent.critics[1].nets[2].fc2.weight
    assert p.data_ptr() == ca.p["W2"][1 * 3 + 2].data_ptr()
    assert p.grad.data_ptr() == ca.g["W2"][5].data_ptr()
    p.data.fill_(3.0)
    assert float(ca.p["W2"][5].min()) == 3.0
    for n in _arena.NAMES:
        assert ca.offsets[n] % 32 == 0  # 128-byte aligned arrays
    n_params = sum(q.numel() for c in agent.critics for q in c.parameters())
    assert n_params == 6 * (23 * 32 + 32 + 32 * 32 + 32 + 32 + 1) and ca.numel >= n_params


def test_deepcopy_and_load_state_dict_keep_aliasing():
    agent = make_agent()
    target = copy.deepcopy(agent)  # main.py:321
    assert target._critic_arena is not agent._critic_arena
    assert torch.equal(target._critic_arena.flat, agent._critic_arena.flat)
    agent.critics[0].nets[0].fc1.weight.data.add_(1.0)
    assert not torch.equal(target._critic_arena.flat, agent._critic_arena.flat)
    target.critics[0].load_state_dict(agent.critics[0].state_dict())
    assert torch.equal(target._critic_arena.p["W1"][0], agent._critic_arena.p["W1"][0])
    assert target.critics[0].nets[0].fc1.weight.data_ptr() == target._critic_arena.p["W1"][0].data_ptr()
    assert target.popart[0] is not agent.popart[0] and target.adv_estimator._agent is target


def test_caller_built_adam_is_recognised():
    from itertools import chain

    agent = make_agent()
    opt = torch.optim.Adam(chain(*(c.parameters() for c in agent.critics)), lr=3e-4, betas=(0.9, 0.999))  # main.py:188-193
    fa = _arena.FlatAdam.attach(opt, agent._critic_arena)
    assert _arena.FlatAdam.attach(opt, agent._critic_arena) is fa
    st = opt.state_dict()["state"]
    assert len(st) == 2 * 3 * 6 and set(st[0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
    assert fa.hyper() == (3e-4, 0.9, 0.999, 1e-8, 0.0)
    other = torch.optim.Adam(agent.actors[0].parameters(), lr=1e-3)
    with pytest.raises(NotImplementedError):
        _arena.FlatAdam.attach(other, agent._critic_arena)


def test_save_load_roundtrip(tmp_path):
    a, b = make_agent(), make_agent()
    a.save(str(tmp_path))
    assert sorted(os.listdir(tmp_path)) == sorted(
        ["encoder.pt", "popart0.pt", "popart1.pt", "critic0.pt", "critic1.pt", "actor0.pt", "actor1.pt", "inverse.pt", "contrastive.pt"])
    b.load(str(tmp_path))
    assert torch.equal(a._critic_arena.flat, b._critic_arena.flat) and torch.equal(a._actor_arena.flat, b._actor_arena.flat)


def test_no_cpu_fallback():
    agent = make_agent()
    with pytest.raises(_lib.SsacError, match="no CPU path"):
        agent.critics[0](torch.zeros(4, 17), torch.zeros(4, 6))
    if not torch.cuda.is_available():
        buf = ssb.replay.ReplayBuffer(16)
        with pytest.raises(_lib.SsacError, match="device-resident"):
            buf.push({"obs": np.zeros(3, np.float32)}, np.zeros(1, np.float32), 0.0, {"obs": np.zeros(3, np.float32)}, False)


def test_unsupported_features_fail_loudly():
    with pytest.raises(NotImplementedError):
        ssb.Agent(2, IdentityEncoder(3), nets.mlps.ContinuousStochasticActor, nets.mlps.ContinuousCritic, discrete=True)

    class Odd(torch.nn.Module):
        def __init__(self, state_size, action_size, hidden_size=8):
            super().__init__()
            self.l = torch.nn.Linear(state_size + action_size, 1)

    with pytest.raises(NotImplementedError):
        ssb.Agent(2, IdentityEncoder(3), nets.mlps.ContinuousStochasticActor, Odd)


def test_big_pixel_encoder_module_shell_on_cpu():
    """nets.cnns.BigPixelEncoder keeps the reference's nn.Module shell (nets/cnns.py:37-69): parameter names / shapes and the
    state_dict of the reference module load into it, its CPU forward reproduces the golden output of the UNMODIFIED reference
    module, and deepcopy / state_dict round trips work (target_agent = deepcopy(agent), Agent.save / load)."""
    import copy

    import numpy as np
    import torch

    import golden_util as gu
    from super_sac_b200.nets import cnns

    fx = gu.load("encoder")
    for tag in ("rgb20", "stack24"):
        obs = fx[f"{tag}/obs"]
        params = gu.sub(fx, f"{tag}/params")
        enc = cnns.BigPixelEncoder(obs.shape[1:], out_dim=fx[f"{tag}/out"].shape[1])
        assert {k: tuple(v.shape) for k, v in enc.state_dict().items()} == {k: tuple(v.shape) for k, v in params.items()}
        enc.load_state_dict({k: torch.as_tensor(v) for k, v in params.items()})
        with torch.no_grad():
            y = enc(torch.as_tensor(obs).float())
        gu.assert_close(y.numpy(), fx[f"{tag}/out"], 2e-5, 2e-6, f"{tag} CPU module forward")
        twin = copy.deepcopy(enc)
        assert all(torch.equal(a, b) for a, b in zip(twin.state_dict().values(), enc.state_dict().values()))
        assert twin._ws is not enc._ws                       # workspaces are per module
        twin.load_state_dict(enc.state_dict())
        assert enc.embedding_dim == fx[f"{tag}/out"].shape[1]
        import io
        import pickle

        buf = io.BytesIO()
        pickle.dump(enc, buf)          # torch.save(module): workspaces stay behind, a fresh pool comes back
        back = pickle.loads(buf.getvalue())
        assert all(torch.equal(a, b) for a, b in zip(back.state_dict().values(), enc.state_dict().values()))
        assert isinstance(back._ws, cnns._Workspace) and not back._ws.free
