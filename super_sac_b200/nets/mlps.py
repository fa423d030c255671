"""The 3-Linear MLP heads of the agent (reference nets/mlps.py:11-41, :78-93, :113-129; discrete: :132-185).

Layer names (fc1, fc2, out / fc3) and shapes match the reference so ``state_dict()`` keys are interchangeable.
Inside an ``Agent`` the Parameters of these modules are views into one flat fp32 arena (see _arena.py) and the
update path runs them as one grouped launch (ssac_mlp_forward / ssac_mlp_backward); the ``forward`` methods
below are the per-module acting path (B = num_envs).
"""
import torch
import torch.nn.functional as F
from torch import nn

from . import distributions, weight_init


class ContinuousStochasticActor(nn.Module):
    def __init__(self, state_size, action_size, log_std_low=-10.0, log_std_high=2.0, hidden_size=256, dist_impl="pyd"):
        super().__init__()
        if dist_impl != "pyd":
            raise NotImplementedError("beta-distribution policies are out of scope (unused by every shipped config)")
        self.fc1 = nn.Linear(state_size, hidden_size)
        self.fc2 = nn.Linear(hidden_size, hidden_size)
        self.fc3 = nn.Linear(hidden_size, 2 * action_size)
        self.log_std_low = log_std_low
        self.log_std_high = log_std_high
        self.apply(weight_init)
        self.dist_impl = dist_impl

    def forward(self, state):
        x = F.relu(self.fc1(state))
        x = F.relu(self.fc2(x))
        return distributions.create_tanh_normal(self.fc3(x), self.log_std_low, self.log_std_high)


class ContinuousDeterministicActor(nn.Module):
    def __init__(self, state_size, action_size, hidden_size=256, **kwargs):
        super().__init__()
        self.fc1 = nn.Linear(state_size, hidden_size)
        self.fc2 = nn.Linear(hidden_size, hidden_size)
        self.out = nn.Linear(hidden_size, action_size)
        self.apply(weight_init)
        self.dist_impl = "deterministic"

    def forward(self, state):
        x = F.relu(self.fc1(state))
        x = F.relu(self.fc2(x))
        return distributions.ContinuousDeterministic(torch.tanh(self.out(x)))


class ContinuousCritic(nn.Module):
    def __init__(self, state_size, action_size, hidden_size=256):
        super().__init__()
        self.fc1 = nn.Linear(state_size + action_size, hidden_size)
        self.fc2 = nn.Linear(hidden_size, hidden_size)
        self.features = None
        self.out = nn.Linear(hidden_size, 1)
        self.apply(weight_init)

    def forward(self, state, action):
        x = torch.cat((state, action), dim=-1)
        x = F.relu(self.fc1(x))
        x = F.relu(self.fc2(x))
        self.features = x
        return self.out(x)


class ContinuousInverseModel(nn.Module):
    """Kept so Agent exposes ``inverse_model`` like the reference (main.py:216-224 builds an optimiser over it);
    the Markov-abstraction update that trains it is out of scope."""

    def __init__(self, state_size, action_size, log_std_low=-10.0, log_std_high=2.0, hidden_size=256, dist_impl="pyd"):
        super().__init__()
        self.fc1 = nn.Linear(state_size * 2, hidden_size)
        self.fc2 = nn.Linear(hidden_size, hidden_size)
        self.fc3 = nn.Linear(hidden_size, 2 * action_size)
        self.log_std_low = log_std_low
        self.log_std_high = log_std_high
        self.apply(weight_init)

    def forward(self, state, next_state):
        x = F.relu(self.fc1(torch.cat((state, next_state), dim=-1)))
        x = F.relu(self.fc2(x))
        return distributions.create_tanh_normal(self.fc3(x), self.log_std_low, self.log_std_high)


class _Categorical:
    """What the acting path and the reference API read from Categorical(logits=...) (nets/mlps.py:147-148):
    normalised ``logits``, ``probs``, ``sample()`` / ``log_prob`` / ``entropy``.  The update path never builds it: the
    softmax and its gradient live in the ssac_discrete_* kernels."""

    def __init__(self, logits):
        self.logits = logits - logits.logsumexp(dim=-1, keepdim=True)
        self.probs = torch.softmax(logits, dim=-1)

    def sample(self):
        return torch.multinomial(self.probs.reshape(-1, self.probs.shape[-1]), 1).reshape(self.probs.shape[:-1])

    def log_prob(self, value):
        return self.logits.gather(-1, value.long().unsqueeze(-1)).squeeze(-1)

    def entropy(self):
        return -(self.probs * self.logits).sum(-1)


class DiscreteActor(nn.Module):
    """S -> H -> H -> A logits (nets/mlps.py:132-149); the last layer keeps the reference's name ``act_p``."""

    def __init__(self, state_size, action_size, hidden_size=256):
        super().__init__()
        self.fc1 = nn.Linear(state_size, hidden_size)
        self.fc2 = nn.Linear(hidden_size, hidden_size)
        self.act_p = nn.Linear(hidden_size, action_size)
        self.apply(weight_init)

    def forward(self, state):
        x = F.relu(self.fc1(state))
        x = F.relu(self.fc2(x))
        return _Categorical(self.act_p(x))


class DiscreteCritic(nn.Module):
    """S -> H -> H -> A action values (nets/mlps.py:170-185)."""

    def __init__(self, state_size, action_size, hidden_size=300):
        super().__init__()
        self.fc1 = nn.Linear(state_size, hidden_size)
        self.fc2 = nn.Linear(hidden_size, hidden_size)
        self.out = nn.Linear(hidden_size, action_size)
        self.features = None
        self.apply(weight_init)

    def forward(self, state):
        x = F.relu(self.fc1(state))
        x = F.relu(self.fc2(x))
        self.features = x
        return self.out(x)


class DiscreteInverseModel(nn.Module):
    """Kept so a discrete Agent exposes ``inverse_model`` like the reference (nets/mlps.py:152-167); the
    Markov-abstraction update that trains it is out of scope."""

    def __init__(self, state_size, action_size, hidden_size, **kwargs):
        super().__init__()
        self.fc1 = nn.Linear(state_size * 2, hidden_size)
        self.fc2 = nn.Linear(hidden_size, hidden_size)
        self.act_p = nn.Linear(hidden_size, action_size)
        self.apply(weight_init)

    def forward(self, state, next_state):
        x = F.relu(self.fc1(torch.cat((state, next_state), dim=-1)))
        x = F.relu(self.fc2(x))
        return _Categorical(self.act_p(x))


class ContrastiveModel(nn.Module):
    def __init__(self, state_size, hidden_size=256):
        super().__init__()
        self.fc1 = nn.Linear(state_size * 2, hidden_size)
        self.fc2 = nn.Linear(hidden_size, hidden_size)
        self.out = nn.Linear(hidden_size, 1)
        self.apply(weight_init)

    def forward(self, states, next_states):
        x = F.relu(self.fc1(torch.cat((states, next_states), dim=-1)))
        x = F.relu(self.fc2(x))
        return torch.sigmoid(self.out(x))
