"""Network definitions.  Mirrors the reference's ``super_sac.nets`` surface (nets/__init__.py:4-35):
``weight_init``, the ``Encoder`` plugin base class, and the ``mlps`` / ``cnns`` / ``distributions`` modules.

The encoder is a *user plugin* (arbitrary nn.Module taking the observation dict): it stays plain PyTorch and
sits upstream of the CUDA path, which returns ``grad(s_rep)`` to it through autograd.
"""
from abc import abstractmethod

from torch import nn


def weight_init(m):
    """Orthogonal init for Linear, delta-orthogonal for conv (reference nets/__init__.py:4-16)."""
    if isinstance(m, nn.Linear):
        nn.init.orthogonal_(m.weight.data)
        m.bias.data.fill_(0.0)
    elif isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
        assert m.weight.size(2) == m.weight.size(3)
        m.weight.data.fill_(0.0)
        m.bias.data.fill_(0.0)
        mid = m.weight.size(2) // 2
        nn.init.orthogonal_(m.weight.data[:, :, mid, mid], nn.init.calculate_gain("relu"))


class Encoder(nn.Module):
    """Plugin boundary: subclass, implement ``forward(obs_dict)`` and ``embedding_dim``
    (reference nets/__init__.py:21-35)."""

    def __init__(self):
        super().__init__()
        self.have_at_least_one_param = nn.Linear(1, 1)

    def forward_rolling(self, obs):
        return self.forward(obs)

    def reset_rolling(self):
        pass

    @property
    @abstractmethod
    def embedding_dim(self):
        raise NotImplementedError


from . import mlps, cnns, distributions  # noqa: E402,F401
