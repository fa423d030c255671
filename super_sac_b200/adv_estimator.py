"""Advantage estimator A(s,a) = Q(s,a) - V(s) (reference adv_estimator.py:8-89: continuous 'mean' / 'max', discrete 'indirect').

Kept as an nn.Module with the reference's constructor so ``agent.adv_estimator(o, a, ensemble_idx)`` keeps working;
the arithmetic runs on the grouped critic / actor kernels (see learning_utils._advantage)."""
import torch
from torch import nn


class AdvantageEstimator(nn.Module):
    def __init__(self, encoder, actors, critics, popart=False, discrete_method="indirect", continuous_method="mean",
                 discrete=False):
        super().__init__()
        assert continuous_method in ["mean", "max"]
        assert discrete_method in ["indirect"]   # adv_estimator.py:41-43: 'direct' raises in the reference too
        # plain attributes (not sub-modules): the Agent owns these objects
        object.__setattr__(self, "encoder", encoder)
        object.__setattr__(self, "actors", actors)
        object.__setattr__(self, "critics", critics)
        object.__setattr__(self, "popart", popart)
        self.cont_method = continuous_method
        self.discrete_method = discrete_method
        self.discrete = bool(discrete)
        self._agent = None

    def bind(self, agent):
        object.__setattr__(self, "_agent", agent)

    def forward(self, obs, action, ensemble_idx, n=4):
        from . import learning_utils as lu

        if self._agent is None:
            raise RuntimeError("AdvantageEstimator must be created by an Agent")
        rd = {"primary_batch": (obs, action, None, None, None)}
        adv, _, _ = lu._advantage(self._agent, rd, ensemble_idx, n=n)
        return adv.unsqueeze(-1)
