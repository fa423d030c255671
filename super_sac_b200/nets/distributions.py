"""Action distributions returned by the actor modules (reference nets/distributions.py:9-15, :64-114).

These objects serve the *acting* path (Agent.forward / sample_action, out of the update hot path) and API
compatibility.  The update path never builds them: sampling, log-probs and their gradients are computed by
the fused head kernels (ssac_tanh_normal_* / ssac_det_head_* in include/ssac_b200.h).
"""
import math

import torch
import torch.nn.functional as F

_LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


class SquashedNormal:
    """tanh(Normal(loc, scale)); ``log_prob`` of a value it did not produce goes through atanh(clamp(+-0.99))."""

    def __init__(self, loc, scale):
        self.loc, self.scale = loc, scale
        self._cache = None

    @property
    def mean(self):
        return torch.tanh(self.loc)

    def _draw(self):
        x = self.loc + self.scale * torch.randn_like(self.loc)
        y = torch.tanh(x)
        self._cache = (x, y)
        return y

    def rsample(self):
        return self._draw()

    def sample(self):
        with torch.no_grad():
            return self._draw()

    def log_prob(self, value):
        if self._cache is not None and value is self._cache[1]:
            x = self._cache[0]
        else:
            y = value.clamp(-0.99, 0.99)
            x = 0.5 * (y.log1p() - (-y).log1p())
        ladj = 2.0 * (math.log(2.0) - x - F.softplus(-2.0 * x))
        nlp = -((x - self.loc) ** 2) / (2 * self.scale**2) - self.scale.log() - _LOG_SQRT_2PI
        return (0.0 - ladj) + nlp


def create_tanh_normal(vec, log_std_low, log_std_high):
    mu, log_std = vec.chunk(2, dim=-1)
    log_std = torch.tanh(log_std)
    log_std = log_std_low + 0.5 * (log_std_high - log_std_low) * (log_std + 1)
    return SquashedNormal(mu, log_std.exp())


class ContinuousDeterministic(torch.distributions.Normal):
    """Normal(loc, 1e-4) whose ``sample()`` is the mode (reference nets/distributions.py:107-114)."""

    def __init__(self, deterministic_actor_output):
        super().__init__(loc=deterministic_actor_output, scale=1e-4, validate_args=False)

    def sample(self):
        return self.loc
