"""PopArt target normalisation with device-resident state (reference popart.py:8-59).

``mu, nu, w, b`` live in one float32[4] device tensor and ``(t, stable)`` in an int32[2] one; the TD-target
kernel (ssac_td_target) de-normalises, updates the statistics (ART), preserves outputs (POP, gated by the
reference's ``_stable`` heuristic) and re-normalises without the host round trip of popart.py:45-47.
Like the reference, the statistics are plain attributes, not buffers: ``state_dict()`` is empty, so
checkpoints written by either implementation load in the other.
"""
import torch
from torch import nn


class PopArtLayer(nn.Module):
    def __init__(self, beta=1e-4, min_steps=1000, init_nu=0):
        super().__init__()
        self.beta = beta
        self.min_steps = min_steps
        self._state = torch.tensor([0.0, float(init_nu), 1.0, 0.0], dtype=torch.float32)
        self._ctl = torch.tensor([1, 0], dtype=torch.int32)

    # reference attribute surface -----------------------------------------------------------
    mu = property(lambda self: self._state[0:1], lambda self, v: self._state[0:1].copy_(torch.as_tensor(v).reshape(1)))
    nu = property(lambda self: self._state[1:2], lambda self, v: self._state[1:2].copy_(torch.as_tensor(v).reshape(1)))
    w = property(lambda self: self._state[2:3], lambda self, v: self._state[2:3].copy_(torch.as_tensor(v).reshape(1)))
    b = property(lambda self: self._state[3:4], lambda self, v: self._state[3:4].copy_(torch.as_tensor(v).reshape(1)))

    @property
    def _t(self):
        return int(self._ctl[0])

    @_t.setter
    def _t(self, v):
        self._ctl[0] = int(v)

    @property
    def _stable(self):
        return bool(self._ctl[1])

    @property
    def sigma(self):
        return (torch.sqrt(self.nu - self.mu**2) + 1e-5).clamp(1e-4, 1e6)

    def normalize_values(self, val):
        return (val - self.mu) / self.sigma

    def to(self, device):
        self._state = self._state.to(device)
        self._ctl = self._ctl.to(device)
        return self

    def __deepcopy__(self, memo):
        new = PopArtLayer(self.beta, self.min_steps)
        new._state = self._state.clone()
        new._ctl = self._ctl.clone()
        new.training = self.training
        return new

    def forward(self, x, normalized=True):
        out = (self.w * x) + self.b
        return out if normalized else (self.sigma * out) + self.mu

    # raw pointers for the C ABI
    def state_ptr(self):
        return self._state.data_ptr()

    def ctl_ptr(self):
        return self._ctl.data_ptr()
