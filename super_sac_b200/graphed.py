"""CUDA-graph replay of the update entry points.

One REDQ-10 critic update is ~25 short kernels; launched one by one from Python the step is bound by launch and
interpreter latency, not by the GPU.  ``GraphedCall`` runs any of the update functions once under
``torch.cuda.graph`` -- every kernel of libssac_b200 is capturable: no allocation, no sync, and everything that
changes between replays (Adam step, PopArt state, Philox offset, log_alpha, buffer fill level) lives in device
memory -- and afterwards replays the whole step with a single ``cudaGraphLaunch``.

    step = GraphedCall(lambda: learning.critic_update(buffer=..., agent=..., ...))
    step.replay()            # one update, no host sync
    logs = step.logs()       # one device->host copy, reference log keys

Restrictions (checked by the caller): static shapes and hyper-parameters, the default Philox randomness source, and
no host-side decisions inside the captured function (``random.choice`` for the logged member is frozen at capture).
"""
import torch

from . import _logs


class GraphedCall:
    def __init__(self, fn, warmup=2):
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):  # allocator / lazy state (Adam moments, rng state) settles before capture
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with _logs.deferred():
            with torch.cuda.graph(self.graph):
                self.result = fn()
        self.n_replays = 0

    def replay(self):
        self.graph.replay()
        self.n_replays += 1

    def _logs_obj(self):
        r = self.result
        if isinstance(r, tuple):
            r = r[0]
        return r

    def logs(self):
        """Logged scalars of the most recent replay (one D2H copy + sync)."""
        return dict(self._logs_obj().fetch(keep=True))
