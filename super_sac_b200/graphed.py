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


# ------------------------------------------------------------------------------------------------
# transparent graph replay behind the drop-in entry points
# ------------------------------------------------------------------------------------------------
_auto = {"on": False, "cache": {}, "min_calls": 2, "lazy": False}


def enable_auto_graphs(on=True, lazy_logs=False):
    """After ``enable_auto_graphs()``, ``learning.critic_update`` / ``online_actor_update`` / ``alpha_update`` replay a
    captured CUDA graph from the third call with identical arguments on (same objects, same hyper-parameters): the call
    a user makes stays ``learning.critic_update(...)``, the ~20 launches and their Python marshalling collapse into one
    ``cudaGraphLaunch`` + one device->host copy of the logged scalars.  ``lazy_logs``: the returned dict materialises its
    values on first access instead of synchronising inside the call (``_logs.LazyLogs``)."""
    _auto["on"] = bool(on)
    _auto["lazy"] = bool(lazy_logs) and bool(on)
    if not on:
        _auto["cache"].clear()


def auto_graphs_enabled():
    return _auto["on"]


class _Entry:
    __slots__ = ("calls", "graph", "result", "on_replay", "refs", "pending", "event")

    def __init__(self):
        self.calls, self.graph, self.result, self.on_replay, self.refs = 0, None, None, None, None
        self.pending, self.event = None, None


def is_static(obj):
    """True for objects (replay dicts) returned by a captured graph: the same buffers on every call."""
    return getattr(obj, "_ssac_static", False)


def run_cached(key, fn, on_replay=None, refs=None):
    """Eager for the first ``min_calls`` calls with this key (allocator and lazy state settle), then capture once
    (capturing does not execute) and replay.  ``on_replay`` updates host-side mirrors (step counters).  ``refs``: the
    objects whose ``id()`` went into ``key`` -- the entry keeps them alive, so that CPython cannot hand one of those ids
    to a new object while the entry exists (a key built from the ids of dead objects could otherwise match by accident).
    The replay dicts a captured call returns are tagged static (``is_static``): they are the graph's own buffers."""
    e = _auto["cache"].get(key)
    if e is None:
        e = _auto["cache"][key] = _Entry()
        e.refs = refs
    if e.graph is None:
        if e.calls < _auto["min_calls"]:
            e.calls += 1
            return fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with _logs.deferred(embed_readback=True):   # this path fetches the logs after every replay
            with torch.cuda.graph(g):
                e.result = fn()
        e.graph, e.on_replay = g, on_replay
        if isinstance(e.result, tuple):
            for part in e.result[1:]:
                for rd in (part if isinstance(part, (list, tuple)) else [part]):
                    try:
                        rd._ssac_static = True
                    except AttributeError:
                        pass
    if e.pending is not None:    # the previous replay's logs live in the pinned buffer this replay overwrites
        e.pending.resolve()
        e.pending = None
    e.graph.replay()
    if e.on_replay is not None:
        e.on_replay()
    res = e.result
    logs = res[0] if isinstance(res, tuple) else res
    if _auto["lazy"] and logs._host is not None:
        if e.event is None:
            e.event = torch.cuda.Event()
        e.event.record()
        out = e.pending = _logs.LazyLogs(logs, e.event)
    else:
        out = dict(logs.fetch(keep=True))
    return (out,) + tuple(res[1:]) if isinstance(res, tuple) else out
