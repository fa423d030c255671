"""Where the update path's random draws come from.

The reference spreads its draws over four generators (torch-CPU randint / randn, torch-device Normal sampling,
python ``random.sample``, numpy) -- SURVEY F14 -- so seeds alone cannot reproduce its stream.  Here every draw is
a device tensor produced by one of two sources with the same interface:

* ``PhiloxSource`` (default): ssac_rng_fill, Philox4x32-10 with (seed, offset) in device memory; graph replays
  advance the offset.
* ``ScriptedSource``: tests push the exact tensors the reference consumed (indices, subsets, eps, noise, shifts).

Downstream kernels are identical for both: they only ever see device pointers.
"""
import collections

import torch

from . import _lib


class PhiloxSource:
    def __init__(self, seed=0):
        self.seed = int(seed)
        self._state = {}

    def _rng(self, device):
        st = self._state.get(device)
        if st is None:
            st = torch.tensor([self.seed, 0, 0, 0], dtype=torch.int64, device=device)
            self._state[device] = st
        return st

    def manual_seed(self, seed):
        self.seed = int(seed)
        self._state.clear()

    def fill(self, device, idx=None, n_filled=0, normal=None, subset=None, N=0, M=0, shift=None, shift_range=0,
             n_filled_dev=None, zero=None):
        """Fill any of the given pre-allocated device tensors in ONE launch (``zero``: a float tensor to clear)."""
        L = _lib.lib()
        n_sub = 0 if subset is None else subset.numel() // M
        L.rng_fill(self._rng(device).data_ptr(),
                   None if idx is None else idx.data_ptr(), 0 if idx is None else idx.numel(), int(n_filled),
                   None if n_filled_dev is None else n_filled_dev.data_ptr(),
                   None if normal is None else normal.data_ptr(), 0 if normal is None else normal.numel(),
                   None if subset is None else subset.data_ptr(), n_sub, int(N), int(M),
                   None if shift is None else shift.data_ptr(), 0 if shift is None else shift.numel(), int(shift_range),
                   None if zero is None else zero.data_ptr(), 0 if zero is None else zero.numel(),
                   _lib.stream_ptr())

    # one-tensor conveniences --------------------------------------------------------------
    def indices(self, out, n_filled, n_filled_dev=None):
        self.fill(out.device, idx=out, n_filled=n_filled, n_filled_dev=n_filled_dev)
        return out

    def normal(self, out):
        self.fill(out.device, normal=out)
        return out

    def subsets(self, out, N, M):
        self.fill(out.device, subset=out, N=N, M=M)
        return out

    def shifts(self, out, shift_range):
        self.fill(out.device, shift=out, shift_range=shift_range)
        return out

    def uniform01(self, out):
        """float64 U[0,1) for prioritised sampling (replay.py:166)."""
        out.copy_(torch.rand(out.shape, dtype=torch.float64, device=out.device))
        return out


class ScriptedSource:
    """FIFO queues of pre-made draws; anything not scripted raises."""

    def __init__(self):
        self.q = collections.defaultdict(collections.deque)

    def push(self, kind, tensor):
        self.q[kind].append(tensor)
        return self

    def _pop(self, kind, out):
        if not self.q[kind]:
            raise RuntimeError(f"ScriptedSource: no scripted '{kind}' draw left")
        t = torch.as_tensor(self.q[kind].popleft())
        out.copy_(t.reshape(out.shape).to(device=out.device, dtype=out.dtype))
        return out

    def indices(self, out, n_filled, n_filled_dev=None):
        return self._pop("indices", out)

    def normal(self, out):
        return self._pop("normal", out)

    def subsets(self, out, N, M):
        return self._pop("subsets", out)

    def shifts(self, out, shift_range):
        return self._pop("shifts", out)

    def uniform01(self, out):
        return self._pop("uniform01", out)

    def fill(self, device, idx=None, n_filled=0, normal=None, subset=None, N=0, M=0, shift=None, shift_range=0,
             n_filled_dev=None, zero=None):
        if zero is not None:
            zero.zero_()
        if idx is not None:
            self.indices(idx, n_filled)
        if normal is not None:
            self.normal(normal)
        if subset is not None:
            self.subsets(subset, N, M)
        if shift is not None:
            self.shifts(shift, shift_range)

    def empty(self):
        return all(len(v) == 0 for v in self.q.values())


_source = PhiloxSource(0)


def source():
    return _source


def set_source(src):
    global _source
    old = _source
    _source = src
    return old


def manual_seed(seed):
    if isinstance(_source, PhiloxSource):
        _source.manual_seed(seed)
