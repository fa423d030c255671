"""Import the UNMODIFIED reference (jakegrigsby/super_sac) and drive it with injected randomness.

This file only works where ``/root/reference`` exists (the build container).  It is used by
``make_golden.py`` to generate the committed fixtures and by ``tests/test_oracle_vs_reference.py``
(skipped when the reference is absent, e.g. on the GPU box).  Nothing in the product imports it.

The reference's ``import super_sac`` pulls in gin / gymnasium / gym / tensorboardX / skimage, none
of which are installed here; five stub modules make the import succeed and leave the whole update
path (learning.py, learning_utils.py, agent.py, replay.py, augmentations.py, popart.py,
adv_estimator.py, nets/*) running unchanged.

Randomness is spread over torch-CPU, torch-device, python ``random`` and numpy.  The
``injected`` context manager replaces exactly the draw sites on the update path with FIFO queues so
the reference, the oracle and the CUDA path can all consume the same indices / subsets / eps:

  torch.randint   replay.py:122 (indices), augmentations.py:180-181,227-231 (shifts)
  random.sample   agent.py:29 (REDQ subset)
  Normal.sample / Normal.rsample   torch.distributions (actor sampling) -> loc + eps * scale
  torch.randn     learning_utils.py:49 (TD3 noise)
  torch.randn_like augmentations.py:203 (DrqAug noise)
  np.random.random replay.py:166 (PER mass)
"""
import contextlib
import os
import random as _py_random
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("SSAC_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "super_sac"))


def import_reference():
    """Returns the reference ``super_sac`` package (device forced to CPU)."""
    if "super_sac" in sys.modules and getattr(sys.modules["super_sac"], "_ssac_ref", False):
        return sys.modules["super_sac"]

    def _stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    _stub("gin", configurable=lambda x=None, **k: x if callable(x) else (lambda f: f))

    class _W:
        pass

    g = _stub("gymnasium", Wrapper=_W, ActionWrapper=_W, ObservationWrapper=_W, RewardWrapper=_W, Env=object)
    g.spaces = _stub("gymnasium.spaces")
    _stub("gym", Wrapper=_W)
    _stub("tensorboardX")
    _stub("skimage")
    _stub("skimage.transform", resize=None)
    _stub("skimage.util")
    _stub("skimage.util.shape", view_as_windows=None)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import super_sac  # noqa
    finally:
        sys.path.remove(REFERENCE_ROOT)
    # force CPU regardless of what the box has (each module did ``from . import device``)
    for mod in (super_sac, super_sac.learning, super_sac.learning_utils, super_sac.agent):
        mod.device = "cpu"
    super_sac._ssac_ref = True
    return super_sac


class _Queue:
    def __init__(self, name, items):
        self.name = name
        self.items = list(items or [])
        self.used = 0

    def pop(self):
        if not self.items:
            raise RuntimeError(f"injected randomness queue '{self.name}' exhausted after {self.used} draws")
        self.used += 1
        return self.items.pop(0)


@contextlib.contextmanager
def injected(randint=None, subsets=None, normal_eps=None, randn=None, randn_like=None, np_random=None):
    """Replace the reference's random draws by scripted values (see module docstring)."""
    import torch.distributions as pyd

    q_randint = _Queue("randint", randint)
    q_subset = _Queue("subsets", subsets)
    q_eps = _Queue("normal_eps", normal_eps)
    q_randn = _Queue("randn", randn)
    q_randn_like = _Queue("randn_like", randn_like)
    q_np = _Queue("np_random", np_random)

    o_randint, o_randn, o_randn_like = torch.randint, torch.randn, torch.randn_like
    o_sample, o_nsample, o_nrsample = _py_random.sample, pyd.Normal.sample, pyd.Normal.rsample
    o_nprandom = np.random.random

    def f_randint(*args, **kwargs):
        out = torch.as_tensor(q_randint.pop()).long()
        size = kwargs.get("size", None)
        if size is None:
            size = args[-1]
        assert tuple(out.shape) == tuple(size), (out.shape, size)
        return out

    def f_sample(population, k):
        out = list(q_subset.pop())
        assert len(out) == k
        return out

    def f_nsample(self, sample_shape=torch.Size()):
        assert tuple(sample_shape) == ()
        eps = torch.as_tensor(q_eps.pop(), dtype=self.loc.dtype)
        with torch.no_grad():
            return self.loc + eps * self.scale

    def f_nrsample(self, sample_shape=torch.Size()):
        assert tuple(sample_shape) == ()
        eps = torch.as_tensor(q_eps.pop(), dtype=self.loc.dtype)
        return self.loc + eps * self.scale

    def f_randn(*shape, **kw):
        out = torch.as_tensor(q_randn.pop(), dtype=torch.float32)
        assert tuple(out.shape) == tuple(shape), (out.shape, shape)
        return out

    def f_randn_like(x, **kw):
        out = torch.as_tensor(q_randn_like.pop(), dtype=x.dtype)
        assert out.shape == x.shape
        return out

    def f_nprandom(size=None):
        out = np.asarray(q_np.pop(), dtype=np.float64)
        assert out.shape == (size,) or out.shape == tuple(np.atleast_1d(size))
        return out

    torch.randint, torch.randn, torch.randn_like = f_randint, f_randn, f_randn_like
    _py_random.sample = f_sample
    pyd.Normal.sample, pyd.Normal.rsample = f_nsample, f_nrsample
    np.random.random = f_nprandom
    try:
        yield dict(randint=q_randint, subsets=q_subset, normal_eps=q_eps, randn=q_randn,
                   randn_like=q_randn_like, np_random=q_np)
    finally:
        torch.randint, torch.randn, torch.randn_like = o_randint, o_randn, o_randn_like
        _py_random.sample = o_sample
        pyd.Normal.sample, pyd.Normal.rsample = o_nsample, o_nrsample
        np.random.random = o_nprandom


class ActionSpace:
    """The only thing GaussianExplorationNoise needs (learning_utils.py:35-36)."""

    def __init__(self, dim):
        self.low = -np.ones(dim, dtype=np.float32)
        self.high = np.ones(dim, dtype=np.float32)
        self.shape = (dim,)
