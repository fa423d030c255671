"""super_sac_b200 -- the off-policy update step of jakegrigsby/super_sac as hand-written sm_100a CUDA.

Drop-in surface (same names as the reference package, reference super_sac/__init__.py:5-11):
``Agent``, ``nets`` (``Encoder`` plugin base, ``mlps``, ``cnns``), ``replay.ReplayBuffer``,
``learning.critic_update / online_actor_update / alpha_update / offline_actor_update``,
``learning_utils.soft_update / hard_update / sample_move_and_augment``, ``augmentations``, ``popart``.

Host code is Python/PyTorch plumbing (tensors, streams, autograd hand-off to user encoders); all arithmetic of
the path runs in ``libssac_b200.so`` through the C ABI in ``include/ssac_b200.h``.  There is no CPU or PyTorch
fallback: a missing library or a non-sm_100 device raises.
"""
import torch

device = torch.device("cuda") if torch.cuda.is_available() else "cpu"

from . import _lib  # noqa: E402
from . import _rng  # noqa: E402
from . import nets  # noqa: E402
from . import popart  # noqa: E402
from . import adv_estimator  # noqa: E402
from . import agent  # noqa: E402
from .agent import Agent  # noqa: E402
from . import replay  # noqa: E402
from . import augmentations  # noqa: E402
from . import learning_utils  # noqa: E402
from . import learning  # noqa: E402

manual_seed = _rng.manual_seed

MLP_IMPLS = {"ffma": 1, "tcgen05": 2}


def set_mlp_impl(name):
    """Select how the ensemble MLP GEMMs run: "tcgen05" (default; 5th-gen tensor cores, 3xTF32 operand splitting,
    fp32 accumulation in TMEM) or "ffma" (exact-fp32 CUDA-core tiles, the in-library cross-check)."""
    _lib.lib().set_default_mlp_impl(MLP_IMPLS[name])


def get_mlp_impl():
    cur = _lib.lib().default_mlp_impl()
    return next(k for k, v in MLP_IMPLS.items() if v == cur)


def set_overlap(on):
    """Run the independent branches of an update on a second stream (weight-gradient GEMMs next to the data-gradient
    GEMMs, the online critics' hidden layers next to the target networks).  On by default; results do not change."""
    _lib.lib().set_overlap(int(bool(on)))


def set_fused_forward(on):
    """Use the single-kernel forward (three layers + head in one launch) for 2 x 256-class networks.  On by default;
    off = one launch per layer (the cross-check path)."""
    _lib.lib().set_fused_forward(int(bool(on)))


def set_pdl(on):
    """Programmatic dependent launch along the update's critical path (each kernel's set-up overlaps the tail of its
    predecessor).  On by default; results do not change."""
    _lib.lib().set_pdl(int(bool(on)))


__all__ = ["Agent", "agent", "nets", "replay", "learning", "learning_utils", "augmentations", "popart",
           "adv_estimator", "device", "manual_seed", "set_mlp_impl", "get_mlp_impl", "set_overlap", "set_fused_forward", "set_pdl"]
