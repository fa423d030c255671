"""Import the UNMODIFIED reference package (jakegrigsby/super_sac) for the timed reference arm and the drop-in proof.

``baseline/_ref/super_sac`` is a verbatim copy of ``/root/reference/super_sac`` made by ``__graft_entry__.build()`` in
the build container (git-ignored: the history stays source-only; it travels to the GPU box with the snapshot, where
``/root/reference`` does not exist).  ``pip install --target baseline/_ref /root/reference`` is not possible here: the
reference's setup.py asks for ``setup_requires=["pytest-runner"]``, which the offline wheelhouse does not carry; the
package is pure Python, so the copy is the install.

The reference's ``import super_sac`` pulls in gin / gymnasium / gym / tensorboardX / skimage, none of which are
installed; five stub modules make the import succeed and leave the update path (learning.py, learning_utils.py,
agent.py, replay.py, augmentations.py, popart.py, nets/*) running unchanged.
Nothing under super_sac_b200/ imports this file.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")


def available(root=None):
    return os.path.isdir(os.path.join(root or REF_ROOT, "super_sac"))


def import_reference(root=None, device="cpu"):
    """Returns the reference ``super_sac`` package with its module-level ``device`` forced to ``device``."""
    root = root or REF_ROOT
    cur = sys.modules.get("super_sac")
    if cur is None or not getattr(cur, "_ssac_ref", False):

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
        sys.path.insert(0, root)
        try:
            import super_sac  # noqa: F401
        finally:
            sys.path.remove(root)
        cur = sys.modules["super_sac"]
        cur._ssac_ref = True
    import torch

    dev = torch.device(device) if device != "cpu" else "cpu"
    # every module did ``from . import device`` at import time (learning.py:13, learning_utils.py:10, agent.py:10,
    # main.py:18, ...): re-point all of them
    for name, mod in list(sys.modules.items()):
        if (name == "super_sac" or name.startswith("super_sac.")) and mod is not None and hasattr(mod, "device"):
            mod.device = dev
    return cur
