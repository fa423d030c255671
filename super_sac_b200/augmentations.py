"""DrQ / DrQv2 / identity augmentations (reference augmentations.py:20-41, :165-293, :489-503).

In the update path these objects only carry *parameters* (pad size, noise flag, the per-sample shifts drawn once per
call and shared by o and o1): ``learning_utils.sample_move_and_augment`` fuses gather + shift + uint8->fp32 + aug_mix
into ssac_gather_aug_u8, reading the uint8 frames straight from the device replay ring.  ``__call__`` on a float
batch (the reference's standalone use) is kept as an index-gather with the same integer-crop semantics.

Drqv2Aug note: the reference evaluates the shift with a bilinear ``grid_sample`` whose fp32 grid lands
~1e-5 px off the pixel centres, so its output differs from the integer crop it encodes by <= 4e-3 on the 0..255
scale.  This implementation is the exact integer crop (replicate padding); DrqAug / DrqNoNoiseAug (reflection
padding + integer crop) are bit-identical to the reference.
"""
import torch

from . import _rng

PAD_NONE, PAD_REPLICATE, PAD_REFLECT, PAD_RAD = 0, 1, 2, 3


class _ShiftAug:
    pad_mode = PAD_NONE
    shift_range_extra = 0  # v2 draws from [0, 2*pad] (inclusive), v1 from [0, 2*pad)

    def __init__(self, batch_size, pad=4, noise=False, *_args, **_kwargs):
        self.batch_size = batch_size
        self.pad = pad
        self.noise = noise
        self.shift = None  # int32 [B,2] = (x, y) on the device, drawn by change_randomization_params

    def change_randomization_params(self, device=None):
        if device is None:
            from . import device as default_device

            device = default_device
        if self.shift is None or self.shift.device != torch.device(device):
            self.shift = torch.zeros((self.batch_size, 2), dtype=torch.int32, device=device)
        _rng.source().shifts(self.shift, 2 * self.pad + self.shift_range_extra)

    def _coords(self, n, shift, device):
        p = torch.arange(n, device=device)[None, :] + shift[:, None].long() - self.pad
        if self.pad_mode == PAD_REFLECT:
            p = torch.where(p < 0, -p, p)
            p = torch.where(p >= n, 2 * (n - 1) - p, p)
            return p
        return p.clamp(0, n - 1)

    def __call__(self, imgs):
        b, c, h, w = imgs.shape
        assert b == self.batch_size
        if self.shift is None or self.shift.device != imgs.device:
            self.change_randomization_params(imgs.device)
        ys = self._coords(h, self.shift[:, 1], imgs.device)
        xs = self._coords(w, self.shift[:, 0], imgs.device)
        out = imgs[torch.arange(b, device=imgs.device)[:, None, None, None], torch.arange(c, device=imgs.device)[None, :, None, None],
                   ys[:, None, :, None], xs[:, None, None, :]].float()
        if self.noise:
            out = out + torch.randn_like(out)
        return out.clamp(0, 255.0)


class DrqAug(_ShiftAug):
    """Reflection pad + random integer crop (+ N(0,1) noise)."""
    pad_mode = PAD_REFLECT

    def __init__(self, batch_size, pad=4, noise=True, *_args, **kwargs):
        super().__init__(batch_size, pad, noise)

    def __repr__(self):
        return "Drqv1"


class DrqNoNoiseAug(DrqAug):
    def __init__(self, batch_size, pad=4, noise=False, *_args, **kwargs):
        super().__init__(batch_size, pad, noise)

    def __repr__(self):
        return "Drqv1NoNoise"


class LargeDrqNoNoiseAug(DrqAug):
    def __init__(self, batch_size, pad=12, noise=False, *_args, **kwargs):
        super().__init__(batch_size, pad, noise)


class LargeDrqAug(DrqAug):
    def __init__(self, batch_size, pad=12, *_args, **kwargs):
        super().__init__(batch_size, pad)


class Drqv2Aug(_ShiftAug):
    """Replicate pad + random shift in [0, 2*pad]^2 (integer-crop restatement, see module docstring)."""
    pad_mode = PAD_REPLICATE
    shift_range_extra = 1

    def __init__(self, batch_size, pad=4, noise=False, *_args, **kwargs):
        super().__init__(batch_size, pad, noise)

    def __repr__(self):
        return "DrqV2"


class RadAug(_ShiftAug):
    """RAD (reference augmentations.py:129-162): bilinear upscale by ``crop`` pixels per axis (cv2.resize,
    INTER_LINEAR, float32) followed by a random H x W window.  In the update path only the window is evaluated, inside
    the gather kernel and straight from the uint8 ring (ssac_gather_aug_u8, pad_mode 3), with cv2's separable fp32
    arithmetic: bit-exact with the reference on frame stacks (> 4 channels); cv2's <= 4-channel path rounds
    differently (<= 1e-3 on the 0..255 scale).  ``shift`` = (w, h) window offsets in [0, crop)."""
    pad_mode = PAD_RAD

    def __init__(self, batch_size, crop=16, *_args, **_kwargs):
        super().__init__(batch_size, pad=crop, noise=False)
        self.crop = crop

    def change_randomization_params(self, device=None):
        if device is None:
            from . import device as default_device

            device = default_device
        if self.shift is None or self.shift.device != torch.device(device):
            self.shift = torch.zeros((self.batch_size, 2), dtype=torch.int32, device=device)
        _rng.source().shifts(self.shift, self.crop)

    def __call__(self, imgs):
        """Standalone use on a float batch [B,C,H,W] (values 0..255 that are whole numbers, as the update path
        produces them): routed through the same kernel via a uint8 view of the batch."""
        from . import _lib

        b, c, h, w = imgs.shape
        assert b == self.batch_size
        if self.shift is None or self.shift.device != imgs.device:
            self.change_randomization_params(imgs.device)
        src = imgs.round().clamp(0, 255).to(torch.uint8).contiguous()
        if not torch.equal(src.float(), imgs):
            raise NotImplementedError("RadAug on non-integer pixel values (the fused path reads the uint8 replay ring)")
        out = torch.empty((b, c, h, w), dtype=torch.float32, device=imgs.device)
        idx = torch.arange(b, dtype=torch.int64, device=imgs.device)
        _lib.lib().gather_aug_u8(src.data_ptr(), out.data_ptr(), idx.data_ptr(), self.shift.data_ptr(), None, b, c, h, w,
                                 self.crop, PAD_RAD, b, _lib.stream_ptr())
        return out

    def __repr__(self):
        return "RAD"


class IdentityAug:
    pad_mode = PAD_NONE
    noise = False
    pad = 0

    def __init__(self, batch_size, *_args, **_kwargs):
        self.batch_size = batch_size
        self.shift = None

    def __call__(self, imgs):
        return imgs

    def change_randomization_params(self, device=None):
        return

    def __repr__(self):
        return "Identity"


class AugmentationSequence:
    def __init__(self, aug_list, keys=None):
        self.aug_list = aug_list
        self.keys = keys

    def fusable(self):
        """The single-aug sequences every shipped config uses run inside the gather kernel."""
        return len(self.aug_list) == 1 and isinstance(self.aug_list[0], (_ShiftAug, IdentityAug))

    def __call__(self, *batches):
        if self.keys is None:
            self.keys = batches[0].keys()
        for aug in self.aug_list:
            aug.change_randomization_params()
        results = []
        for original in batches:
            batch = {k: v.clone() for k, v in original.items()}
            for key in self.keys:
                for aug in self.aug_list:
                    with torch.no_grad():
                        batch[key] = aug(batch[key])
            results.append(batch)
        return tuple(results) if len(results) > 1 else results[0]

    def __repr__(self):
        return f"AugmentationSequence: ({[repr(a) for a in self.aug_list]})"
