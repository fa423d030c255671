"""CPU restatement of the DrQ / DrQv2 augmentations + aug_mix blend (numpy).  TEST INFRASTRUCTURE ONLY.

augmentations.py:165-211 (DrqAug: reflection pad + integer crop [+ N(0,1) noise] + clamp),
augmentations.py:214-269 (Drqv2Aug: replicate pad + shift; restated as the integer crop it encodes --
the reference evaluates it with a bilinear grid_sample whose fp32 grid arithmetic lands ~1e-5 px off the
pixel centres, so it differs from the integer crop by <= 4e-3 on the 0..255 scale),
learning_utils.py:193-206 (uint8 -> float cast, first int(B*aug_mix) rows replaced by their augmentation).
Byte/index work: the CUDA path must match this bit-for-bit.
"""
import numpy as np


def _reflect(p, n):
    """nn.ReflectionPad2d index map (no edge repeat): padded coord (already minus pad) -> source coord."""
    p = np.where(p < 0, -p, p)
    return np.where(p >= n, 2 * (n - 1) - p, p)


def drq_v1_crop(imgs, w1, h1, pad=4, noise=None):
    """imgs [B,C,H,W] any dtype -> float32.  w1,h1 in [0, 2*pad): augmentations.py:179-195."""
    imgs = np.asarray(imgs)
    B, C, H, W = imgs.shape
    ys = _reflect(np.arange(H)[None, :] + np.asarray(h1)[:, None] - pad, H)  # [B,H]
    xs = _reflect(np.arange(W)[None, :] + np.asarray(w1)[:, None] - pad, W)  # [B,W]
    out = imgs[np.arange(B)[:, None, None, None], np.arange(C)[None, :, None, None],
               ys[:, None, :, None], xs[:, None, None, :]].astype(np.float32)
    if noise is not None:
        out = out + np.asarray(noise, dtype=np.float32)
    return np.clip(out, 0.0, 255.0).astype(np.float32)


def drq_v2_crop(imgs, shift, pad=4):
    """Integer-crop restatement of Drqv2Aug.random_crop.  shift [B,1,1,2] or [B,2] in [0, 2*pad];
    shift[...,0] moves along width (grid x), shift[...,1] along height: augmentations.py:226-257."""
    imgs = np.asarray(imgs)
    B, C, H, W = imgs.shape
    shift = np.asarray(shift).reshape(B, 2)
    ys = np.clip(np.arange(H)[None, :] + shift[:, 1:2] - pad, 0, H - 1)
    xs = np.clip(np.arange(W)[None, :] + shift[:, 0:1] - pad, 0, W - 1)
    out = imgs[np.arange(B)[:, None, None, None], np.arange(C)[None, :, None, None],
               ys[:, None, :, None], xs[:, None, None, :]].astype(np.float32)
    return np.clip(out, 0.0, 255.0).astype(np.float32)


def mix(original_f32, augmented_f32, aug_mix):
    """learning_utils.py:200-206."""
    B = original_f32.shape[0]
    k = int(B * aug_mix)
    out = original_f32.copy()
    out[:k] = augmented_f32[:k]
    return out
