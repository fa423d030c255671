"""CPU restatement of the DrQ / DrQv2 augmentations + aug_mix blend (numpy).  TEST INFRASTRUCTURE ONLY.

augmentations.py:165-211 (DrqAug: reflection pad + integer crop [+ N(0,1) noise] + clamp),
augmentations.py:214-269 (Drqv2Aug: replicate pad + shift; restated as the integer crop it encodes --
the reference evaluates it with a bilinear grid_sample whose fp32 grid arithmetic lands ~1e-5 px off the
pixel centres, so it differs from the integer crop by <= 4e-3 on the 0..255 scale),
augmentations.py:129-162 (RadAug: cv2 bilinear upscale + integer crop),
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


def _cv2_linear_axis(n_src, n_dst, horizontal):
    """Source taps and fp32 weights of cv2.resize(INTER_LINEAR) along one axis (OpenCV resize.cpp, resizeGeneric_ for
    float32 images): fx = float((d + 0.5) * scale - 0.5) with scale = 1 / (n_dst / n_src) in double, tap = floor(fx),
    weight = fx - tap in fp32.  Horizontally an out-of-range tap is moved onto the edge with weight 0; vertically the
    two row indices are clamped and keep their weights."""
    scale = 1.0 / (n_dst / n_src)
    f = ((np.arange(n_dst, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if horizontal:
        lo, hi = s < 0, s >= n_src - 1
        f[lo], s[lo] = 0.0, 0
        f[hi], s[hi] = 0.0, n_src - 1
        s0, s1 = s, np.minimum(s + 1, n_src - 1)
    else:
        s0, s1 = np.clip(s, 0, n_src - 1), np.clip(s + 1, 0, n_src - 1)
    return s0, s1, (np.float32(1.0) - f).astype(np.float32), f


def rad_crop(imgs, h_off, w_off, crop=16):
    """RadAug.__call__ (augmentations.py:129-162): every image is upscaled to (H+crop, W+crop) with
    cv2.resize(INTER_LINEAR) on float32 HWC data and the window [h:h+H, w:w+W] is cut out.  imgs [B,C,H,W] (any dtype)
    -> float32.  Restates cv2's separable fp32 arithmetic -- horizontal pass S[x0]*a0 + S[x1]*a1, vertical pass
    r0*b0 + r1*b1, each product and sum rounded to fp32 -- which reproduces cv2 4.x bit for bit on images with more
    than 4 channels (frame stacks); cv2's <=4-channel path rounds differently (<= 1e-3 on the 0..255 scale)."""
    imgs = np.asarray(imgs)
    B, C, H, W = imgs.shape
    S = imgs.astype(np.float32)
    ys0, ys1, b0, b1 = _cv2_linear_axis(H, H + crop, False)
    xs0, xs1, a0, a1 = _cv2_linear_axis(W, W + crop, True)
    out = np.empty((B, C, H, W), dtype=np.float32)
    for i in range(B):
        Y, X = np.arange(H) + int(h_off[i]), np.arange(W) + int(w_off[i])
        y0, y1, x0, x1 = ys0[Y], ys1[Y], xs0[X], xs1[X]
        wa0, wa1, wb0, wb1 = a0[X], a1[X], b0[Y][None, :, None], b1[Y][None, :, None]
        r0 = ((S[i][:, y0][:, :, x0] * wa0).astype(np.float32) + (S[i][:, y0][:, :, x1] * wa1).astype(np.float32)).astype(np.float32)
        r1 = ((S[i][:, y1][:, :, x0] * wa0).astype(np.float32) + (S[i][:, y1][:, :, x1] * wa1).astype(np.float32)).astype(np.float32)
        out[i] = ((r0 * wb0).astype(np.float32) + (r1 * wb1).astype(np.float32)).astype(np.float32)
    return out


def mix(original_f32, augmented_f32, aug_mix):
    """learning_utils.py:200-206."""
    B = original_f32.shape[0]
    k = int(B * aug_mix)
    out = original_f32.copy()
    out[:k] = augmented_f32[:k]
    return out
