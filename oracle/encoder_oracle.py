"""CPU restatement of the DrQ pixel encoder -- TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product (super_sac_b200) never
does.  It restates ``BigPixelEncoder`` (reference nets/cnns.py:37-69) with an EXPLICIT backward, tap by tap, in float32
torch-CPU matmuls, so that every intermediate the CUDA kernels produce (activations, per-layer gradients) has a checker:

    img = obs / 255 - 0.5                                   cnns.py:58
    x   = relu(conv3x3 stride 2 (C -> 32))                  cnns.py:40,59
    x   = relu(conv3x3 stride 1 (32 -> 32)) three times     cnns.py:41-43,60-62
    x   = fc(x.view(B, -1))   (NCHW flatten)                cnns.py:52,63-64
    out = tanh(LayerNorm(x))                                cnns.py:53,65-66

Parity is PINNED: tests/golden/encoder.npz holds outputs and autograd gradients of the UNMODIFIED reference module
(tests/golden/make_golden.py run_encoder_case) and tests/test_oracle_golden.py checks this file against them.
"""
import numpy as np
import torch

PARAM_NAMES = ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "conv3.weight", "conv3.bias", "conv4.weight",
               "conv4.bias", "fc.weight", "fc.bias", "ln.weight", "ln.bias"]


def _t(x):
    return torch.as_tensor(np.asarray(x), dtype=torch.float32)


def conv3x3_forward(x, w, b, stride):
    """x [B,Ci,H,W], w [Co,Ci,3,3] -> [B,Co,Ho,Wo]: nn.Conv2d(kernel 3, no padding) as nine tap matmuls (cnns.py:40-43)."""
    B, Ci, H, W = x.shape
    Ho, Wo = (H - 3) // stride + 1, (W - 3) // stride + 1
    out = b.view(1, -1, 1, 1).expand(B, w.shape[0], Ho, Wo).clone()
    for kh in range(3):
        for kw in range(3):
            xs = x[:, :, kh:kh + stride * (Ho - 1) + 1:stride, kw:kw + stride * (Wo - 1) + 1:stride]   # [B,Ci,Ho,Wo]
            out += torch.einsum("bchw,oc->bohw", xs, w[:, :, kh, kw])
    return out


def conv3x3_backward(x, w, dz, stride, need_dx=True):
    """Gradients of conv3x3_forward given dz = dL/d(out): (dx, dw, db)."""
    B, Ci, H, W = x.shape
    Ho, Wo = dz.shape[2], dz.shape[3]
    dw = torch.zeros_like(w)
    dx = torch.zeros_like(x) if need_dx else None
    for kh in range(3):
        for kw in range(3):
            sl = (slice(None), slice(None), slice(kh, kh + stride * (Ho - 1) + 1, stride),
                  slice(kw, kw + stride * (Wo - 1) + 1, stride))
            dw[:, :, kh, kw] = torch.einsum("bohw,bchw->oc", dz, x[sl])
            if need_dx:
                dx[sl] += torch.einsum("bohw,oc->bchw", dz, w[:, :, kh, kw])
    return dx, dw, dz.sum(dim=(0, 2, 3))


def forward(params, obs):
    """params: dict name -> array (PARAM_NAMES); obs [B,C,H,W] in 0..255.  Returns (out [B,O], cache)."""
    p = {k: _t(v) for k, v in params.items()}
    x0 = _t(obs) / 255.0 - 0.5
    acts = [x0]
    x = x0
    for l, stride in ((1, 2), (2, 1), (3, 1), (4, 1)):
        x = torch.relu(conv3x3_forward(x, p[f"conv{l}.weight"], p[f"conv{l}.bias"], stride))
        acts.append(x)
    flat = x.reshape(x.shape[0], -1)
    z = flat @ p["fc.weight"].t() + p["fc.bias"]
    mean = z.mean(dim=1, keepdim=True)
    var = ((z - mean) ** 2).mean(dim=1, keepdim=True)          # LayerNorm: biased variance, eps 1e-5
    rstd = 1.0 / torch.sqrt(var + 1e-5)
    xhat = (z - mean) * rstd
    out = torch.tanh(xhat * p["ln.weight"] + p["ln.bias"])
    return out, dict(p=p, acts=acts, flat=flat, xhat=xhat, rstd=rstd, out=out)


def backward(cache, dout, masks=None):
    """Explicit backward of ``forward``: dict name -> gradient (+ "dz1".."dz4": dL/d(pre-activation) per conv layer).
    masks (optional): {layer: bool [B,32,h,w]} ReLU patterns to differentiate through instead of the cache's own -- a
    pre-activation within rounding of zero lands on either side of the ReLU in two correct fp32 forwards, and a parity
    test of the backward has to run both sides on the same pattern."""
    p, acts, xhat, rstd, out = cache["p"], cache["acts"], cache["xhat"], cache["rstd"], cache["out"]
    dout = _t(dout)
    g = {}
    dl = dout * (1.0 - out * out)
    g["ln.weight"] = (dl * xhat).sum(0)
    g["ln.bias"] = dl.sum(0)
    dxh = dl * p["ln.weight"]
    dz = rstd * (dxh - dxh.mean(dim=1, keepdim=True) - xhat * (dxh * xhat).mean(dim=1, keepdim=True))
    g["fc.weight"] = dz.t() @ cache["flat"]
    g["fc.bias"] = dz.sum(0)
    dx = (dz @ p["fc.weight"]).reshape(acts[4].shape)
    for l, stride in ((4, 1), (3, 1), (2, 1), (1, 2)):
        dzl = dx * (_t(masks[l]) if masks is not None else (acts[l] > 0).float())
        g[f"dz{l}"] = dzl
        dx, dw, db = conv3x3_backward(acts[l - 1], p[f"conv{l}.weight"], dzl, stride, need_dx=l > 1)
        g[f"conv{l}.weight"], g[f"conv{l}.bias"] = dw, db
    return g
