"""Diagnostic: per-tensor errors of the native encoder against the oracle at the BASELINE geometry."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import encoder_oracle as eo  # noqa: E402
import test_conv_encoder as tce  # noqa: E402
from super_sac_b200.nets import cnns  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rng = np.random.default_rng(5)
torch.manual_seed(5)
C, H, W, O = 9, 84, 84, 50
enc = cnns.BigPixelEncoder((C, H, W), out_dim=O)
with torch.no_grad():
    for p in enc.parameters():
        p.add_(0.02 * torch.randn_like(p))
params = {k: v.detach().numpy() for k, v in enc.named_parameters()}
obs = rng.integers(0, 256, (B, C, H, W)).astype(np.float32)
dout = rng.standard_normal((B, O)).astype(np.float32)
ref_out, cache = eo.forward(params, obs)
g = eo.backward(cache, dout)
nat = tce._Native(params, B, C, H, W, O)
out = nat.forward(obs)
grads = nat.backward(dout)
for n in eo.PARAM_NAMES:
    w = g[n].numpy().astype(np.float64)
    e = np.abs(grads[n] - w)
    print(f"{n:14s} max|want| {np.abs(w).max():.4e}  max err {e.max():.3e}  rel-to-max {e.max()/np.abs(w).max():.3e}")
for l in (4, 3, 2, 1):
    got, full, _ = nat.dz(l)
    w = g[f"dz{l}"].numpy().astype(np.float64)
    e = np.abs(got - w)
    tol = 1e-4 * np.abs(w).max() + 1e-4 * np.abs(w)
    print(f"dz{l}: mismatches {(e > tol).sum()} of {e.size}; max err {e.max():.3e} vs max {np.abs(w).max():.3e}")
    cs = full.reshape(-1, 32).double().sum(0).cpu().numpy()
    name = f"conv{l}.bias"
    print(f"  colsum(torch over the dz buffer) vs kernel {name}: {np.abs(cs - grads[name]).max():.3e}; vs oracle {np.abs(cs - g[name].numpy()).max():.3e}")
