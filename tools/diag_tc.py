import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
from oracle import update_oracle as uo
from super_sac_b200 import _ops
from super_sac_b200._arena import MLPArena
DEV = "cuda"
def run(G, D, H, O, B, impl):
    gen = torch.Generator().manual_seed(1)
    st = uo.MLPStack(G, D, H, O).random_init(gen)
    ar = MLPArena(G, D, H, O, DEV)
    for n in uo.PARAM_NAMES: ar.p[n].copy_(getattr(st, n).to(DEV))
    x = torch.randn(B, D, generator=gen); dy = torch.randn(G, B, O, generator=gen) / B
    xd = x.to(DEV); h1 = torch.empty((G, B, H), device=DEV); h2 = torch.empty_like(h1); y = torch.empty((G, B, O), device=DEV)
    _ops.mlp_forward(ar, 0, G, xd, B, h1, h2, y, impl=impl)
    grads = st.zeros_like(); dxw = torch.zeros(G, B, D)
    x64 = x.double()
    for g in range(G):
        yw, h1w, h2w = uo.mlp_forward(st, g, x)
        dxw[g] = uo.mlp_backward(st, g, x, h1w, h2w, dy[g], grads, need_dx=True)
        if g == 0:
            h1_64 = torch.relu(x64 @ st.W1[g].double().t() + st.b1[g].double()); h2_64 = torch.relu(h1_64 @ st.W2[g].double().t() + st.b2[g].double())
            print(f"  h2 err vs f64: gpu {float((h2[0].cpu().double()-h2_64).abs().max()):.3e}  cpu-fp32 {float((h2w.double()-h2_64).abs().max()):.3e}  scale {float(h2_64.abs().max()):.3f}")
    dx = torch.empty((G, B, D), device=DEV)
    _ops.mlp_backward(ar, 0, G, xd, B, h1, h2, dy.to(DEV), want_dw=True, dx=dx, lddx=D, impl=impl)
    torch.cuda.synchronize()
    for n in uo.PARAM_NAMES:
        got, want = ar.g[n].cpu(), getattr(grads, n)
        print(f"  {n}: max|err| {float((got-want).abs().max()):.3e} scale {float(want.abs().max()):.3e} zeros_got {int((got==0).sum())}/{got.numel()}")
    print(f"  dx: max|err| {float((dx.cpu()-dxw).abs().max()):.3e} scale {float(dxw.abs().max()):.3e}")
for cfg in [(2, 23, 256, 1, 256), (1, 24, 128, 1, 128), (1, 32, 128, 1, 128)]:
    for impl in (1, 2):
        print(cfg, "impl", impl); run(*cfg, impl)
