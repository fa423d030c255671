#!/usr/bin/env python
"""Warm, graph-replayed timings of individual C-ABI entry points at BASELINE shapes (CUDA events)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from super_sac_b200 import _ops, _lib
from super_sac_b200._arena import MLPArena
DEV = "cuda"

def timeit(fn, reps=20, iters=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * iters)

def mlp(G, D, H, O, B, impl):
    ar = MLPArena(G, D, H, O, DEV); ar.flat.normal_(0, 0.05)
    x = torch.randn(B, D, device=DEV); h1 = torch.empty(G, B, H, device=DEV); h2 = torch.empty_like(h1)
    y = torch.empty(G, B, O, device=DEV); dy = torch.randn(G, B, O, device=DEV); dx = torch.empty(G, B, D, device=DEV)
    f = timeit(lambda: _ops.mlp_forward(ar, 0, G, x, B, h1, h2, y, impl=impl))
    b = timeit(lambda: _ops.mlp_backward(ar, 0, G, x, B, h1, h2, dy, want_dw=True, impl=impl))
    bx = timeit(lambda: _ops.mlp_backward(ar, 0, G, x, B, h1, h2, dy, want_dw=False, dx=dx, lddx=D, impl=impl))
    return f, b, bx

if __name__ == "__main__":
    for name, shp in [("critic C2 G=10", (10, 23, 256, 1, 256)), ("target C2 G=2", (2, 23, 256, 1, 256)), ("actor C2 G=1", (1, 17, 256, 12, 256)),
                      ("critic C5 G=2 H=1024 B=1024", (2, 23, 1024, 1, 1024))]:
        for impl in (1, 2):
            f, b, bx = mlp(*shp, impl)
            print(f"{name:30s} impl={impl}: fwd {f:7.1f} us   bwd(dW) {b:7.1f} us   bwd(dx only) {bx:7.1f} us")
    n = 721930
    p = torch.randn(n, device=DEV); t = torch.randn(n, device=DEV); g = torch.randn(n, device=DEV); m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV)
    ctl = torch.zeros(2, dtype=torch.int32, device=DEV)
    L = _lib.lib()
    print(f"polyak C2 (8.7 MB)   {timeit(lambda: L.polyak(t.data_ptr(), p.data_ptr(), n, 0.005, _lib.stream_ptr())):6.1f} us")
    print(f"adam   C2 (20 MB)    {timeit(lambda: L.adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, ctl.data_ptr(), 3e-4, .9, .999, 1e-8, 0., None, 0., 0, _lib.stream_ptr())):6.1f} us")
    # split critic backward at the REDQ shape: u GEMM (_pre) and the two branches of _post
    G, D, H, B = 10, 23, 256, 256
    ar = MLPArena(G, D, H, 1, DEV); ar.flat.normal_(0, 0.05)
    x = torch.randn(B, D, device=DEV); h1 = torch.rand(G, B, H, device=DEV); h2 = torch.rand(G, B, H, device=DEV)
    dq = torch.randn(G, B, 1, device=DEV) / B
    ws = torch.empty(L.mlp_backward_ws(G, B, H), dtype=torch.float32, device=DEV)
    W1, _, W2, _, W3, _ = ar.ptrs(0)
    gr = ar.ptrs(0, grad=True)
    pre = lambda a: L.mlp_backward_pre(W2, W3, G, H, B, h1.data_ptr(), h2.data_ptr(), ws.data_ptr(), a, 2, _lib.stream_ptr())
    post = lambda: L.mlp_backward_post(W3, G, D, H, x.data_ptr(), D, 0, B, h1.data_ptr(), h2.data_ptr(), dq.data_ptr(), ws.data_ptr(), *gr, 2, _lib.stream_ptr())
    print(f"split bwd C2: _pre (u GEMM, gated A)            {timeit(lambda: pre(0)):6.1f} us")
    print(f"split bwd C2: _post, branches side by side      {timeit(post):6.1f} us")
    L.set_overlap(0)
    print(f"split bwd C2: _post, branches in series         {timeit(post):6.1f} us   (gW1 reduction = series - gW2 GEMM)")
    L.set_overlap(1)
