#!/usr/bin/env python
"""Kernel-level measurements next to bench.py (which carries every BASELINE config: --config, and "secondary").

    python tools/bench_configs.py [--only mlp,kernels,<config names>] [--steps K]

  mlp      ensemble MLP forward / backward at the C2..C5 shapes: algorithmic TFLOP/s against the measured bf16 peak and the
           3xTF32 ceiling (bf16 / 6)
  kernels  achieved GB/s of gather+augment, Polyak, Adam against the measured HBM peak
  sac|redq|sunrise|drqv2|afbc   device-resident ms/step of one config (same code path as bench.py's secondary block)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tools")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import torch  # noqa: E402

import bench  # noqa: E402
import benchlib as bl  # noqa: E402
import super_sac_b200 as ssb  # noqa: E402

DEV = torch.device("cuda", 0)


def mlp_kernels():
    """Tensor-core ensemble MLP at the shapes of the BASELINE configs: algorithmic TFLOP/s of the forward and of the
    backward (data + weight gradients) against the measured bf16 peak and against the 3xTF32 ceiling (bf16 / 6)."""
    from super_sac_b200 import _arena, _ops

    peaks = bl.measured_peaks()
    for name, G, D, H, B in (("C2 critics 10 x (23-256-256-1), B=256", 10, 23, 256, 256),
                             ("C3 critics 10 x (23-256-256-1), B=256 per member", 2, 23, 256, 256),
                             ("C5 critics 2 x (23-1024-1024-1), B=1024", 2, 23, 1024, 1024),
                             ("C4 critics 2 x (56-1024-1024-1), B=512", 2, 56, 1024, 512)):
        ar = _arena.MLPArena(G, D, H, 1, DEV)
        for n in _arena.NAMES:
            ar.p[n].normal_(0, 0.05)
        x = torch.randn(B, D, device=DEV)
        h1 = torch.empty(G, B, H, device=DEV); h2 = torch.empty_like(h1); y = torch.empty(G, B, 1, device=DEV)
        dy = torch.randn(G, B, 1, device=DEV) / B
        fwd_fl = 2.0 * G * B * (D * H + H * H + H)
        bwd_fl = 2.0 * G * B * (2 * H + 2 * H * H + D * H)

        def graph_time(fn, per=10, iters=20):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(per):
                    fn()
            g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / (per * iters)

        f_ms = graph_time(lambda: _ops.mlp_forward(ar, 0, G, x, B, h1, h2, y, keep_hidden=True))
        b_ms = graph_time(lambda: _ops.mlp_backward(ar, 0, G, x, B, h1, h2, dy, want_dw=True))
        for what, fl, ms in (("forward", fwd_fl, f_ms), ("backward (data + weight gradients)", bwd_fl, b_ms)):
            tf = fl / ms / 1e9
            print(json.dumps({"kernel": f"ensemble MLP {what}, {name}", "algorithmic_GFLOP": fl / 1e9, "us": ms * 1e3, "TFLOPs": tf,
                              "frac_of_bf16_peak": tf / peaks["bf16_tflops"], "frac_of_3xTF32_ceiling": tf / (peaks["bf16_tflops"] / 6),
                              "peak_bf16_TFLOPs": peaks["bf16_tflops"]}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="kernels,mlp,sac,sunrise,drqv2,afbc")
    ap.add_argument("--steps", type=int, default=300)
    args = ap.parse_args()
    which = args.only.split(",")
    torch.cuda.set_device(DEV)
    if "kernels" in which:
        for row in bench.hbm_kernels():
            print(json.dumps(row), flush=True)
    if "mlp" in which:
        mlp_kernels()
    for name in ("sac", "redq", "sunrise", "drqv2", "afbc"):
        if name in which:
            W = bl.Workload(ssb, name, DEV, seed=0, fill_on_device=True)
            ms, mode, _ = bench.time_config(W, min(args.steps, {"drqv2": 30, "afbc": 100}.get(name, args.steps)), 5)
            print(json.dumps({"config": name, "workload": W.cfg["workload"], "ms_per_step": ms, "updates_per_s": 1e3 / ms, "mode": mode}), flush=True)
            del W
            torch.cuda.empty_cache()
