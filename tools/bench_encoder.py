"""Times the native DrQ encoder (csrc/ssac_conv.cu) at BASELINE config 4 (B = 512, 9 x 84 x 84 -> 50) with CUDA events:
forward with saved activations, backward, no-grad forward, and -- for comparison, clearly labelled -- the same module through
PyTorch / cuDNN fp32 (TF32 off, the reference's arithmetic).  One JSON line per figure.

    python tools/bench_encoder.py [--batch 512] [--iters 10]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from super_sac_b200.nets import cnns  # noqa: E402

FLOP_PER_SAMPLE = 2 * (41 * 41 * 32 * 81 + (39 * 39 + 37 * 37 + 35 * 35) * 32 * 288 + 39200 * 50)   # forward, 88.5 MFLOP


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B = a.batch
    enc = cnns.BigPixelEncoder((9, 84, 84), 50).cuda()
    obs = torch.randint(0, 256, (B, 9, 84, 84), device="cuda").float()
    dout = torch.randn(B, 50, device="cuda")
    flop = FLOP_PER_SAMPLE * B

    def fwd_bwd():
        enc.zero_grad(set_to_none=True)
        y = enc(obs)
        y.backward(dout)

    def fwd_nograd():
        with torch.no_grad():
            enc(obs)

    def fwd_grad():
        enc(obs)   # forward with saved activations; the graph is dropped

    res = {}
    for impl in ("native", "torch"):
        os.environ["SSAC_ENCODER_IMPL"] = impl
        res[impl] = dict(fwd_nograd_us=timed(fwd_nograd, a.iters), fwd_bwd_us=timed(fwd_bwd, a.iters))
        r = res[impl]
        r["fwd_tflops"] = flop / r["fwd_nograd_us"] / 1e6
        r["fwd_bwd_tflops"] = 3 * flop / r["fwd_bwd_us"] / 1e6
        print(json.dumps({"impl": impl + (" (PyTorch/cuDNN fp32, comparison only)" if impl == "torch" else ""), "batch": B, **r}))
    os.environ["SSAC_ENCODER_IMPL"] = "native"
    print(json.dumps({"speedup_fwd": res["torch"]["fwd_nograd_us"] / res["native"]["fwd_nograd_us"],
                      "speedup_fwd_bwd": res["torch"]["fwd_bwd_us"] / res["native"]["fwd_bwd_us"]}))


if __name__ == "__main__":
    main()
