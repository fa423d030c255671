#!/usr/bin/env python
"""Secondary measurements for the other BASELINE.json configs and for the HBM-bound kernels (one JSON line each).

    python tools/bench_configs.py [--only sac,sunrise,drqv2,afbc,kernels] [--steps K]

bench.py (the driver contract) stays on the headline REDQ-10 config; this script records, with the same timing hygiene
(CUDA events, warm-up, graph replay where capture is possible), what the other shapes of SURVEY §8 do on a B200:
  sac      C1  SAC, 2 critics, obs 3 / act 1, B=256
  sunrise  C3  SUNRISE, 5 members x 2 critics, weighted Bellman backups (T=20), B=256
  drqv2    C4  DrQv2 pixels: uint8 9x84x84 ring, fused gather+shift, conv encoder (PyTorch/cuDNN plugin), H=1024, B=512
  afbc     C5  offline AFBC: 2 M-transition device ring, B=1024, H=1024, DR3 0.01, clips 40, PER actor update
  kernels      achieved GB/s of gather+augment, Polyak, Adam against the measured HBM peak
"""
import argparse
import copy
import json
import os
import sys
from itertools import chain

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import cuda_util as cu  # noqa: E402
import super_sac_b200 as ssb  # noqa: E402
from super_sac_b200 import _lib, augmentations, graphed, learning, learning_utils as lu, nets  # noqa: E402

DEV = torch.device("cuda", 0)


def timed(step, steps, warmup=10):
    for k in range(warmup):
        step(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        step(k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def graph_or_eager(fn):
    try:
        g = graphed.GraphedCall(fn)
        return (lambda: g.replay()), "cuda-graph replay"
    except Exception as e:  # noqa: BLE001
        torch.cuda.synchronize()
        return fn, f"eager ({type(e).__name__}: {str(e)[:80]})"


def state_config(name, E, N, M, S, A, H, B, steps, weight_type=None, temp=None):
    ssb.manual_seed(0)
    torch.manual_seed(0)
    agent = ssb.Agent(act_space_size=A, encoder=cu.IdentityEncoder(S), actor_network_cls=nets.mlps.ContinuousStochasticActor,
                      critic_network_cls=nets.mlps.ContinuousCritic, ensemble_size=E, num_critics=N, hidden_size=H,
                      auto_rescale_targets=False, log_std_low=-5.0, log_std_high=2.0)
    agent.to(DEV)
    target = copy.deepcopy(agent)
    c_opt, a_opt, e_opt, las, al_opts = cu.optimizers(agent, dict(E=E))
    cfg = dict(S=S, A=A)
    buf = ssb.replay.ReplayBuffer(500_000, device=DEV)
    s, a, r, s1, d = bench.synthetic_transitions(cfg, 500_000)
    buf.load_experience({"obs": s}, a, r, {"obs": s1}, d)
    kw = dict(buffer=buf, agent=agent, target_agent=target, critic_optimizer=c_opt, encoder_optimizer=e_opt, log_alphas=las,
              batch_size=B, gamma=0.99, critic_clip=None, encoder_clip=None, target_critic_ensemble_n=M,
              weighted_bellman_temp=temp, weight_type=weight_type, pop=False,
              augmenter=augmentations.AugmentationSequence([augmentations.IdentityAug(B)]), encoder_lambda=0.0,
              random_process=None, noise_clip=None, aug_mix=0.0)

    def upd():
        out = learning.critic_update(**kw)
        for ac, tc in zip(agent.critics, target.critics):
            lu.soft_update(tc, ac, 0.005)
        return out

    step, mode = graph_or_eager(upd)
    ms = timed(lambda k: step(), steps)

    def full():   # main.py:380-543 with UTD 1: critic update + Polyak + actor update + temperature update
        _, rds = learning._critic_update_impl(**kw)
        for ac, tc in zip(agent.critics, target.critics):
            lu.soft_update(tc, ac, 0.005)
        learning._online_actor_update_impl(buffer=buf, agent=agent, pop=False, actor_optimizer=a_opt, log_alphas=las,
                                           batch_size=B, clip=None, random_process=None, noise_clip=None,
                                           augmenter=kw["augmenter"], aug_mix=0.0, premade_replay_dicts=rds)
        return learning.alpha_update(buffer=buf, agent=agent, optimizers=al_opts, batch_size=B, log_alphas=las,
                                     augmenter=kw["augmenter"], aug_mix=0.0, target_entropy=-float(A),
                                     premade_replay_dicts=rds, discrete=False)

    fstep, fmode = graph_or_eager(full)
    fms = timed(lambda k: fstep(), steps)
    print(json.dumps({"config": name, "metric": "sac_gradient_updates_per_sec", "value": 1e3 / ms, "ms_per_step": ms, "mode": mode,
                      "full_step_utd1": {"ms": fms, "steps_per_sec": 1e3 / fms, "mode": fmode,
                                         "what": "critic_update + Polyak + online_actor_update + alpha_update"},
                      "shape": dict(E=E, N=N, M=M, S=S, A=A, H=H, B=B, weight_type=weight_type)}), flush=True)


def drqv2_config(steps):
    ssb.manual_seed(0)
    torch.manual_seed(0)
    C, HW, A, H, B, cap = 9, 84, 6, 1024, 512, 20_000
    enc = nets.cnns.BigPixelEncoder((C, HW, HW), 50)

    class PixEnc(nets.Encoder):
        def __init__(self):
            super().__init__()
            self.net = enc
            self.embedding_dim_ = 50

        @property
        def embedding_dim(self):
            return 50

        def forward(self, obs):
            return self.net(obs["pixels"])

    agent = ssb.Agent(act_space_size=A, encoder=PixEnc(), actor_network_cls=nets.mlps.ContinuousDeterministicActor,
                      critic_network_cls=nets.mlps.ContinuousCritic, ensemble_size=1, num_critics=2, hidden_size=H,
                      auto_rescale_targets=False)
    agent.to(DEV)
    target = copy.deepcopy(agent)
    c_opt = torch.optim.Adam(chain(*(c.parameters() for c in agent.critics)), lr=1e-4)
    e_opt = torch.optim.Adam(agent.encoder.parameters(), lr=1e-4, capturable=True)
    las = [torch.tensor([-30.0], device=DEV, requires_grad=True)]
    buf = ssb.replay.ReplayBuffer(cap, device=DEV)
    g = torch.Generator(device=DEV).manual_seed(0)
    # fill the device ring directly (12.7 GB of host staging would only measure PCIe)
    buf.load_experience({"pixels": np.zeros((2, C, HW, HW), np.uint8)}, np.zeros((2, A), np.float32), np.zeros(2, np.float32),
                        {"pixels": np.zeros((2, C, HW, HW), np.uint8)}, np.zeros(2, bool))
    st = buf._storage
    st.s_stack["pixels"].random_(0, 256, generator=g)
    st.s1_stack["pixels"].random_(0, 256, generator=g)
    st.action_stack.uniform_(-1, 1, generator=g)
    st.reward_stack.normal_(generator=g)
    st._max_filled, st._next_idx = cap, 0
    buf._n_filled_dev.fill_(cap)
    noise = lu.GaussianExplorationNoise(cu.ActionSpace(A), start_scale=1.0, final_scale=0.1)
    kw = dict(buffer=buf, agent=agent, target_agent=target, critic_optimizer=c_opt, encoder_optimizer=e_opt, log_alphas=las,
              batch_size=B, gamma=0.99**3, critic_clip=None, encoder_clip=None, target_critic_ensemble_n=2,
              weighted_bellman_temp=None, weight_type=None, pop=False,
              augmenter=augmentations.AugmentationSequence([augmentations.Drqv2Aug(B)]), encoder_lambda=0.0, random_process=noise,
              noise_clip=0.3, aug_mix=1.0)

    def upd():
        out = learning.critic_update(**kw)
        lu.soft_update(target.critics[0], agent.critics[0], 0.01)
        lu.soft_update(target.encoder, agent.encoder, 1.0)
        return out

    step, mode = graph_or_eager(upd)
    ms = timed(lambda k: step(), steps, warmup=5)
    # the sampling part alone (fused gather + shift + cast of o and o1): algorithmic 2 x (32.5 MB read + 130 MB written)
    aug = kw["augmenter"]
    ms_s = timed(lambda k: lu.sample_move_and_augment(buf, B, aug, 1.0, per=False), 50, warmup=5)
    bytes_s = 2 * (B * C * HW * HW * 1 + B * C * HW * HW * 4)
    peaks = bench.measured_peaks()
    print(json.dumps({"config": "drqv2 (C4)", "metric": "sac_gradient_updates_per_sec", "value": 1e3 / ms, "ms_per_step": ms, "mode": mode,
                      "shape": dict(obs="u8 9x84x84", B=B, H=H, N=2, encoder="BigPixelEncoder (PyTorch/cuDNN plugin)", aug="Drqv2Aug pad 4"),
                      "sample_move_and_augment": {"ms": ms_s, "algorithmic_MB": bytes_s / 1e6, "GBps": bytes_s / ms_s / 1e6,
                                                  "frac_of_hbm_peak": bytes_s / ms_s / 1e6 / peaks["hbm_gbs"],
                                                  "note": "includes the index / shift draw and the small-array gather launches"}}),
          flush=True)


def afbc_config(steps):
    ssb.manual_seed(0)
    torch.manual_seed(0)
    S, A, H, B, N, cap = 17, 6, 1024, 1024, 2, 2_000_000
    agent = ssb.Agent(act_space_size=A, encoder=cu.IdentityEncoder(S), actor_network_cls=nets.mlps.ContinuousStochasticActor,
                      critic_network_cls=nets.mlps.ContinuousCritic, ensemble_size=1, num_critics=N, hidden_size=H,
                      auto_rescale_targets=False, log_std_low=-5.0, log_std_high=2.0)
    agent.to(DEV)
    target = copy.deepcopy(agent)
    c_opt, a_opt, e_opt, las, _ = cu.optimizers(agent, dict(E=1, init_alpha=1e-15))
    buf = ssb.replay.ReplayBuffer(cap, alpha=0.6, beta=1.0, device=DEV)
    s, a, r, s1, d = bench.synthetic_transitions(dict(S=S, A=A), cap)
    buf.load_experience({"obs": s}, a, r, {"obs": s1}, d)
    aug = augmentations.AugmentationSequence([augmentations.IdentityAug(B)])
    kw = dict(buffer=buf, agent=agent, target_agent=target, critic_optimizer=c_opt, encoder_optimizer=e_opt, log_alphas=las,
              batch_size=B, gamma=0.99, critic_clip=40.0, encoder_clip=40.0, target_critic_ensemble_n=2,
              weighted_bellman_temp=None, weight_type=None, pop=False, augmenter=aug, encoder_lambda=0.0, random_process=None,
              noise_clip=None, aug_mix=0.0, update_priorities=True, dr3_coeff=0.01)

    def upd():
        out = learning.critic_update(**kw)
        lu.soft_update(target.critics[0], agent.critics[0], 0.005)
        learning.offline_actor_update(buffer=buf, agent=agent, actor_optimizer=a_opt, encoder_optimizer=e_opt, batch_size=B,
                                      actor_clip=40.0, update_encoder=False, encoder_clip=40.0, augmenter=aug, actor_lambda=0.0,
                                      aug_mix=0.0, per=True, filter_=True)
        return out

    step, mode = graph_or_eager(upd)
    ms = timed(lambda k: step(), steps, warmup=5)
    print(json.dumps({"config": "afbc offline (C5)", "metric": "offline_steps_per_sec", "value": 1e3 / ms, "ms_per_step": ms, "mode": mode,
                      "step": "critic_update(DR3 0.01, clip 40, priority refresh) + Polyak + offline_actor_update(PER, filtered BC)",
                      "shape": dict(S=S, A=A, H=H, B=B, N=N, buffer=cap)}), flush=True)


def hbm_kernels():
    L, sp = _lib.lib(), _lib.stream_ptr()
    peaks = bench.measured_peaks()
    out = []
    # gather + DrQv2 shift + cast, C4 batch, 20k-frame ring (1.27 GB > L2), output 130 MB > L2
    C, HW, B, cap = 9, 84, 512, 20_000
    src = torch.empty((cap, C, HW, HW), dtype=torch.uint8, device=DEV).random_(0, 256)
    dst = torch.empty((B, C, HW, HW), device=DEV)
    idx = torch.randint(0, cap, (B,), device=DEV)
    shift = torch.randint(0, 9, (B, 2), device=DEV, dtype=torch.int32)
    ms = timed(lambda k: L.gather_aug_u8(src.data_ptr(), dst.data_ptr(), idx.data_ptr(), shift.data_ptr(), None, B, C, HW, HW, 4, 1, B, sp), 100)
    nbytes = B * C * HW * HW * 5 + B * 8
    out.append(("ssac_gather_aug_u8 (C4: 512 x 9x84x84 u8 -> f32)", nbytes, ms, "ring and output larger than L2"))
    for name, n, note in (("C2 721 930 params", 721_930, "L2-resident, as in the real step"), ("C5 2 150 402 params", 2_150_402, "L2-resident, as in the real step"),
                          ("64 Mi params", 1 << 26, "larger than L2: HBM bound")):
        p = torch.randn(n, device=DEV); t = torch.randn(n, device=DEV); g = torch.randn(n, device=DEV)
        m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV); ctl = torch.zeros(2, dtype=torch.int32, device=DEV)
        ms = timed(lambda k: L.polyak(t.data_ptr(), p.data_ptr(), n, 0.005, sp), 50)
        out.append((f"ssac_polyak ({name})", 12 * n, ms, note))
        ms = timed(lambda k: L.adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, ctl.data_ptr(), 3e-4, .9, .999, 1e-8, 0., None, 0., 0, sp), 50)
        out.append((f"ssac_adam_step ({name})", 28 * n, ms, note))
        ms = timed(lambda k: L.adam_polyak_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), t.data_ptr(), n, ctl.data_ptr(), 3e-4, .9, .999, 1e-8, 0., None, 0., 0, 0.005, sp), 50)
        out.append((f"ssac_adam_polyak_step ({name})", 36 * n, ms, note))
        del p, t, g, m, v
    for name, nb, ms, note in out:
        print(json.dumps({"kernel": name, "algorithmic_MB": nb / 1e6, "us": ms * 1e3, "GBps": nb / ms / 1e6,
                          "frac_of_hbm_peak": nb / ms / 1e6 / peaks["hbm_gbs"], "peak_GBps": peaks["hbm_gbs"], "note": note}), flush=True)


def mlp_kernels():
    """Tensor-core ensemble MLP at the shapes of the BASELINE configs: algorithmic TFLOP/s of the forward and of the
    backward (data + weight gradients) against the measured bf16 peak and against the 3xTF32 ceiling (bf16 / 6)."""
    from super_sac_b200 import _arena, _ops

    peaks = bench.measured_peaks()
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
    ap.add_argument("--only", default="sac,sunrise,drqv2,afbc,kernels,mlp")
    ap.add_argument("--steps", type=int, default=300)
    args = ap.parse_args()
    which = args.only.split(",")
    torch.cuda.set_device(DEV)
    if "kernels" in which:
        hbm_kernels()
    if "mlp" in which:
        mlp_kernels()
    if "sac" in which:
        state_config("sac (C1)", 1, 2, 2, 3, 1, 256, 256, args.steps)
    if "sunrise" in which:
        state_config("sunrise (C3)", 5, 2, 2, 17, 6, 256, 256, args.steps, weight_type="sunrise", temp=20.0)
    if "drqv2" in which:
        drqv2_config(min(args.steps, 50))
    if "afbc" in which:
        afbc_config(min(args.steps, 100))
