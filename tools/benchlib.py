"""Shared pieces of bench.py and tools/bench_configs.py: the five BASELINE.json workloads built through the reference's
own API surface (the SAME builder constructs them from ``super_sac_b200`` or from the unmodified reference package --
that is what "drop-in" means), synthetic data, clock sampling, graph timing helpers.

    C1 sac      SAC, 2 critics, obs 3 / act 1, H=256, B=256                       (experiments/gym/sac.gin)
    C2 redq     REDQ N=10, M=2, obs 17 / act 6, H=256, B=256, UTD 20              (redq.gin)   <- the headline
    C3 sunrise  SUNRISE E=5 x N=2, weighted Bellman backups T=20, B=256           (sunrise.gin)
    C4 drqv2    DrQv2 pixels: u8 9x84x84 ring, Drqv2Aug, BigPixelEncoder, H=1024, B=512, TD3 noise (dmc/drqv2.gin)
    C5 afbc     offline AFBC: 2 M transitions, B=1024, H=1024, DR3 0.01, clips 40, PER actor sampling (d4rl/basic_afbc.gin)
"""
import json
import math
import os
import subprocess
import threading
import time
from itertools import chain

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONFIGS = {
    "redq": dict(E=1, N=10, M=2, S=17, A=6, H=256, B=256, target_delay=2, tau=0.005, lr=3e-4, buffer=1_000_000, utd=20,
                 workload="REDQ-10 critic_update+Polyak, obs17/act6, B=256, 2x256 MLP, M=2, target_delay=2"),
    "sac": dict(E=1, N=2, M=2, S=3, A=1, H=256, B=256, target_delay=2, tau=0.005, lr=3e-4, buffer=100_000, utd=1,
                workload="SAC (2 critics) critic_update+Polyak, obs3/act1, B=256, 2x256 MLP, target_delay=2"),
    "sunrise": dict(E=5, N=2, M=2, S=17, A=6, H=256, B=256, target_delay=2, tau=0.005, lr=3e-4, buffer=500_000, utd=1,
                    weight_type="sunrise", temp=20.0,
                    workload="SUNRISE 5 members x 2 critics critic_update+Polyak, weighted Bellman backups T=20, obs17/act6, B=256"),
    "drqv2": dict(E=1, N=2, M=2, S=50, A=6, H=1024, B=512, target_delay=1, tau=0.01, lr=1e-4, buffer=20_000, utd=1, pixels=(9, 84, 84),
                  workload="DrQv2 pixel critic_update+Polyak: u8 9x84x84 ring, Drqv2Aug(pad 4), BigPixelEncoder(50), 2 critics "
                           "56-1024-1024-1, B=512, TD3 noise, encoder tau 1.0"),
    "afbc": dict(E=1, N=2, M=2, S=17, A=6, H=1024, B=1024, target_delay=1, tau=0.005, lr=3e-4, buffer=2_000_000, utd=1, offline=True,
                 workload="offline AFBC step: critic_update(DR3 0.01, clip 40, priority refresh)+Polyak+offline_actor_update(PER, "
                          "filtered BC), 2M-transition buffer, B=1024, 3x1024 MLP"),
}


def synthetic_transitions(cfg, n, seed=0):
    rng = np.random.default_rng(seed)
    s = rng.standard_normal((n, cfg["S"]), dtype=np.float32)
    a = rng.uniform(-1, 1, (n, cfg["A"])).astype(np.float32)
    r = rng.standard_normal(n, dtype=np.float32)
    s1 = rng.standard_normal((n, cfg["S"]), dtype=np.float32)
    d = (rng.uniform(size=n) < 0.01)
    return s, a, r, s1, d


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], source="MEASURED_PEAKS.json (measured)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="B200_PROFILING.md fallback")


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons every 100 ms for the whole arm; ``region()`` marks the timed stretches and
    the summary is taken over the samples that fall inside them (plus one sample of slack on either side)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu, self.regions = [], None, gpu_index, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [x.strip() for x in line.split(",")]))

    class _Region:
        def __init__(self, owner):
            self.owner = owner

        def __enter__(self):
            self.t0 = time.monotonic()

        def __exit__(self, *a):
            self.owner.regions.append((self.t0, time.monotonic()))

    def region(self):
        return ClockSampler._Region(self)

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"], samples=0)
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        slack = 0.15
        inside = [r for t, r in self.rows if any(a - slack <= t <= b + slack for a, b in self.regions)] if self.regions else []
        rows = inside if inside else [r for _, r in self.rows]
        ok = [r for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        sm = [float(r[1]) for r in ok]
        mx = [float(r[2]) for r in ok if r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in ok if r[3].replace(".", "").isdigit()]
        reasons = set()
        for r in ok:
            for nm, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                    samples=len(sm), samples_total=len(self.rows), power_w_max=max(pw) if pw else None,
                    what="samples inside the timed regions of this arm (value, e2e, secondary configs)")


# ------------------------------------------------------------------------------------------------ workloads
class _ActionSpace:
    def __init__(self, dim):
        self.low = -np.ones(dim, dtype=np.float32)
        self.high = np.ones(dim, dtype=np.float32)
        self.shape = (dim,)


def _encoders(pkg):
    class IdentityEncoder(pkg.nets.Encoder):   # experiments/gym/train_gym.py:18-28
        def __init__(self, dim):
            super().__init__()
            self._dim = dim

        @property
        def embedding_dim(self):
            return self._dim

        def forward(self, obs_dict):
            return obs_dict["obs"]

    class PixelEncoder(pkg.nets.Encoder):      # experiments/dmc/train_dmc_from_pixels.py:15-27
        def __init__(self, shape, dim):
            super().__init__()
            self.net = pkg.nets.cnns.BigPixelEncoder(shape, dim)
            self._dim = dim

        @property
        def embedding_dim(self):
            return self._dim

        def forward(self, obs_dict):
            return self.net(obs_dict["pixels"])

    return IdentityEncoder, PixelEncoder


class Workload:
    """One BASELINE config built from ``pkg`` (``super_sac_b200`` or the unmodified reference ``super_sac``): the same
    constructor calls main.py:188-244 / :321 makes, then ``step(k)`` = one update of that config as the reference's
    training loop runs it (main.py:380-414 online, :456-477 offline)."""

    def __init__(self, pkg, name, device, seed=0, buffer_size=None, fill_on_device=False, overrides=None, torch_seed=None):
        import copy

        cfg = self.cfg = dict(CONFIGS[name])
        cfg.update(overrides or {})   # e.g. this rank's share of the ensemble when it is sharded over the GPUs
        self.name, self.pkg, self.device = name, pkg, torch.device(device)
        ours = pkg.__name__ == "super_sac_b200"
        if ours:
            pkg.manual_seed(seed)
        torch.manual_seed(seed if torch_seed is None else torch_seed)
        IdentityEncoder, PixelEncoder = _encoders(pkg)
        E, N, S, A, H, B = cfg["E"], cfg["N"], cfg["S"], cfg["A"], cfg["H"], cfg["B"]
        pixels = cfg.get("pixels")
        det = pixels is not None
        enc = PixelEncoder(pixels, S) if pixels else IdentityEncoder(S)
        self.agent = pkg.Agent(act_space_size=A, encoder=enc,
                               actor_network_cls=pkg.nets.mlps.ContinuousDeterministicActor if det else pkg.nets.mlps.ContinuousStochasticActor,
                               critic_network_cls=pkg.nets.mlps.ContinuousCritic, ensemble_size=E, num_critics=N, hidden_size=H,
                               auto_rescale_targets=False, log_std_low=-5.0, log_std_high=2.0)
        self.agent.to(self.device)
        self.target = copy.deepcopy(self.agent)
        self.target.to(self.device)
        lr = cfg["lr"]
        self.critic_opt = torch.optim.Adam(chain(*(c.parameters() for c in self.agent.critics)), lr=lr, betas=(0.9, 0.999))
        self.actor_opt = torch.optim.Adam(chain(*(a.parameters() for a in self.agent.actors)), lr=lr, betas=(0.9, 0.999))
        self.enc_opt = torch.optim.Adam(self.agent.encoder.parameters(), lr=1e-4, betas=(0.9, 0.999))
        init_alpha = 1e-15 if (cfg.get("offline") or det) else 0.1
        self.log_alphas, self.alpha_opts = [], []
        for _ in range(E):
            la = torch.Tensor([math.log(init_alpha)]).to(self.device)
            la.requires_grad = True
            self.log_alphas.append(la)
            self.alpha_opts.append(torch.optim.Adam([la], lr=1e-4, betas=(0.5, 0.999)))
        n = buffer_size or cfg["buffer"]
        self.buffer_size = n
        bkw = dict(device=self.device) if ours else {}
        self.buffer = pkg.replay.ReplayBuffer(n, **bkw)
        if pixels:
            C, Hh, Ww = pixels
            if ours and fill_on_device:
                # fill the device ring directly (12.7 GB of host staging would only measure PCIe)
                z = np.zeros((2, C, Hh, Ww), np.uint8)
                self.buffer.load_experience({"pixels": z}, np.zeros((2, A), np.float32), np.zeros(2, np.float32), {"pixels": z}, np.zeros(2, bool))
                st = self.buffer._storage
                g = torch.Generator(device=self.device).manual_seed(seed)
                st.s_stack["pixels"].random_(0, 256, generator=g)
                st.s1_stack["pixels"].random_(0, 256, generator=g)
                st.action_stack.uniform_(-1, 1, generator=g)
                st.reward_stack.normal_(generator=g)
                st._max_filled, st._next_idx = n, 0
                self.buffer._n_filled_dev.fill_(n)
            else:
                rng = np.random.default_rng(seed)
                self.buffer.load_experience({"pixels": rng.integers(0, 256, (n, C, Hh, Ww), dtype=np.uint8)},
                                            rng.uniform(-1, 1, (n, A)).astype(np.float32), rng.standard_normal(n).astype(np.float32),
                                            {"pixels": rng.integers(0, 256, (n, C, Hh, Ww), dtype=np.uint8)}, rng.uniform(size=n) < 0.01)
            self.augmenter = pkg.augmentations.AugmentationSequence([pkg.augmentations.Drqv2Aug(B)])
            self.noise = pkg.learning_utils.GaussianExplorationNoise(_ActionSpace(A), start_scale=1.0, final_scale=0.1)
            self.obs_key = "pixels"
        else:
            s, a, r, s1, d = synthetic_transitions(cfg, n, seed)
            self.buffer.load_experience({"obs": s}, a, r, {"obs": s1}, d)
            self.augmenter = pkg.augmentations.AugmentationSequence([pkg.augmentations.IdentityAug(B)])
            self.noise = None
            self.obs_key = "obs"
        offline = bool(cfg.get("offline"))
        self.kw = dict(buffer=self.buffer, agent=self.agent, target_agent=self.target, critic_optimizer=self.critic_opt,
                       encoder_optimizer=self.enc_opt, log_alphas=self.log_alphas, batch_size=B,
                       gamma=0.99**3 if pixels else 0.99, critic_clip=40.0 if offline else None,
                       encoder_clip=40.0 if offline else None, target_critic_ensemble_n=cfg["M"],
                       weighted_bellman_temp=cfg.get("temp"), weight_type=cfg.get("weight_type"), pop=False,
                       augmenter=self.augmenter, encoder_lambda=0.0, random_process=self.noise,
                       noise_clip=0.3 if pixels else None, aug_mix=1.0 if pixels else 0.0, discrete=False, per=False,
                       update_priorities=offline, dr3_coeff=0.01 if offline else 0.0)

    # ---- the pieces of one update ---------------------------------------------------------------------------------
    def critic_update(self):
        return self.pkg.learning.critic_update(**self.kw)

    def polyak(self):
        lu, cfg = self.pkg.learning_utils, self.cfg
        for ac, tc in zip(self.agent.critics, self.target.critics):
            lu.soft_update(tc, ac, cfg["tau"])
        lu.soft_update(self.target.encoder, self.agent.encoder, 1.0 if cfg.get("pixels") else 0.01)

    def offline_actor(self):
        B = self.cfg["B"]
        return self.pkg.learning.offline_actor_update(
            buffer=self.buffer, agent=self.agent, actor_optimizer=self.actor_opt, encoder_optimizer=self.enc_opt, batch_size=B,
            actor_clip=40.0, update_encoder=False, encoder_clip=40.0, augmenter=self.augmenter, actor_lambda=0.0, aug_mix=0.0,
            per=True, filter_=True)

    def step(self, k, with_polyak=None):
        """One update: critic_update + the conditional Polyak step (main.py:409-414) [+ the offline actor update]."""
        out = self.critic_update()
        if with_polyak if with_polyak is not None else (k % self.cfg["target_delay"] == 0):
            self.polyak()
        if self.cfg.get("offline"):
            self.offline_actor()
        return out

    def actor_and_alpha(self, replay_dicts):
        B, A = self.cfg["B"], self.cfg["A"]
        L = self.pkg.learning
        L.online_actor_update(buffer=self.buffer, agent=self.agent, pop=False, actor_optimizer=self.actor_opt, log_alphas=self.log_alphas,
                              batch_size=B, clip=None, random_process=self.noise, noise_clip=0.3 if self.noise else None,
                              augmenter=self.augmenter, aug_mix=self.kw["aug_mix"], premade_replay_dicts=replay_dicts)
        return L.alpha_update(buffer=self.buffer, agent=self.agent, optimizers=self.alpha_opts, batch_size=B, log_alphas=self.log_alphas,
                              augmenter=self.augmenter, aug_mix=self.kw["aug_mix"], target_entropy=-float(A),
                              premade_replay_dicts=replay_dicts, discrete=False)

    # ---- one host transition per step (the e2e path pushes it into the replay ring) ------------------------------------
    def host_transitions(self, n, seed=123):
        rng = np.random.default_rng(seed)
        cfg = self.cfg
        if cfg.get("pixels"):
            C, Hh, Ww = cfg["pixels"]
            obs = rng.integers(0, 256, (n, C, Hh, Ww), dtype=np.uint8)
            obs1 = rng.integers(0, 256, (n, C, Hh, Ww), dtype=np.uint8)
        else:
            obs = rng.standard_normal((n, cfg["S"]), dtype=np.float32)
            obs1 = rng.standard_normal((n, cfg["S"]), dtype=np.float32)
        return obs, rng.uniform(-1, 1, (n, cfg["A"])).astype(np.float32), rng.standard_normal(n, dtype=np.float32), obs1, rng.uniform(size=n) < 0.01

    def push(self, tr, j):
        obs, a, r, obs1, d = tr
        self.buffer.push({self.obs_key: obs[j]}, a[j], float(r[j]), {self.obs_key: obs1[j]}, bool(d[j]))


def critic_flops(cfg):
    D, H, B, N, M, E = cfg["S"] + cfg["A"], cfg["H"], cfg["B"], cfg["N"], cfg["M"], cfg["E"]
    fwd = 2 * B * (D * H + H * H + H)            # one critic net forward
    bwd = 2 * B * (2 * H + 2 * H * H + D * H)    # dz2, gW3, gW2, dz1, gW1 (no dX: identity encoder)
    O = cfg["A"] if cfg.get("pixels") else 2 * cfg["A"]
    actor = 2 * B * (cfg["S"] * H + H * H + O * H)
    return dict(bwd_group=E * N * bwd, fwd_group=E * N * fwd, update=E * (actor + M * fwd + N * (fwd + bwd)))


# ------------------------------------------------------------------------------------------------ timing helpers
def timed_events(step, steps, warmup=3):
    """ms per step of ``step(k)`` with CUDA events on the current stream (warm-up first, synchronise on both sides)."""
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


def graph_time(fn, per=10, iters=20):
    """ms per call of ``fn`` replayed back to back from a CUDA graph (no launch gaps): kernel-level rooflines."""
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


def buf_bytes(buf):
    st = buf._storage
    n = st.action_stack.numel() * 4 + st.reward_stack.numel() * 4 + st.done_stack.numel()
    for d in (st.s_stack, st.s1_stack):
        for v in d.values():
            n += v.numel() * v.element_size()
    return n


# ------------------------------------------------------------------------------------------------ CPU arms
def reference_package():
    """The unmodified reference (baseline/_ref) on the CPU, or None when it did not travel."""
    import sys

    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from baseline import ref_import

    return ref_import.import_reference(device="cpu") if ref_import.available() else None


def cpu_step_fn(name, seed=0, buffer_size=None):
    """step(k) of config ``name`` on the host cores: the UNMODIFIED reference when baseline/_ref travelled ("reference"),
    else the oracle port of the state configs ("port").  Returns (step, kind)."""
    ref = reference_package()
    if ref is not None:
        n = buffer_size or min(CONFIGS[name]["buffer"], 2_000 if CONFIGS[name].get("pixels") else 100_000)
        w = Workload(ref, name, "cpu", seed=seed, buffer_size=n)
        return w.step, "reference"
    return oracle_step_fn(CONFIGS[name], seed=seed), "port"


def oracle_step_fn(cfg, n_buf=100_000, seed=0):
    """One CPU critic update (+Polyak by the target_delay rule) of the oracle port (oracle/update_oracle.py), sampling its
    batch from a host numpy buffer like the reference does (replay.py:121-126).  State configs without DR3 / PER only."""
    import sys

    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import update_oracle as uo

    if cfg.get("pixels") or cfg.get("offline"):
        raise RuntimeError("the oracle port times the state configs only; ship baseline/_ref for the others")
    gen = torch.Generator().manual_seed(seed)
    agent = uo.OracleAgent(cfg["E"], cfg["N"], cfg["S"], cfg["A"], cfg["H"], log_std_low=-5.0, log_std_high=2.0)
    agent.actors.random_init(gen)
    agent.critics.random_init(gen)
    target = agent.clone()
    opt = uo.Adam(agent.critics.tensors(), lr=cfg["lr"])
    s, a, r, s1, d = synthetic_transitions(cfg, n_buf, seed)
    rng = np.random.default_rng(seed + 1)
    log_alphas = [torch.tensor([math.log(0.1)]) for _ in range(cfg["E"])]
    hp = dict(gamma=0.99, weight_type=cfg.get("weight_type"), weight_temp=cfg.get("temp"))
    E, N, M, B, A = cfg["E"], cfg["N"], cfg["M"], cfg["B"], cfg["A"]

    def step(k):
        batches, rands = [], []
        for _ in range(E):
            idx = rng.integers(0, n_buf, B)
            t = torch.from_numpy
            batches.append(({"obs": t(s[idx])}, t(a[idx]), t(r[idx]).reshape(-1, 1), {"obs": t(s1[idx])},
                            t(d[idx].astype(np.float32)).reshape(-1, 1)))
            rands.append(dict(eps=torch.randn(B, A), subset=[int(x) for x in rng.permutation(N)[:M]]))
        logs, _ = uo.critic_update(agent, target, batches, rands, hp, log_alphas, opt)
        if k % cfg["target_delay"] == 0:
            uo.soft_update(target.critics.tensors(), agent.critics.tensors(), cfg["tau"])
        return logs

    return step


def time_cpu(name, steps, warmup, threads=None, budget_s=None):
    """(updates/s, seconds, kind, threads) of ``steps`` CPU updates of config ``name`` with ``threads`` intra-op threads
    (default: all cores).  ``budget_s`` stops early (a bounded sample) once that much time has been spent."""
    threads = threads or (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    step, kind = cpu_step_fn(name)
    for k in range(warmup):
        step(k)
    t0 = time.perf_counter()
    done = 0
    for k in range(steps):
        step(k)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done / dt, dt, kind, threads, done


def cpu_thread_sweep(name, steps, warmup, budget_s=6.0):
    """The CPU arm swings several-fold with the thread count (over-subscription on 30+ core hosts): time the same
    sample at 1 / 8 / 16 / all threads and quote the best."""
    ncpu = os.cpu_count() or 1
    rows = []
    for th in sorted({1, min(8, ncpu), min(16, ncpu), ncpu}):
        ups, dt, kind, _, done = time_cpu(name, steps, warmup, threads=th, budget_s=budget_s)
        rows.append(dict(threads=th, updates_per_s=ups, steps=done, seconds=dt, kind=kind))
    best = max(rows, key=lambda r: r["updates_per_s"])
    return best, rows
