#!/usr/bin/env python
"""bench.py -- SAC gradient updates/sec of the off-policy update step (BASELINE.json; headline: configs[1], REDQ-10).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config redq|sac|sunrise|drqv2|afbc]

One "step" = one update = one ``critic_update`` + the conditional Polyak target update of main.py:380-414 (for the
offline config also its ``offline_actor_update``), on synthetic batches of the config's shape.  ONE JSON line:

  value     whole-job updates/s with everything resident in HBM: CUDA-graph replay of the update (eager launches for
            the configs that cannot be captured yet), CUDA events, no host reads
  e2e       the same step through the drop-in Python API with HOST inputs: every step pushes one host transition
            into the device replay ring (H2D), calls learning.critic_update + learning_utils.soft_update, and reads
            the logged scalars back (D2H)
  roofline  the dominant kernel of the config, CUDA-event timed inside this process against MEASURED_PEAKS.json
  cpu_baseline  the reference's own CPU implementation on this box's cores (the UNMODIFIED reference from
            baseline/_ref when it travelled, else the oracle port), best of a 1/8/16/all-thread sweep
  secondary (default config only) the other BASELINE configs' step times, the HBM-bound kernels at C4/C5 sizes and the
            unmodified reference on this GPU through stock PyTorch-CUDA (clearly labelled: a different arm)

--impl reference times the reference's CPU implementation alone, on the same workload.  Under torchrun (--gpus N > 1)
every rank runs an independent learner replica (weak scaling, no data-path collective); the ensemble-sharded single
learner (critics / members partitioned over the ranks, exchange over NVLink) is checked for parity first
("sharded_parity") and timed next to the replicas ("sharded_ensemble").
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402,F401
import torch  # noqa: E402

import benchlib as bl  # noqa: E402
from benchlib import CONFIGS  # noqa: E402

METRIC = "sac_gradient_updates_per_sec"
_emit = lambda line: print(line, flush=True)   # noqa: E731  (replaced in main)


def base_line(args, cfg, world):
    return {"metric": METRIC, "unit": "updates/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "name": args.config}}


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference_arm(args, cfg, rank, world):
    """The reference's own CPU implementation of the path on the box's host cores (rank 0 only)."""
    if rank != 0:
        return
    warm = max(args.warmup, 3)
    best, sweep = bl.cpu_thread_sweep(args.config, min(args.steps, 40), 2, budget_s=4.0)   # pick the thread count first
    ups, dt, kind, threads, done = bl.time_cpu(args.config, args.steps, warm, threads=best["threads"], budget_s=150.0)
    line = base_line(args, cfg, args.gpus)
    line.update({
        "impl": "reference", "value": ups, "ms_per_step": 1e3 * dt / done, "steps": done,
        "arm": ("UNMODIFIED reference (baseline/_ref/super_sac, stub-imported) on the host CPU: learning.critic_update + soft_update, "
                "host numpy replay" if kind == "reference" else "CPU oracle port of the reference update (oracle/update_oracle.py)"),
        "cpu_baseline": {"value": ups, "unit": "updates/s", "cores": threads, "kind": kind, "host_cores": os.cpu_count(),
                         "sample": f"{done} updates after {warm} warm-up, torch-CPU fp32, {threads} intra-op threads "
                                   f"(best of the sweep {[(r['threads'], round(r['updates_per_s'], 1)) for r in sweep]})"},
        "e2e": {"value": ups, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })
    _emit(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def count_launches(fn, fused=False):
    """Count kernel launches by wrapping the C-ABI entry points with their known launch multiplicities (fused: the
    single-kernel forward applies, i.e. tcgen05 and a 2x256-class network)."""
    from super_sac_b200 import _lib as L

    lib = L.lib()
    per_call = {"mlp_forward": 1 if fused else 3, "actor_forward_sample": 1 if fused else 3, "scatter_fields": 1, "polyak": 1,
                "polyak_multi": 1, "adam_step": 1, "adam_polyak_step": 1, "sumsq": 1, "rng_fill": 1, "gather_rows": 1,
                "gather_aug_u8": 1, "tanh_normal_forward": 1, "tanh_normal_backward": 1, "tanh_normal_logprob": 1,
                "det_head_forward": 1, "det_head_backward": 1, "td_target": 1, "critic_loss_seed": 1, "actor_loss_seed": 1,
                "backup_weights": 1, "tree_set": 1, "tree_sample": 1, "dr3_dot": 1, "advantage": 1, "min_over_nets": 1,
                "alpha_step": 1, "sum_groups": 1, "mlp_backward_dact": 2}
    per_call = {k: v for k, v in per_call.items() if hasattr(lib, k)}
    counter = {"n": 0}
    saved = {}

    def wrap(name, f, mult):
        def g(*a):
            if name == "critic_forward_loss":
                counter["n"] += ({0: 1, 1: 1, 2: 1} if fused else {0: 3, 1: 2, 2: 1})[a[24]]   # phase: all / hidden / output + loss
            elif name == "mlp_backward":
                want_dw = a[17] is not None   # dz2, gW3, gW2, dz1, gW1 (+dx): 5 launches with weight grads, 2 (+1) without
                counter["n"] += (5 if want_dw else 2) + (1 if a[24] is not None else 0)
            elif name == "mlp_backward_pre":
                counter["n"] += 1    # u GEMM (v = W3 .* (h2 > 0) is generated while its A tiles are staged)
            elif name == "mlp_backward_post":
                counter["n"] += 2    # gW2 GEMM, gW1 / gb1 / gW3 / gb3 reduction
            else:
                counter["n"] += mult
            return f(*a)
        return g

    for name, mult in list(per_call.items()) + [("critic_forward_loss", 0), ("mlp_backward", 0), ("mlp_backward_pre", 0), ("mlp_backward_post", 0)]:
        saved[name] = getattr(lib, name)
        setattr(lib, name, wrap(name, saved[name], mult))
    try:
        fn()
    finally:
        for name, f in saved.items():
            setattr(lib, name, f)
    return counter["n"]


def graph_or_eager(fn):
    from super_sac_b200 import graphed

    try:
        g = graphed.GraphedCall(fn)
        return g.replay, "CUDA-graph replay", g
    except Exception as e:  # noqa: BLE001  (configs with host-side decisions / autograd hand-off are not capturable yet)
        torch.cuda.synchronize()
        return fn, f"eager launches (not capturable: {type(e).__name__}: {str(e)[:80]})", None


def time_config(W, steps, warmup):
    """Device-resident ms/step of workload ``W``: graph replay where the update can be captured, else eager."""
    delay = W.cfg["target_delay"]
    with_pol, mode, g1 = graph_or_eager(lambda: W.step(0, with_polyak=True))
    if delay > 1 and g1 is not None:
        no_pol, _, _ = graph_or_eager(lambda: W.step(1, with_polyak=False))
    else:
        no_pol = (lambda: W.step(1, with_polyak=False)) if delay > 1 else with_pol
    step = lambda k: (with_pol if k % delay == 0 else no_pol)()
    ms = bl.timed_events(step, steps, warmup)
    return ms, mode, g1


def run_gpu_arm(args, cfg, rank, world, local_rank):
    import super_sac_b200 as ssb
    from super_sac_b200 import _lib, graphed, learning

    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    _lib.require_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=device)
    sampler = bl.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ssb.set_mlp_impl(args.mlp_impl)
    ssb.set_overlap(not args.no_overlap)
    ssb.set_pdl(not args.no_pdl)

    sharded_parity = None
    if dist is not None:
        sharded_parity = run_sharded_parity(rank, world, device, dist)

    W = bl.Workload(ssb, args.config, device, seed=rank, fill_on_device=True)
    ring_mb = bl.buf_bytes(W.buffer) / 1e6
    warm = max(args.warmup, 3)

    # ---- count our kernel launches per step (eager, via the C-ABI entry points) --------------------
    fused_fwd = cfg["H"] <= 256 and cfg["H"] % 16 == 0 and cfg["S"] + cfg["A"] <= 32 and args.mlp_impl == "tcgen05"
    launch_count = count_launches(lambda: (W.step(0), W.step(1)), fused_fwd) / 2.0

    # ---- device-resident timed region ------------------------------------------------------------
    delay = cfg["target_delay"]
    with_pol, mode, g1 = graph_or_eager(lambda: W.step(0, with_polyak=True))
    if delay > 1:
        no_pol = graph_or_eager(lambda: W.step(1, with_polyak=False))[0] if g1 is not None else (lambda: W.step(1, with_polyak=False))
    else:
        no_pol = with_pol

    def gstep(k):
        (with_pol if k % delay == 0 else no_pol)()

    # The reference's training loop runs UTD updates back to back per environment step (main.py:380-414; REDQ: 20).  Such
    # a block is captured as ONE graph in which the target side of update k+1 overlaps update k's backward / Adam
    # (learning_utils.pipelined_updates: bit-identical to the sequential schedule, tests/test_cuda_update_parity.py).
    utd = cfg.get("utd", 1)
    block = None
    if utd > 1 and utd % delay == 0 and g1 is not None and not args.no_pipeline:
        from super_sac_b200 import learning_utils as lu

        def utd_block():
            with lu.pipelined_updates():
                for u in range(utd):
                    out = W.step(u)
            return out

        try:
            block = graphed.GraphedCall(utd_block, warmup=1)
            mode = "CUDA-graph replay of the UTD block (%d updates per graph, target side of update k+1 next to update k)" % utd
        except Exception as e:  # noqa: BLE001
            torch.cuda.synchronize()
            block = None
            print("[bench] UTD block not capturable: %s: %s" % (type(e).__name__, str(e)[:300]), file=sys.stderr, flush=True)
            mode += " (UTD block not capturable: %s)" % type(e).__name__

    def run_steps(n):
        k = 0
        if block is not None:
            for _ in range(n // utd):
                block.replay()
            k = n - n % utd
        for j in range(k, n):
            gstep(j)

    run_steps(max(warm, utd if block is not None else 0))
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    with sampler.region():
        ev0.record()
        run_steps(args.steps)
        ev1.record()
        torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * args.steps / (ms * 1e-3)
    glogs = (block or g1).logs() if g1 is not None else {}

    # ---- e2e through the drop-in API with host inputs ---------------------------------------------
    e2e_steps = min(args.steps, 2000)
    tr = W.host_transitions(256 if cfg.get("pixels") else 4096, seed=123 + rank)
    n_tr = len(tr[1])
    eager_logs = learning._critic_update_impl(**W.kw)[0]
    d2h = 4 * eager_logs._n

    def time_e2e(lazy):
        """push (H2D) + update + Polyak + logged scalars read on the host, every step.  lazy: the drop-in call returns
        after cudaGraphLaunch and step k's scalars are read while step k+1 runs (one step late, every step); strict: the
        call itself waits for them."""
        # the drop-in calls replay captured graphs from their third call on; lazy: consecutive updates also overlap on the device
        graphed.enable_auto_graphs(True, lazy_logs=lazy, pipeline=lazy and not args.no_pipeline)
        prev = {"logs": None, "sum": 0.0}

        def e2e_step(k):
            W.push(tr, k % n_tr)
            logs = W.step(k)[0]
            if prev["logs"] is not None:
                prev["sum"] += float(prev["logs"]["losses/critic_overall_loss"])   # the D2H read of the previous step's result
            prev["logs"] = logs if lazy else None
            if not lazy:
                prev["sum"] += float(logs["losses/critic_overall_loss"])

        for k in range(6):
            e2e_step(k)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        with sampler.region():
            t0 = time.perf_counter()
            for k in range(e2e_steps):
                e2e_step(k)
            if prev["logs"] is not None:
                prev["sum"] += float(prev["logs"]["losses/critic_overall_loss"])
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        graphed.enable_auto_graphs(False)
        if not np.isfinite(prev["sum"]):
            raise RuntimeError("e2e: non-finite loss")
        if dist is not None:
            t = torch.tensor([dt], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return world * e2e_steps / dt

    e2e_strict = time_e2e(False)
    e2e_value = time_e2e(True)
    h2d = W.buffer._stage_bytes   # one pinned staging row per pushed transition (s, s1, a, r, d, tree index, priority, fill level)

    sharded = None
    can_shard = (cfg["N"] >= world) if cfg["E"] == 1 else (cfg["E"] >= world)
    if dist is not None and can_shard and not cfg.get("pixels") and not cfg.get("offline"):
        with sampler.region():
            sharded = time_sharded(cfg, args, device, rank, world, dist)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- the full SAC step, rank 0 --------------------------------------------------------
    sac_step = None
    if not cfg.get("offline"):
        try:
            with sampler.region():
                sac_step = time_full_sac_step(W)
        except Exception as e:   # a secondary figure must not take the headline down with it
            sac_step = {"unavailable": "%s: %s" % (type(e).__name__, e)}

    # ---- roofline of the dominant kernel (rank 0) -----------------------------------------------------
    with sampler.region():
        roofline = dominant_kernel_roofline(args, cfg, W)

    secondary = None
    if args.config == "redq" and not args.skip_secondary:
        del W
        torch.cuda.empty_cache()
        secondary = run_secondary(args, sampler)

    best, sweep = bl.cpu_thread_sweep(args.config, 60 if not cfg.get("pixels") else 3, 3 if not cfg.get("pixels") else 1, budget_s=5.0)
    clocks = sampler.stop()
    line = base_line(args, cfg, world)
    line.update({
        "value": value, "ms_per_step": ms / args.steps,
        "run": {"mode": mode + ("; update = critic_update (+Polyak every %d. step)" % delay if delay > 1 else " of the update"),
                "parallelism": "1 learner" if world == 1 else f"{world} independent learner replicas (no data-path collective)",
                "l2": "replay ring %.0f MB (random rows) vs 126 MB L2; parameters + moments are L2-resident by design" % ring_mb,
                "impl": "ensemble MLP GEMMs: " + ssb.get_mlp_impl(),
                "overlap": "two-stream fork/join inside the update" if not args.no_overlap else "off",
                "pdl": "programmatic dependent launch along the critical path" if not args.no_pdl else "off"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "updates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "path": "buffer.push(host transition, H2D) + learning.critic_update (auto CUDA graph where "
                                            "capturable) + soft_update + logged scalars read back (D2H) on every step",
                "logs_read": "every step, one step late: graphed.enable_auto_graphs(lazy_logs=True) returns after cudaGraphLaunch and "
                             "the host reads step k's loss while step k+1 runs",
                "strict_value": e2e_strict, "strict_logs_read": "the drop-in call itself waits for its scalars (lazy_logs=False)"},
        "gpu_launches": int(round(launch_count * args.steps)),
        "gpu_launches_per_step": launch_count,
        "roofline": roofline,
        "cpu_baseline": {"value": best["updates_per_s"], "unit": "updates/s", "cores": best["threads"], "kind": best["kind"],
                         "host_cores": os.cpu_count(),
                         "sample": f"{best['steps']} updates of the same workload, torch-CPU fp32, best of the thread sweep "
                                   f"{[(r['threads'], round(r['updates_per_s'], 1)) for r in sweep]}"},
        "full_sac_step": sac_step, "flops_per_update": bl.critic_flops(cfg)["update"],
        "sharded_parity": sharded_parity, "sharded_ensemble": sharded, "secondary": secondary,
        "sample_logs": {k: float(v) for k, v in list(glogs.items())[:4]},
    })
    _emit(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def run_secondary(args, sampler):
    """The other BASELINE configs (device-resident ms/step, CPU reference on a bounded sample), the HBM-bound kernels at
    C4 / C5 sizes, and the unmodified reference on this GPU through stock PyTorch-CUDA."""
    import super_sac_b200 as ssb

    out = {}
    for name in ("sac", "sunrise", "afbc", "drqv2"):
        try:
            W = bl.Workload(ssb, name, torch.device("cuda", torch.cuda.current_device()), seed=0, fill_on_device=True)
            steps = {"sac": 400, "sunrise": 200, "afbc": 60, "drqv2": 20}[name]
            with sampler.region():
                ms, mode, _ = time_config(W, steps, 5)
            ent = {"workload": W.cfg["workload"], "ms_per_step": ms, "updates_per_s": 1e3 / ms, "mode": mode}
            if name == "drqv2":
                with sampler.region():
                    ent["sample_move_and_augment"] = time_pixel_sampling(W)
                    try:
                        ent["nstep_frame_ring"] = time_nstep_ring_sampling(W)
                    except Exception as e:  # noqa: BLE001  (a secondary figure)
                        ent["nstep_frame_ring"] = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}
            del W
            torch.cuda.empty_cache()
            cpu_steps = {"sac": 40, "sunrise": 20, "afbc": 5, "drqv2": 2}[name]
            ups, dt, kind, th, done = bl.time_cpu(name, cpu_steps, 1, threads=min(16, os.cpu_count() or 1), budget_s=8.0)
            ent["cpu_baseline"] = {"value": ups, "unit": "updates/s", "cores": th, "kind": kind, "sample": f"{done} updates"}
        except Exception as e:  # noqa: BLE001
            ent = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}
        out[name] = ent
    try:
        with sampler.region():
            out["pixel_encoder"] = time_pixel_encoder()
    except Exception as e:  # noqa: BLE001
        out["pixel_encoder"] = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}
    try:
        with sampler.region():
            out["hbm_kernels"] = hbm_kernels()
    except Exception as e:  # noqa: BLE001
        out["hbm_kernels"] = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}
    out["reference_on_this_gpu_pytorch_cuda"] = reference_on_gpu(args)
    return out


def reference_on_gpu(args):
    """A DIFFERENT arm, for context only: the unmodified reference (baseline/_ref) running the REDQ-10 update on this B200
    through stock PyTorch-CUDA -- what a user gets today on this GPU without this library."""
    import subprocess

    code = ("import sys,json,time,torch; sys.path.insert(0,%r); sys.path.insert(0,%r)\n"
            "import benchlib as bl\n"
            "from baseline import ref_import\n"
            "ref=ref_import.import_reference(device='cuda')\n"
            "W=bl.Workload(ref,'redq','cuda',seed=0,buffer_size=100000)\n"
            "[W.step(k) for k in range(10)]\n"
            "torch.cuda.synchronize(); t0=time.perf_counter()\n"
            "[W.step(k) for k in range(100)]\n"
            "torch.cuda.synchronize(); dt=time.perf_counter()-t0\n"
            "print(json.dumps({'updates_per_s':100/dt,'ms_per_step':10*dt}))\n") % (os.path.join(ROOT, "tools"), ROOT)
    try:
        from baseline import ref_import

        if not ref_import.available():
            return {"unavailable": "baseline/_ref did not travel"}
        res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
        if res.returncode != 0:
            return {"unavailable": res.stderr[-300:]}
        d = json.loads(res.stdout.strip().splitlines()[-1])
        d["what"] = ("UNMODIFIED reference learning.critic_update + soft_update, REDQ-10 B=256, host numpy replay + H2D per batch, "
                     "stock PyTorch-CUDA eager on this B200 (wall clock, 100 updates)")
        return d
    except Exception as e:  # noqa: BLE001
        return {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}


def time_pixel_sampling(W):
    from super_sac_b200 import learning_utils as lu

    B = W.cfg["B"]
    C, H, Wd = W.cfg["pixels"]
    ms = bl.timed_events(lambda k: lu.sample_move_and_augment(W.buffer, B, W.augmenter, 1.0, per=False), 50, warmup=5)
    nbytes = 2 * (B * C * H * Wd * 1 + B * C * H * Wd * 4)
    peaks = bl.measured_peaks()
    return {"ms": ms, "algorithmic_MB": nbytes / 1e6, "GBps": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peaks["hbm_gbs"],
            "note": "o and o1: u8 read + fp32 written; includes the index / shift draw and the small-array gather launches"}


def time_nstep_ring_sampling(W, n_frames=12_000):
    """the same DrQv2 batch drawn from the frame-deduplicated one-step ring (NStepReplayBuffer, n_step 3,
    3 stacked RGB frames): device bytes per stored transition and the time of sample_move_and_augment, next to the classic
    ring (two copies of the whole stack per transition, n-step returns built on the host)."""
    import super_sac_b200 as ssb
    from super_sac_b200 import learning_utils as lu

    B = W.cfg["B"]
    C, H, Wd = W.cfg["pixels"]
    k, cf = 3, C // 3
    rng = np.random.default_rng(5)
    nb = ssb.replay.NStepReplayBuffer(n_frames, n_step=3, gamma=0.99, frame_stack=k, device=W.device, validate=False)
    A = W.cfg["A"]
    t, frames = 0, None
    t0 = time.perf_counter()
    while t < n_frames:
        T = 500
        frames = rng.integers(0, 256, (T + 1, cf, H, Wd), dtype=np.uint8)
        stack = lambda i: {"pixels": np.concatenate([frames[max(i - j, 0)] for j in range(k - 1, -1, -1)], 0)}   # noqa: E731
        for i in range(T):
            nb.push(stack(i), rng.uniform(-1, 1, A).astype(np.float32), float(rng.standard_normal()), stack(i + 1), False,
                    terminate_traj=(i == T - 1))
        t += T
    torch.cuda.synchronize()
    push_us = (time.perf_counter() - t0) / t * 1e6
    ms = bl.timed_events(lambda k_: lu.sample_move_and_augment(nb, B, W.augmenter, 1.0, per=False), 50, warmup=5)
    nbytes = 2 * (B * C * H * Wd * 1 + B * C * H * Wd * 4)
    peaks = bl.measured_peaks()
    classic = 2 * C * H * Wd + 4 * A + 5
    return {"ms": ms, "algorithmic_MB": nbytes / 1e6, "GBps": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peaks["hbm_gbs"],
            "bytes_per_transition": nb.bytes_per_transition(), "classic_ring_bytes_per_transition": classic,
            "host_push_us_incl_frame_generation": push_us, "ring_frames": n_frames,
            "note": "one-step ring, every frame stored once; n-step return, next-state stack and done flag assembled by the sampler"}


def time_pixel_encoder(B=512, iters=10):
    """The native DrQ encoder (csrc/ssac_conv.cu behind nets.cnns.BigPixelEncoder) alone at the C4 geometry: tensor-core
    bound implicit GEMMs in 3xTF32.  Algorithmic FLOPs = the reference module's own (88.5 MFLOP per sample forward, SURVEY 8d;
    backward = 2x), not the pitch-layout work the kernels actually do."""
    import super_sac_b200 as ssb

    dev = torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(0)
    enc = ssb.nets.cnns.BigPixelEncoder((9, 84, 84), 50).to(dev)
    obs = torch.randint(0, 256, (B, 9, 84, 84), device=dev).float()     # 130 MB: larger than L2 together with the activations
    dout = torch.randn(B, 50, device=dev)
    flop = 2 * (41 * 41 * 32 * 81 + (39 * 39 + 37 * 37 + 35 * 35) * 32 * 288 + 39200 * 50) * B

    def fwd():
        with torch.no_grad():
            enc(obs)

    def fwd_bwd():
        enc.zero_grad(set_to_none=True)
        enc(obs).backward(dout)

    out = {"workload": f"BigPixelEncoder(9x84x84 -> 50), B={B}, fp32 via 3xTF32 tcgen05 implicit GEMMs"}
    peak = bl.measured_peaks()["bf16_tflops"]
    for name, fn, mult in (("forward", fwd, 1), ("forward_backward", fwd_bwd, 3)):
        ms = bl.timed_events(lambda k: fn(), iters, warmup=3)    # per call
        tf = mult * flop / (ms * 1e-3) / 1e12
        out[name] = {"us": ms * 1e3, "algorithmic_TFLOPs": tf, "frac_of_bf16_peak": tf / peak,
                     "frac_of_3xTF32_ceiling": tf / (peak / 6)}
    return out


def hbm_kernels():
    """Achieved GB/s of the HBM-bound kernels against the measured copy bandwidth (algorithmic bytes)."""
    from super_sac_b200 import _lib

    L, sp = _lib.lib(), _lib.stream_ptr()
    peaks = bl.measured_peaks()
    dev = torch.device("cuda", torch.cuda.current_device())
    rows = []
    C, HW, B, cap = 9, 84, 512, 20_000
    src = torch.empty((cap, C, HW, HW), dtype=torch.uint8, device=dev).random_(0, 256)
    dst = torch.empty((B, C, HW, HW), device=dev)
    idx = torch.randint(0, cap, (B,), device=dev)
    shift = torch.randint(0, 9, (B, 2), device=dev, dtype=torch.int32)
    ms = bl.timed_events(lambda k: L.gather_aug_u8(src.data_ptr(), dst.data_ptr(), idx.data_ptr(), shift.data_ptr(), None, B, C, HW, HW, 4, 1, B, sp), 100, 10)
    rows.append(("ssac_gather_aug_u8 (C4: 512 x 9x84x84 u8 -> f32)", B * C * HW * HW * 5 + B * 8, ms, "ring and output larger than L2"))
    del src, dst
    for name, n, note in (("C5 2 150 402 params", 2_150_402, "L2-resident, as in the real step"), ("64 Mi params", 1 << 26, "larger than L2: HBM bound")):
        p = torch.randn(n, device=dev); t = torch.randn(n, device=dev); g = torch.randn(n, device=dev)
        m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev); ctl = torch.zeros(2, dtype=torch.int32, device=dev)
        ms = bl.timed_events(lambda k: L.polyak(t.data_ptr(), p.data_ptr(), n, 0.005, sp), 50, 5)
        rows.append((f"ssac_polyak ({name})", 12 * n, ms, note))
        ms = bl.timed_events(lambda k: L.adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, ctl.data_ptr(), 3e-4, .9, .999, 1e-8, 0., None, 0., 0, sp), 50, 5)
        rows.append((f"ssac_adam_step ({name})", 28 * n, ms, note))
        del p, t, g, m, v
    return [{"kernel": nm, "algorithmic_MB": nb / 1e6, "us": ms * 1e3, "GBps": nb / ms / 1e6, "frac_of_hbm_peak": nb / ms / 1e6 / peaks["hbm_gbs"],
             "note": note} for nm, nb, ms, note in rows]


def run_sharded_parity(rank, world, device, dist):
    """Before any timing: the ensemble-sharded update (critics over the ranks, then SUNRISE members over the ranks) must
    equal the single-GPU update on the same scripted draws (rtol 1e-4)."""
    try:
        import sharded_check

        sharded_check.run_checks(rank, world, device)
        ok = torch.ones(1, device=device)
        msg = "ok"
    except Exception as e:  # noqa: BLE001
        ok = torch.zeros(1, device=device)
        msg = "failed on rank %d: %s: %s" % (rank, type(e).__name__, str(e)[:300])
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return "ok" if float(ok.item()) == 1.0 else (msg if msg != "ok" else "failed on another rank")


def time_sharded(cfg, args, device, rank, world, dist):
    """Multi-GPU single learner: ONE learner whose ensemble is partitioned over the ranks -- the N critics of
    a REDQ-style agent, or the E members of a SUNRISE-style ensemble -- with the exchange (target Q rows / member batches
    and values) inside every update.  Returns updates/s of that single learner (max over ranks)."""
    import super_sac_b200 as ssb
    from super_sac_b200 import graphed, parallel

    members = cfg["E"] > 1
    if members:
        lo, hi = parallel.enable_member_sharding(cfg["E"])
        over, seed, tseed = dict(E=hi - lo), 100 + rank, 0      # every rank samples its own members' batches
    else:
        lo, hi = parallel.enable_critic_sharding(cfg["N"])
        over, seed, tseed = dict(N=hi - lo), 1234, 1234         # identical Philox stream and actor replica on every rank
    try:
        W = bl.Workload(ssb, args.config, device, seed=seed, buffer_size=200_000, overrides=over, torch_seed=tseed)
        if not members:
            W.kw["target_critic_ensemble_n"] = cfg["M"]

        def upd():
            out = W.critic_update()
            W.polyak()
            return out

        mode = "cuda-graph replay (exchange captured)"
        per_call = 1
        utd = cfg.get("utd", 1)
        step = None
        if not members and utd > 1 and utd % cfg["target_delay"] == 0 and not args.no_pipeline:
            # like the headline: the UTD block as ONE graph, software pipelined -- the target side of update k+1 INCLUDING
            # its exchange over NVLink runs next to update k's backward (learning.py, lu.pipelined_updates)
            from super_sac_b200 import learning_utils as lu

            def utd_block():
                with lu.pipelined_updates():
                    for u in range(utd):
                        out = W.step(u)
                return out

            try:
                W.critic_update()          # creates the exchange sites outside the capture
                torch.cuda.synchronize()
                dist.barrier()
                step = graphed.GraphedCall(utd_block, warmup=1).replay
                per_call = utd
                mode = "cuda-graph replay of the UTD block (%d updates per graph; exchange captured, on the target side's stream)" % utd
            except Exception as e:  # noqa: BLE001
                torch.cuda.synchronize()
                print("[bench] sharded UTD block not capturable: %s: %s" % (type(e).__name__, str(e)[:300]), file=sys.stderr, flush=True)
                step = None
        if step is None:
            try:
                g = graphed.GraphedCall(upd)
                step = g.replay
            except Exception as e:  # noqa: BLE001  (capture of the exchange is driver / NCCL dependent)
                mode = "eager launches (graph capture of the exchange unavailable: %s)" % type(e).__name__
                step = upd
        steps = max(1, min(args.steps, 1000) // per_call)
        for _ in range(max(2, 10 // per_call)):
            step()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        n_units = cfg["E"] if members else cfg["N"]
        steps *= per_call
        return {"value": steps / (ms * 1e-3), "unit": "updates/s of ONE learner", "ms_per_step": ms / steps, "mode": mode,
                "partitioned": "members" if members else "critics", "exchange": parallel.exchange_name(),
                "per_rank": [parallel.local_range(n_units, world, r)[1] - parallel.local_range(n_units, world, r)[0] for r in range(world)],
                "exchanged_per_update": ("member batches [E,B,S+A] + target values [E*N,E,B] fp32" if members else
                                         "target Q rows [N,B] fp32 (%d B per critic)" % (4 * cfg["B"]))}
    finally:
        parallel.disable()


def time_full_sac_step(W, iters=30):
    """the full SAC step of main.py:380-543 -- UTD critic updates (+ their Polyak steps), then ONE actor
    update and ONE temperature update on the last batch -- captured as a single CUDA graph where possible."""
    import contextlib

    from super_sac_b200 import learning, learning_utils as lu

    cfg = W.cfg
    utd = cfg.get("utd", 1)

    def full_step():
        rds = None
        with lu.pipelined_updates() if utd > 1 else contextlib.nullcontext():
            for u in range(utd):
                _, rds = learning._critic_update_impl(**W.kw)
                if u % cfg["target_delay"] == 0:
                    W.polyak()
            return W.actor_and_alpha(rds)

    step, mode, _ = graph_or_eager(full_step)
    ms = bl.timed_events(lambda k: step(), iters, 3)
    return {"value": 1e3 / ms, "unit": "SAC steps/s", "ms_per_step": ms, "critic_updates_per_step": utd,
            "gradient_updates_per_sec": (utd + 1) * 1e3 / ms, "mode": mode,
            "what": "%d x (critic_update + Polyak every %d) + online_actor_update + alpha_update" % (utd, cfg["target_delay"])}


def dominant_kernel_roofline(args, cfg, W):
    """CUDA-event time of ONE launch of the config's dominant kernel, replayed back to back from a CUDA graph so that no
    launch gap is counted, against the measured peak.  MLP configs: the ensemble-critic forward (all E*N critics on one
    batch; operands L2-warm, as they are inside the real step).  Pixel config: the fused gather + shift + cast."""
    import super_sac_b200 as ssb
    from super_sac_b200 import _ops

    peaks = bl.measured_peaks()
    if cfg.get("pixels"):
        # the pixel update is dominated by the native encoder's tensor-core convolutions (conv_halo_kernel: 11 of the ~74
        # launches, ~46 % of the step); the HBM-bound pixel gather rides along as `hbm`
        enc = time_pixel_encoder(B=cfg["B"])
        r = hbm_kernels()[0]
        f = enc["forward"]
        return {"kernel": "native BigPixelEncoder forward (s2d + 4 x conv_halo_kernel + split-K FC; 3xTF32 tcgen05 implicit GEMMs)",
                "bound": "tensor", "achieved": f["algorithmic_TFLOPs"], "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": f["frac_of_bf16_peak"], "frac_of_3xTF32_ceiling": f["frac_of_3xTF32_ceiling"], "traffic": None,
                "us_per_launch": f["us"], "forward_backward": enc["forward_backward"], "peak_source": peaks["source"],
                "note": "fp32 parity needs 3xTF32 and the 32-channel layers give the MMAs N = 64 / 32 only: ncu has the tensor "
                        "pipe active 72-86 % of the convolution kernels' time (profiles/r2_15)",
                "hbm": {"kernel": r["kernel"], "achieved": r["GBps"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": r["frac_of_hbm_peak"], "us_per_launch": r["us"],
                        "algorithmic_bytes_per_launch": r["algorithmic_MB"] * 1e6}}
    ca = W.agent._critic_arena
    G, B, H, D = cfg["E"] * cfg["N"], cfg["B"], cfg["H"], cfg["S"] + cfg["A"]
    dev = ca.device
    X = torch.randn(B, D, device=dev)
    h1 = torch.empty(G, B, H, device=dev)
    h2 = torch.empty_like(h1)
    q = torch.empty(G, B, 1, device=dev)
    ms = bl.graph_time(lambda: _ops.mlp_forward(ca, 0, G, X, B, h1, h2, q, keep_hidden=True), per=20, iters=10)
    fl = bl.critic_flops(cfg)
    achieved = fl["fwd_group"] / (ms * 1e-3) / 1e12
    fused = cfg["H"] <= 256 and cfg["H"] % 16 == 0 and D <= 32 and ssb.get_mlp_impl() == "tcgen05"
    return {"kernel": ("mlp3_forward_kernel (fc1+fc2+fc3 of all %d critics in ONE launch, B=%d)" if fused else
                       "ensemble critic forward (%d nets, B=%d; grouped tcgen05 GEMM, one launch per layer)") % (G, B),
            "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
            "frac": achieved / peaks["bf16_tflops"], "frac_of_3xTF32_ceiling": achieved / (peaks["bf16_tflops"] / 6),
            "traffic": 3004416 if (fused and args.config == "redq") else None,
            "us_per_launch": ms * 1e3, "algorithmic_flops_per_launch": fl["fwd_group"], "peak_source": peaks["source"],
            "note": "fp32 parity needs 3xTF32 (three kind::tf32 MMAs per product at half the bf16 rate: the bf16 figure is 6x "
                    "out of reach by construction); traffic = dram bytes of one cold-cache ncu launch (profiles/), operands are "
                    "L2-resident inside the step"}


def _keep_stdout_for_the_json_line():
    """Everything that writes to fd 1 during the run (NCCL's version banner, diagnostics of the parity check) goes to
    stderr; the ONE JSON line is written to the original stdout."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(saved, "w")
    real_print = print

    def emit(line):
        out.write(line + "\n")
        out.flush()

    return emit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="redq", choices=sorted(CONFIGS))
    ap.add_argument("--mlp-impl", default="tcgen05", choices=["tcgen05", "ffma"])
    ap.add_argument("--no-overlap", action="store_true", help="serialise the independent branches of the update (A/B switch)")
    ap.add_argument("--no-pdl", action="store_true", help="plain stream-ordered launches instead of programmatic dependent launch")
    ap.add_argument("--no-pipeline", action="store_true", help="one graph per update instead of the pipelined UTD block (A/B switch)")
    ap.add_argument("--skip-secondary", action="store_true", help="headline config only (no other configs / kernels / GPU-reference arm)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    global _emit
    _emit = _keep_stdout_for_the_json_line()
    if args.impl == "reference":
        run_reference_arm(args, cfg, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (super_sac_b200 has no CPU path); "
                         "use --impl reference for the CPU arm")
    run_gpu_arm(args, cfg, rank, world, local_rank)


if __name__ == "__main__":
    main()
