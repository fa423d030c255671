#!/usr/bin/env python
"""SUNRISE (C3: 5 members x 2 critics, B=256, sunrise weights T=20) with the members sharded over the ranks
(SURVEY 8e): updates/s of ONE learner, next to the single-GPU figure of tools/bench_configs.py.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29540 \
        tools/bench_sunrise_sharded.py [--steps 300]"""
import argparse
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import cuda_util as cu  # noqa: E402
import super_sac_b200 as ssb  # noqa: E402
from super_sac_b200 import augmentations, learning, learning_utils as lu, nets, parallel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=300)
args = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
E, N, S, A, H, B = 5, 2, 17, 6, 256, 256
lo, hi = parallel.enable_member_sharding(E)
ssb.manual_seed(100 + rank)          # every rank samples its own members' batches
torch.manual_seed(0)
agent = ssb.Agent(act_space_size=A, encoder=cu.IdentityEncoder(S), actor_network_cls=nets.mlps.ContinuousStochasticActor,
                  critic_network_cls=nets.mlps.ContinuousCritic, ensemble_size=hi - lo, num_critics=N, hidden_size=H,
                  auto_rescale_targets=False, log_std_low=-5.0, log_std_high=2.0)
agent.to(dev)
target = copy.deepcopy(agent)
c_opt, a_opt, e_opt, las, _ = cu.optimizers(agent, dict(E=hi - lo))
buf = ssb.replay.ReplayBuffer(200_000, device=dev)
s, a, r, s1, d = bench.synthetic_transitions(dict(S=S, A=A), 200_000)
buf.load_experience({"obs": s}, a, r, {"obs": s1}, d)
kw = dict(buffer=buf, agent=agent, target_agent=target, critic_optimizer=c_opt, encoder_optimizer=e_opt, log_alphas=las,
          batch_size=B, gamma=0.99, critic_clip=None, encoder_clip=None, target_critic_ensemble_n=N,
          weighted_bellman_temp=20.0, weight_type="sunrise", pop=False,
          augmenter=augmentations.AugmentationSequence([augmentations.IdentityAug(B)]), encoder_lambda=0.0,
          random_process=None, noise_clip=None, aug_mix=0.0)


def upd():
    learning._critic_update_impl(**kw)
    for ac, tc in zip(agent.critics, target.critics):
        lu.soft_update(tc, ac, 0.005)


for _ in range(5):
    upd()
torch.cuda.synchronize()
mode = "eager launches"
step = upd
try:
    g = torch.cuda.CUDAGraph()
    from super_sac_b200 import _logs
    with _logs.deferred():
        with torch.cuda.graph(g):
            upd()
    step, mode = g.replay, "cuda-graph replay (NCCL all-gathers captured)"
except Exception as e:  # noqa: BLE001
    mode = "eager launches (capture unavailable: %s)" % type(e).__name__
for _ in range(5):
    step()
dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"config": "sunrise (C3), members sharded", "n_gpus": world, "value": 1e3 / float(ms), "unit": "updates/s of ONE learner",
                      "ms_per_step": float(ms), "mode": mode, "members_per_rank": [parallel.local_range(E, world, r)[1] -
                                                                                    parallel.local_range(E, world, r)[0] for r in range(world)],
                      "collectives_per_update": "all-gather of the member batches [E,B,S+A] + all-gather of the target values [E*N,E,B]"}))
sys.stdout.flush()
torch.cuda.synchronize()
os._exit(0)   # a captured NCCL graph can make the orderly process-group shutdown hang; the measurement is done
