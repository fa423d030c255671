import sys, os, copy, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from itertools import chain
import numpy as np, torch
import cuda_util as cu
import super_sac_b200 as ssb
from super_sac_b200 import augmentations, graphed, learning, learning_utils as lu, nets

def build():
    ssb.manual_seed(5); torch.manual_seed(5)
    agent = ssb.Agent(act_space_size=6, encoder=cu.IdentityEncoder(17), actor_network_cls=nets.mlps.ContinuousStochasticActor,
                      critic_network_cls=nets.mlps.ContinuousCritic, ensemble_size=1, num_critics=10, hidden_size=256,
                      auto_rescale_targets=False, log_std_low=-5.0, log_std_high=2.0)
    agent.to("cuda"); target = copy.deepcopy(agent)
    rng = np.random.default_rng(0); n = 4096
    buf = ssb.replay.ReplayBuffer(n, device="cuda")
    buf.load_experience({"obs": rng.standard_normal((n, 17), dtype=np.float32)}, rng.uniform(-1, 1, (n, 6)).astype(np.float32),
                        rng.standard_normal(n, dtype=np.float32), {"obs": rng.standard_normal((n, 17), dtype=np.float32)},
                        rng.uniform(size=n) < 0.05)
    c_opt = torch.optim.Adam(chain(*(c.parameters() for c in agent.critics)), lr=3e-4)
    e_opt = torch.optim.Adam(agent.encoder.parameters(), lr=1e-4)
    la = [torch.tensor([-2.3], device="cuda", requires_grad=True)]
    aug = augmentations.AugmentationSequence([augmentations.IdentityAug(256)])
    def block(pipelined):
        outs = []
        with (lu.pipelined_updates() if pipelined else contextlib.nullcontext()):
            for k in range(6):
                logs, rds = learning._critic_update_impl(
                    buffer=buf, agent=agent, target_agent=target, critic_optimizer=c_opt, encoder_optimizer=e_opt,
                    log_alphas=la, batch_size=256, gamma=0.99, critic_clip=None, encoder_clip=None,
                    target_critic_ensemble_n=2, weighted_bellman_temp=None, weight_type=None, pop=False, augmenter=aug,
                    encoder_lambda=0.0, random_process=None, noise_clip=None, aug_mix=0.0)
                outs.append(logs)
                if k % 2 == 0:
                    lu.soft_update(target.critics[0], agent.critics[0], 0.005)
        return outs
    return agent, block

def show(name, outs):
    for k, l in enumerate(outs):
        d = dict(l.fetch(keep=True)) if hasattr(l, "fetch") else dict(l)
        print(name, k, {kk.split("/")[-1]: round(float(v), 6) for kk, v in d.items()})

for pipelined in (False, True):
    a, block = build()
    for _ in range(3):
        outs = block(pipelined)
    show("eager pipe=%d" % pipelined, outs)
    a, block = build()
    holder = {}
    def fn():
        holder["o"] = block(pipelined)
        return holder["o"][-1]
    g = graphed.GraphedCall(fn, warmup=2)
    g.replay(); torch.cuda.synchronize()
    show("graph pipe=%d" % pipelined, holder["o"])
