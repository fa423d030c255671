"""Is the divergence of two Adam trajectories of the encoder a bug or sensitivity?  Three trajectories on the same problem:
CPU PyTorch, GPU PyTorch/cuDNN (fp32, TF32 off), GPU native kernels; pairwise worst parameter difference over the steps."""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from super_sac_b200.nets import cnns  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(21)
base = cnns.BigPixelEncoder((3, 20, 20), 10)
with torch.no_grad():
    for p in base.parameters():
        p.add_(0.05 * torch.randn_like(p))
rng = np.random.default_rng(21)
obs = torch.as_tensor(rng.integers(0, 256, (24, 3, 20, 20)).astype(np.float32))
tgt = torch.as_tensor(rng.uniform(-0.8, 0.8, (24, 10)).astype(np.float32))
lr = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-3
nets = {"cpu": copy.deepcopy(base), "cudnn": copy.deepcopy(base).cuda(), "native": copy.deepcopy(base).cuda()}
opts = {k: torch.optim.Adam(v.parameters(), lr=lr) for k, v in nets.items()}
data = {"cpu": (obs, tgt), "cudnn": (obs.cuda(), tgt.cuda()), "native": (obs.cuda(), tgt.cuda())}
torch.set_num_threads(1)
for step in range(1, 61):
    for k in nets:
        os.environ["SSAC_ENCODER_IMPL"] = "torch" if k == "cudnn" else "native"
        opts[k].zero_grad()
        o, t = data[k]
        loss = ((nets[k](o) - t) ** 2).mean()
        loss.backward()
        opts[k].step()
    if step in (1, 2, 5, 10, 20, 40, 60):
        def diff(a, b):
            return max(float((pa.detach().cpu() - pb.detach().cpu()).abs().max()) for pa, pb in zip(nets[a].parameters(), nets[b].parameters()))
        print(f"step {step:3d}: cpu-cudnn {diff('cpu', 'cudnn'):.3e}  cpu-native {diff('cpu', 'native'):.3e}  cudnn-native {diff('cudnn', 'native'):.3e}  loss {float(loss):.5f}")
