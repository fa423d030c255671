"""The drift test's exact problem instance: CPU PyTorch vs GPU cuDNN vs native (torch Adam) vs native (fused Adam)."""
import copy, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from super_sac_b200 import _encoder_opt, nets
from super_sac_b200.nets import cnns
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

class Enc(nets.Encoder):
    def __init__(self):
        super().__init__()
        self.net = cnns.BigPixelEncoder((3, 20, 20), 10)
    @property
    def embedding_dim(self):
        return 10
    def forward(self, obs_dict):
        return self.net(obs_dict["obs"])

torch.manual_seed(21)
cpu = Enc()
with torch.no_grad():
    for p in cpu.parameters():
        p.add_(0.05 * torch.randn_like(p))
lr = 1e-3
rng = np.random.default_rng(21)
obs = torch.as_tensor(rng.integers(0, 256, (24, 3, 20, 20)).astype(np.float32))
tgt = torch.as_tensor(rng.uniform(-0.8, 0.8, (24, 10)).astype(np.float32))
ns = {"cpu": cpu, "cudnn": copy.deepcopy(cpu).cuda(), "native_torchadam": copy.deepcopy(cpu).cuda(), "native_fused": copy.deepcopy(cpu).cuda()}
os_ = {k: torch.optim.Adam(v.parameters(), lr=lr) for k, v in ns.items()}
torch.set_num_threads(1)
for step in range(1, 21):
    for k, net in ns.items():
        os.environ["SSAC_ENCODER_IMPL"] = "torch" if k == "cudnn" else "native"
        o, t = (obs, tgt) if k == "cpu" else (obs.cuda(), tgt.cuda())
        os_[k].zero_grad()
        ((net({"obs": o}) - t) ** 2).mean().backward()
        if k == "native_fused":
            assert _encoder_opt.fused_step(net, os_[k], None) is net.net
        else:
            os_[k].step()
    if step in (1, 5, 10, 15, 20):
        d = lambda a, b: max(float((pa.detach().cpu() - pb.detach().cpu()).abs().max()) for pa, pb in zip(ns[a].net.parameters(), ns[b].net.parameters()))
        print(f"step {step:2d}: cpu-cudnn {d('cpu','cudnn'):.2e}  cpu-native(torch Adam) {d('cpu','native_torchadam'):.2e}  cpu-native(fused) {d('cpu','native_fused'):.2e}  fused-torchAdam {d('native_fused','native_torchadam'):.2e}")
