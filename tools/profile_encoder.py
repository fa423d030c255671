"""One forward + backward of the native DrQ encoder at B = 512 (for an ncu launch list / --set full capture):
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_encoder.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from super_sac_b200.nets import cnns  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
enc = cnns.BigPixelEncoder((9, 84, 84), 50).cuda()
obs = torch.randint(0, 256, (B, 9, 84, 84), device="cuda").float()
dout = torch.randn(B, 50, device="cuda")
for _ in range(iters):
    enc.zero_grad(set_to_none=True)
    enc(obs).backward(dout)
torch.cuda.synchronize()
