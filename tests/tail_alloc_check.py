"""Stand-alone check run by test_cuda_kernels.test_tma_operands_at_allocation_tail in a subprocess with
PYTORCH_NO_CUDA_MEMORY_CACHING=1 (every tensor its own cudaMalloc): TMA operands that end exactly at the end of their
allocation.  A TMA box hanging over the end of such a tensor faults on B200 although the rows are declared out of bounds
(csrc/ssac_mlp_tc.cu make_map), so make_map must refuse the map and the kernel must stage that operand through registers."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from super_sac_b200 import _arena, _ops

which = sys.argv[1]
G, D, H, O, B = 2, 4, 32, 1, 16
ar = _arena.MLPArena(G, D, H, O, "cuda")
n = ar.numel
big = torch.zeros(2 << 20, dtype=torch.uint8, device="cuda")
torch.manual_seed(0)
ref_flat = torch.randn(n, device="cuda") * 0.1
ar.flat.copy_(ref_flat)
X = torch.randn(B, D, device="cuda")
h1 = torch.empty(G, B, H, device="cuda")
h2 = torch.empty_like(h1)
y = torch.empty(G, B, 1, device="cuda")
ni = torch.tensor([1, 0], dtype=torch.int32, device="cuda")
_ops.mlp_forward(ar, 0, G, X, B, h1, h2, y, ldx=D, net_index=ni, keep_hidden=True)
torch.cuda.synchronize()
want = y.clone()
if which == "arena_tail":          # the weights end at the last byte of a cudaMalloc block
    flat = big.view(torch.float32)[-n:]
    flat.copy_(ref_flat)
    ar.flat = flat
    ar._make_views()
if which == "h_tail":              # the saved activations do
    hb = big.view(torch.float32)[-2 * G * B * H:]
    h1, h2 = hb[: G * B * H].view(G, B, H), hb[G * B * H:].view(G, B, H)
for keep in (False, True):
    y.zero_()
    _ops.mlp_forward(ar, 0, G, X, B, h1, h2, y, ldx=D, net_index=ni, keep_hidden=keep)
    torch.cuda.synchronize()
    assert torch.allclose(y, want, rtol=1e-5, atol=1e-6), (which, keep, float((y - want).abs().max()))
print("ok", which)
