"""Tensor-level wrappers over the C ABI (pointer marshalling only -- no arithmetic happens in Python)."""
import ctypes

import torch

from . import _lib


def _p(t):
    return None if t is None else t.data_ptr()


def check_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.SsacError("super_sac_b200 runs on sm_100 CUDA devices only (got a CPU tensor); there is no CPU path")
    for t in tensors:
        if t is not None:
            _lib.require_device(t.device.index if t.device.index is not None else torch.cuda.current_device())
            break


def _bwd_ws(G, B, H, device):
    # a fresh tensor per call: the caching allocator makes this free, and under graph capture it comes from the
    # graph's private pool (a cached, later-resized workspace would leave captured graphs with a stale pointer)
    return torch.empty(_lib.lib().mlp_backward_ws(G, B, H), dtype=torch.float32, device=device)


def mlp_forward(arena, g0, G, x, B, h1, h2, y, ldx=None, x_gs=0, net_index=None, impl=0, keep_hidden=True):
    """y[g] = MLP_{g0+g}(x[g]);  x is [B, ldx] (x_gs = 0, shared) or [G, B, ldx].  keep_hidden=False: h1 / h2 are
    scratch (the single-kernel forward skips writing them)."""
    check_cuda(x, y)
    W1, b1, W2, b2, W3, b3 = arena.ptrs(g0)
    _lib.lib().mlp_forward(W1, b1, W2, b2, W3, b3, _p(net_index), G, arena.D, arena.H, arena.O, x.data_ptr(),
                           arena.D if ldx is None else ldx, x_gs, B, _p(h1), _p(h2), int(bool(keep_hidden)),
                           y.data_ptr(), impl, _lib.stream_ptr())


def mlp_backward(arena, g0, G, x, B, h1, h2, dy, ldx=None, x_gs=0, dh2_extra=None, extra_scale=0.0, want_dw=True,
                 accumulate=False, dx=None, lddx=0, net_index=None, impl=0):
    check_cuda(x, h1, h2)
    W1, _, W2, _, W3, _ = arena.ptrs(g0)
    if want_dw:
        gW1, gb1, gW2, gb2, gW3, gb3 = arena.ptrs(g0, grad=True)
    else:
        gW1 = gb1 = gW2 = gb2 = gW3 = gb3 = None
    ws = _bwd_ws(G, B, arena.H, x.device)
    _lib.lib().mlp_backward(W1, W2, W3, _p(net_index), G, arena.D, arena.H, arena.O, x.data_ptr(),
                            arena.D if ldx is None else ldx, x_gs, B, h1.data_ptr(), h2.data_ptr(), _p(dy),
                            _p(dh2_extra), float(extra_scale), gW1, gb1, gW2, gb2, gW3, gb3, int(bool(accumulate)),
                            _p(dx), lddx, ws.data_ptr(), impl, _lib.stream_ptr())


def polyak_ranges(target_flat, source_flat, ranges, tau):
    L = _lib.lib()
    s = _lib.stream_ptr()
    for off, n in ranges:
        L.polyak(target_flat.data_ptr() + 4 * off, source_flat.data_ptr() + 4 * off, n, float(tau), s)


def gather_rows(srcs, dsts, row_elems, dst_ld, modes, idx, B):
    n = len(srcs)
    L = _lib.lib()
    L.gather_rows(_lib.host_array(ctypes.c_void_p, [s.data_ptr() for s in srcs]),
                  _lib.host_array(ctypes.c_void_p, [d.data_ptr() for d in dsts]),
                  _lib.host_array(ctypes.c_int64, list(row_elems)), _lib.host_array(ctypes.c_int64, list(dst_ld)),
                  _lib.host_array(ctypes.c_int32, list(modes)), n, idx.data_ptr(), B, _lib.stream_ptr())
