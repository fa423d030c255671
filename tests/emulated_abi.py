"""A numpy stand-in for the handful of C-ABI entry points the SAC-Discrete updates call.  TEST INFRASTRUCTURE ONLY.

CPU tensors' ``data_ptr()`` are host addresses, so the emulation reads and writes the caller's buffers through the very
pointers, strides and argument order the product hands to libssac_b200.so: tests/test_discrete_host.py runs
super_sac_b200.discrete on CPU tensors against it and compares with the reference's golden vectors -- a check of the
HOST logic (argument order, strides, loss normalisation, log plumbing), not of the kernels; the kernels are checked on the
GPU (tests/test_discrete_parity.py).  Semantics follow include/ssac_b200.h.  Never imported by the product.
"""
import ctypes
import math

import numpy as np


def _f32(ptr, n):
    return np.ctypeslib.as_array((ctypes.c_float * int(n)).from_address(ptr))


def _i32(ptr, n):
    return np.ctypeslib.as_array((ctypes.c_int32 * int(n)).from_address(ptr))


def _rows(ptr, B, D, ld):
    """[B, D] view of a row-major matrix with row stride ld."""
    flat = _f32(ptr, (B - 1) * ld + D)
    return np.lib.stride_tricks.as_strided(flat, shape=(B, D), strides=(4 * ld, 4))


def _softmax_stats(z):
    m = z.max(-1, keepdims=True)
    lp = z - m - np.log(np.exp(z - m).sum(-1, keepdims=True, dtype=np.float32))
    return np.exp(lp), lp


class EmulatedLib:
    def __init__(self):
        self.calls = []

    def default_mlp_impl(self):
        return 1

    def mlp_backward_ws(self, G, B, H):
        return 16

    # ---- grouped MLP ------------------------------------------------------------------------------------------------
    def _nets(self, net_index, G):
        return list(range(G)) if not net_index else [int(v) for v in _i32(net_index, G)]

    def mlp_forward(self, W1, b1, W2, b2, W3, b3, net_index, G, D, H, O, x, ldx, x_gs, B, h1, h2, keep, y, impl, stream):
        self.calls.append("mlp_forward")
        assert x_gs == 0
        X = _rows(x, B, D, ldx)
        for g, n in enumerate(self._nets(net_index, G)):
            w1, bb1 = _f32(W1 + 4 * n * H * D, H * D).reshape(H, D), _f32(b1 + 4 * n * H, H)
            w2, bb2 = _f32(W2 + 4 * n * H * H, H * H).reshape(H, H), _f32(b2 + 4 * n * H, H)
            w3, bb3 = _f32(W3 + 4 * n * O * H, O * H).reshape(O, H), _f32(b3 + 4 * n * O, O)
            a1 = np.maximum(X @ w1.T + bb1, 0)
            a2 = np.maximum(a1 @ w2.T + bb2, 0)
            if h1:
                _f32(h1 + 4 * g * B * H, B * H)[:] = a1.ravel()
                _f32(h2 + 4 * g * B * H, B * H)[:] = a2.ravel()
            _f32(y + 4 * g * B * O, B * O)[:] = (a2 @ w3.T + bb3).ravel()

    def mlp_backward(self, W1, W2, W3, net_index, G, D, H, O, x, ldx, x_gs, B, h1, h2, dy, dh2_extra, extra_scale, gW1, gb1,
                     gW2, gb2, gW3, gb3, accumulate, dx, lddx, ws, impl, stream):
        self.calls.append("mlp_backward")
        assert x_gs == 0 and ws
        X = _rows(x, B, D, ldx)
        for g, n in enumerate(self._nets(net_index, G)):
            w1 = _f32(W1 + 4 * n * H * D, H * D).reshape(H, D)
            w2 = _f32(W2 + 4 * n * H * H, H * H).reshape(H, H)
            w3 = _f32(W3 + 4 * n * O * H, O * H).reshape(O, H)
            a1 = _f32(h1 + 4 * g * B * H, B * H).reshape(B, H)
            a2 = _f32(h2 + 4 * g * B * H, B * H).reshape(B, H)
            d_y = _f32(dy + 4 * g * B * O, B * O).reshape(B, O) if dy else np.zeros((B, O), np.float32)
            dh2 = d_y @ w3
            if dh2_extra:
                dh2 = dh2 + np.float32(extra_scale) * _f32(dh2_extra + 4 * g * B * H, B * H).reshape(B, H)
            dz2 = dh2 * (a2 > 0)
            dz1 = (dz2 @ w2) * (a1 > 0)
            if gW1:
                for ptr, val, cnt in ((gW1, dz1.T @ X, H * D), (gb1, dz1.sum(0), H), (gW2, dz2.T @ a1, H * H),
                                      (gb2, dz2.sum(0), H), (gW3, d_y.T @ a2, O * H), (gb3, d_y.sum(0), O)):
                    out = _f32(ptr + 4 * n * cnt, cnt)
                    out[:] = (out if accumulate else 0) + val.astype(np.float32).ravel()
            if dx:
                _rows(dx + 4 * g * B * lddx, B, D, lddx)[:] = dz1 @ w1

    # ---- SAC-Discrete heads (csrc/ssac_discrete.cu) ----------------------------------------------------------------
    def discrete_value(self, logits, q_t, M, B, A, log_alpha, v, ent, stream):
        self.calls.append("discrete_value")
        p, lp = _softmax_stats(_f32(logits, B * A).reshape(B, A))
        q = _f32(q_t, M * B * A).reshape(M, B, A).min(0)
        alpha = np.exp(_f32(log_alpha, 1)[0])
        _f32(v, B)[:] = (p * (q - alpha * lp)).sum(-1)
        if ent:
            _f32(ent, 1)[0] += (alpha * lp).mean()

    def discrete_gather_q(self, q, act, G, B, A, out, stream):
        self.calls.append("discrete_gather_q")
        a = _f32(act, B).astype(np.int64)
        _f32(out, G * B).reshape(G, B)[:] = _f32(q, G * B * A).reshape(G, B, A)[:, np.arange(B), a]

    def discrete_critic_loss_seed(self, q, N, B, A, act, y, w, imp, popart, pop, E, n_total, dy, loss, stream):
        self.calls.append("discrete_critic_loss_seed")
        n_total = n_total or N
        a = _f32(act, B).astype(np.int64)
        qs = _f32(q, N * B * A).reshape(N, B, A)[:, np.arange(B), a]
        pw, pb = (_f32(popart, 4)[2], _f32(popart, 4)[3]) if (popart and pop) else (np.float32(1), np.float32(0))
        td = _f32(y, B)[None, :] - (pw * qs + pb)
        ww = (_f32(w, B) if w else 1.0) * (_f32(imp, B) if imp else 1.0) * np.ones(B, np.float32)
        inv = 1.0 / (B * E * n_total)
        d = np.zeros((N, B, A), np.float32)
        d[:, np.arange(B), a] = -2.0 * ww * td * pw * inv
        _f32(dy, N * B * A)[:] = d.ravel()
        ls = _f32(loss, 2)
        ls[0] += (ww * td * td).sum() * inv
        ls[1] = td[N - 1].mean()

    def discrete_actor_seed(self, logits, q, N, B, A, log_alpha, popart, pop, E, dlogits, loss, stream):
        self.calls.append("discrete_actor_seed")
        p, lp = _softmax_stats(_f32(logits, B * A).reshape(B, A))
        vals = _f32(q, N * B * A).reshape(N, B, A).min(0)
        if popart and pop:
            vals = _f32(popart, 4)[2] * vals + _f32(popart, 4)[3]
        alpha = np.exp(_f32(log_alpha, 1)[0])
        g = vals - alpha * lp
        f = (p * g).sum(-1, keepdims=True)
        _f32(dlogits, B * A)[:] = ((-1.0 / (E * B)) * p * (g - f)).ravel()
        _f32(loss, 1)[0] += -f.sum() / (E * B)

    def discrete_neg_entropy(self, logits, B, A, out, stream):
        self.calls.append("discrete_neg_entropy")
        p, lp = _softmax_stats(_f32(logits, B * A).reshape(B, A))
        _f32(out, B)[:] = (p * lp).sum(-1)

    def discrete_advantage(self, logits, E, q, N, B, A, act, popart, adv, mask, priority, stream):
        self.calls.append("discrete_advantage")
        pm = np.mean([_softmax_stats(_f32(logits + 4 * e * B * A, B * A).reshape(B, A))[0] for e in range(E)], axis=0)
        mq = _f32(q, N * B * A).reshape(N, B, A).min(0)
        if popart:
            mq = _f32(popart, 4)[2] * mq + _f32(popart, 4)[3]
        a = _f32(act, B).astype(np.int64)
        ad = mq[np.arange(B), a] - (pm * mq).sum(-1)
        if adv:
            _f32(adv, B)[:] = ad
        if mask:
            _f32(mask, B)[:] = ad >= 0
        if priority:
            np.ctypeslib.as_array((ctypes.c_double * B).from_address(priority))[:] = np.maximum(ad, 0).astype(np.float64) + 1e-4

    def discrete_bc_seed(self, logits, act, mask, B, A, E, dlogits, loss, stream):
        self.calls.append("discrete_bc_seed")
        p, lp = _softmax_stats(_f32(logits, B * A).reshape(B, A))
        a = _f32(act, B).astype(np.int64)
        m = _f32(mask, B) if mask else np.ones(B, np.float32)
        onehot = np.zeros((B, A), np.float32)
        onehot[np.arange(B), a] = 1
        _f32(dlogits, B * A)[:] = ((-m / (B * E))[:, None] * (onehot - p)).ravel()
        _f32(loss, 1)[0] += -(m * lp[np.arange(B), a]).sum() / B

    # ---- shared with the continuous path (csrc/ssac_elementwise.cu) ------------------------------------------------
    def td_target(self, q_t, M, B, logp, log_alpha, r, d, gamma, popart, popart_ctl, pop, beta, min_steps, y, logs, stream):
        self.calls.append("td_target")
        assert not logp
        v = _f32(q_t, M * B).reshape(M, B).min(0)
        if popart:
            st, ctl = _f32(popart, 4), _i32(popart_ctl, 2)
            mu, nu, pw, pb = (np.float32(x) for x in st)
            sigma = np.clip(np.sqrt(nu - mu * mu) + np.float32(1e-5), 1e-4, 1e6).astype(np.float32)
            if pop:
                v = sigma * (pw * v + pb) + mu
        yy = _f32(r, B) + np.float32(gamma) * (1 - _f32(d, B)) * v
        if popart:
            t = int(ctl[0]) + 1
            beta_t = beta / (1.0 - (1.0 - beta) ** t)
            nmu = np.float32((1 - beta_t) * mu + beta_t * yy.mean())
            nnu = np.float32((1 - beta_t) * nu + beta_t * (yy * yy).mean())
            nsig = np.clip(np.sqrt(nnu - nmu * nmu) + np.float32(1e-5), 1e-4, 1e6).astype(np.float32)
            stable = (t > min_steps) and ((1 - sigma) / nsig <= 0.1)
            if stable:
                pw, pb = pw * (sigma / nsig), (sigma * pb + mu - nmu) / nsig
            st[:] = (nmu, nnu, pw, pb)
            ctl[:] = (t, int(stable))
            yy = (yy - nmu) / nsig
        _f32(y, B)[:] = yy
        _f32(logs, 3)[:] = (yy.mean(), yy.std(ddof=1), 0.0)

    def backup_weights(self, q, E, N, B, temperature, kind, w, logs, stream):
        self.calls.append("backup_weights")
        assert kind == 0
        std = _f32(q, E * N * B).reshape(E, N, B).min(1).std(0, ddof=1)
        ww = 1.0 / (1.0 + np.exp(std * temperature)) + 0.5
        _f32(w, B)[:] = ww
        _f32(logs, 4)[:] = (ww.mean(), ww.max(), ww.min(), ww.std(ddof=1))

    def dr3_dot(self, f, f1, N, B, H, out, stream):
        self.calls.append("dr3_dot")
        _f32(out, 1)[0] = (_f32(f, N * B * H) * _f32(f1, N * B * H)).sum() / (N * B)

    def sumsq(self, x, n, out, accumulate, stream):
        o = _f32(out, 1)
        o[0] = (o[0] if accumulate else 0) + (_f32(x, n).astype(np.float64) ** 2).sum()

    def adam_step(self, p, g, m, v, n, ctl, lr, b1, b2, eps, wd, gnorm_sq, max_norm, wb, stream):
        self.calls.append("adam_step")
        P, G, M_, V, c = _f32(p, n), _f32(g, n), _f32(m, n), _f32(v, n), _i32(ctl, 2)
        if gnorm_sq and max_norm > 0:
            G *= np.float32(min(1.0, max_norm / (math.sqrt(_f32(gnorm_sq, 1)[0]) + 1e-6)))
        t = int(c[0]) + 1
        ge = G + np.float32(wd) * P if wd else G
        M_[:] = M_ + np.float32(1 - b1) * (ge - M_)
        V[:] = V * np.float32(b2) + np.float32(1 - b2) * ge * ge
        denom = np.sqrt(V) / np.float32(math.sqrt(1 - b2**t)) + np.float32(eps)
        P[:] = P - np.float32(lr / (1 - b1**t)) * (M_ / denom)
        c[0] = t

    def alpha_step(self, log_alpha, logp, B, target_entropy, state, ctl, lr, b1, b2, eps, logs, stream):
        self.calls.append("alpha_step")
        la, st, c = _f32(log_alpha, 1), _f32(state, 2), _i32(ctl, 2)
        tt = _f32(logp, B) + np.float32(target_entropy)
        loss, grad = -(la[0] * tt).mean(), -tt.mean()
        t = int(c[0]) + 1
        st[0] = st[0] + (1 - b1) * (grad - st[0])
        st[1] = st[1] * b2 + (1 - b2) * grad * grad
        la[0] -= (lr / (1 - b1**t)) * st[0] / (math.sqrt(st[1]) / math.sqrt(1 - b2**t) + eps)
        c[0] = t
        _f32(logs, 2)[:] = (loss, math.exp(la[0]))
