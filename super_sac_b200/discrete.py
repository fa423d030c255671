"""SAC-Discrete updates: the ``discrete=True`` branches of reference learning.py:18-141 (critic),
:344-421 (actor), :222-263 (temperature) and learning_utils.py:322-328 (TD target), :373-376 (sunrise weights).

Same flat arenas, grouped MLP launches (ssac_mlp_forward / ssac_mlp_backward with O = number of actions: every critic
net emits its whole Q row), fused Adam and device-side logs as the continuous path in learning.py; the categorical
arithmetic around the networks (softmax, expectation over actions, gather at the taken action and its scatter back into
the dense output gradient) runs in the ssac_discrete_* kernels (csrc/ssac_discrete.cu).  This is the first correct path:
eager launches, members in series, no CUDA-graph capture / cross-update pipelining yet.  The offline (AFBC) actor update
with the indirect advantage filter (adv_estimator.py:45-56) and the priority refresh are here too.  Not implemented
(raise): softmax Bellman weights (they draw Categorical samples), the invariance regularisers, sharded ensembles.
"""
import random

import torch

from . import _arena, _encoder_opt, _lib, _logs, _ops, _rng, parallel
from . import learning_utils as lu


def _rows(rep):
    """(tensor, row stride) of a [B, S] fp32 device matrix the kernels can read in place (unit column stride)."""
    rep = rep.detach()
    if rep.dtype != torch.float32 or rep.dim() != 2 or rep.stride(1) != 1:
        rep = rep.float().contiguous()
    _ops.check_cuda(rep)
    return rep, rep.stride(0)


def _forward(arena, g0, G, X, ldx, B, net_index=None, keep=False):
    """y [G,B,O] of nets g0..g0+G (or the ``net_index`` subset relative to g0) on the rows of X."""
    dev = X.device
    h1 = torch.empty((G, B, arena.H), dtype=torch.float32, device=dev)
    h2 = torch.empty_like(h1)
    y = torch.empty((G, B, arena.O), dtype=torch.float32, device=dev)
    _ops.mlp_forward(arena, g0, G, X, B, h1, h2, y, ldx=ldx, net_index=net_index, keep_hidden=keep)
    return (y, h1, h2) if keep else y


def _actions(a, B):
    """The taken actions as a dense float [B] vector holding the indices (replay row [B,1], ``a.long()`` in the reference)."""
    return a.reshape(B).float().contiguous()


def _check(agent):
    if not agent.discrete:
        raise ValueError("discrete=True needs an Agent built with discrete=True")
    if parallel.is_sharded() or parallel.members_sharded():
        raise NotImplementedError("sharded ensembles cover the continuous path only")
    _ops.check_cuda(agent._critic_arena.flat)


def compute_td_targets(logs, replay_dict, agent, target_agent, ensemble_idx, ensemble_n, log_alphas, pop, gamma):
    """learning_utils.py:298-354 with discrete=True: y = r + gamma (1-d) E_{a~pi(s1)}[min_M Q_target(s1, a) - alpha log pi(a|s1)],
    PopArt as in the continuous branch.  Returns (y [B,1], (s1_rep, None)): the reference hands back the policy's
    probabilities in the second slot (:328), which nothing on the update path reads -- they never leave the kernel here."""
    dlogs, user_logs = _logs.as_device_logs(logs, agent._critic_arena.device)
    i, M = ensemble_idx, ensemble_n
    o, a, r, o1, d = replay_dict["primary_batch"]
    aa, N = agent._actor_arena, agent.num_critics
    A = agent.act_space_size
    assert 0 < M <= N
    L, stream = _lib.lib(), _lib.stream_ptr()
    with torch.no_grad():
        s1_rep = target_agent.encoder(o1)
    X1, ld1 = _rows(s1_rep)
    B, dev = X1.shape[0], X1.device
    logits = _forward(aa, i, 1, X1, ld1, B)
    subset = torch.empty(M, dtype=torch.int32, device=dev)
    _rng.source().subsets(subset, N, M)                                     # agent.py:29
    qt = _forward(target_agent._critic_arena, i * N, M, X1, ld1, B, net_index=subset)   # [M,B,A]
    v = torch.empty(B, dtype=torch.float32, device=dev)
    ev, eslot = dlogs.slots(1)                                              # zero-initialised with the buffer
    L.discrete_value(logits.data_ptr(), qt.data_ptr(), M, B, A, log_alphas[i].data_ptr(), v.data_ptr(), ev.data_ptr(), stream)
    popart = agent.popart[i]
    y = torch.empty((B, 1), dtype=torch.float32, device=dev)
    lv, slot = dlogs.slots(3)
    # PopArt de-normalisation, r + gamma (1-d) v, statistics update and re-normalisation: the continuous kernel with the
    # state value in the place of min Q - alpha logp (M = 1, no entropy term: it is inside v already)
    L.td_target(v.data_ptr(), 1, B, None, log_alphas[i].data_ptr(), r.data_ptr(), d.data_ptr(), float(gamma),
                popart.state_ptr() if popart else None, popart.ctl_ptr() if popart else None, int(bool(pop)),
                float(popart.beta) if popart else 0.0, int(popart.min_steps) if popart else 0, y.data_ptr(), lv.data_ptr(),
                stream)
    dlogs.defer(f"td_targets/mean_td_target_{i}", slot)
    dlogs.defer(f"td_targets/std_td_target_{i}", slot + 1)
    dlogs.defer(f"td_targets/entropy_bonus_{i}", eslot)
    if user_logs is not None:
        user_logs.update(dlogs.finalize())
    return y, (s1_rep, None)


def compute_backup_weights(logs, replay_dict, agent, target_agent, weight_type, weight_temp, batch_size):
    """learning_utils.py:357-398 with discrete=True: sunrise weights from the std over members of min_N Q_target(s, a_b)."""
    E, N = agent.ensemble_size, agent.num_critics
    if weight_type is None or weight_temp is None or E == 1:
        return 1.0
    if weight_type != "sunrise":
        raise NotImplementedError("discrete agents: only the sunrise Bellman weights are implemented "
                                  "(softmax weights draw Categorical samples)")
    dlogs, user_logs = _logs.as_device_logs(logs, agent._critic_arena.device)
    o, a, *_ = replay_dict["primary_batch"]
    A = agent.act_space_size
    with torch.no_grad():
        s_rep = target_agent.encoder(o)
    X, ld = _rows(s_rep)
    B, dev = X.shape[0], X.device
    L, stream = _lib.lib(), _lib.stream_ptr()
    q_all = _forward(target_agent._critic_arena, 0, E * N, X, ld, B)        # every target net's Q row: one launch
    q_sel = torch.empty((E * N, B), dtype=torch.float32, device=dev)
    act = _actions(a, B)
    L.discrete_gather_q(q_all.data_ptr(), act.data_ptr(), E * N, B, A, q_sel.data_ptr(), stream)
    w = torch.empty((B, 1), dtype=torch.float32, device=dev)
    lv, slot = dlogs.slots(4)
    L.backup_weights(q_sel.data_ptr(), E, N, B, float(weight_temp), 0, w.data_ptr(), lv.data_ptr(), stream)
    for j, name in enumerate(("mean", "max", "min", "std")):
        dlogs.defer(f"bellman_weights/{name}", slot + j)
    if user_logs is not None:
        user_logs.update(dlogs.finalize())
    return w


def critic_update(buffer, agent, target_agent, critic_optimizer, encoder_optimizer, log_alphas, batch_size, gamma,
                  critic_clip, encoder_clip, target_critic_ensemble_n, weighted_bellman_temp, weight_type, pop, augmenter,
                  encoder_lambda, aug_mix, per, update_priorities, dr3_coeff):
    """learning.py:18-141 with discrete=True."""
    _check(agent)
    if encoder_lambda:
        raise NotImplementedError("encoder invariance regulariser (lambda = 0 in every shipped config) is out of scope")
    lu.pipeline_barrier()
    ca = agent._critic_arena
    dev = ca.device
    L, stream = _lib.lib(), _lib.stream_ptr()
    E, N, B, A = agent.ensemble_size, agent.num_critics, batch_size, agent.act_space_size
    logs = _logs.DeviceLogs(dev)
    loss_all, loss_slot = logs.slots(2 * E)   # per member: {loss contribution, mean td error of its last net}
    opt = _arena.FlatAdam.attach(critic_optimizer, ca)
    enc_outs, replay_dicts = [], []
    for i in range(E):
        loss_v = loss_all[2 * i:2 * i + 2]
        rd = lu.sample_move_and_augment(buffer=buffer, batch_size=B, augmenter=augmenter, aug_mix=aug_mix, per=per)
        td_target, (s1_rep, _) = lu.compute_td_targets(
            logs=logs, replay_dict=rd, agent=agent, target_agent=target_agent, ensemble_idx=i,
            ensemble_n=target_critic_ensemble_n, log_alphas=log_alphas, pop=pop, gamma=gamma, random_process=None,
            noise_clip=None, discrete=True)
        w = lu.compute_backup_weights(logs=logs, replay_dict=rd, agent=agent, target_agent=target_agent,
                                      weight_type=weight_type, weight_temp=weighted_bellman_temp, batch_size=B, discrete=True)
        o, a, *_ = rd["primary_batch"]
        s_rep = agent.encoder(o)
        need_ds = torch.is_tensor(s_rep) and s_rep.requires_grad
        X, ld = _rows(s_rep)
        S = X.shape[1]
        act = _actions(a, B)
        q, h1, h2 = _forward(ca, i * N, N, X, ld, B, keep=True)               # [N,B,A]: every net's whole Q row
        popart = agent.popart[i]
        imp = rd["imp_weights"].float().contiguous() if per else None          # per=False: ones(1), i.e. no weighting
        dy = torch.empty((N, B, A), dtype=torch.float32, device=dev)
        L.discrete_critic_loss_seed(q.data_ptr(), N, B, A, act.data_ptr(), td_target.data_ptr(),
                                    w.data_ptr() if torch.is_tensor(w) else None, None if imp is None else imp.data_ptr(),
                                    popart.state_ptr() if popart else None, int(bool(pop)), E, 0, dy.data_ptr(),
                                    loss_v.data_ptr(), stream)
        extra, extra_scale, f1 = None, 0.0, None
        if dr3_coeff > 0:
            # DR3 (learning.py:100-108): second forward on s1; both feature sets carry gradient
            X1, ld1 = _rows(s1_rep)
            _, h1b, h2b = _forward(ca, i * N, N, X1, ld1, B, keep=True)
            dv, dslot = logs.slots(1)
            L.dr3_dot(h2.data_ptr(), h2b.data_ptr(), N, B, ca.H, dv.data_ptr(), stream)
            logs.defer(f"dr3_dotproduct_{i}", dslot)
            loss_v[0:1].add_(dv, alpha=dr3_coeff / (E * N))
            extra, extra_scale, f1 = h2b, dr3_coeff / (E * N) / (N * B), (X1, ld1, h1b, h2b)
        dxg = torch.empty((N, B, S), dtype=torch.float32, device=dev) if need_ds else None
        _ops.mlp_backward(ca, i * N, N, X, B, h1, h2, dy, ldx=ld, dh2_extra=extra, extra_scale=extra_scale, want_dw=True,
                          accumulate=False, dx=dxg, lddx=S)
        if f1 is not None:
            X1, ld1, h1b, h2b = f1
            _ops.mlp_backward(ca, i * N, N, X1, B, h1b, h2b, None, ldx=ld1, dh2_extra=h2, extra_scale=extra_scale,
                              want_dw=True, accumulate=True)
        if need_ds:
            enc_outs.append((s_rep, dxg.sum(0)))
        replay_dicts.append(rd)

    encoder_optimizer.zero_grad()
    if enc_outs:
        torch.autograd.backward([s for s, _ in enc_outs], [g for _, g in enc_outs])
    if critic_clip:
        opt.grad_norm_sq(stream)
    enc_net = None
    if enc_outs:
        enc_net = _encoder_opt.fused_step(agent.encoder, encoder_optimizer, encoder_clip)
        if enc_net is None:
            if encoder_clip:
                torch.nn.utils.clip_grad_norm_(agent.encoder.parameters(), encoder_clip)
            encoder_optimizer.step()
    member = random.choice(range(E))
    opt.step(stream, max_norm=critic_clip if critic_clip else None)   # (clipping rescales the stored gradients, as the reference's does)
    gslot = lu._member_grad_norm_slot(logs, ca, member * N, (member + 1) * N)
    logs.defer("losses/last_member_critic_td_error", loss_slot + 2 * (E - 1) + 1)
    logs.defer("losses/critic_overall_loss", [loss_slot + 2 * i for i in range(E)])
    logs.defer("gradients/critic_random_grad", gslot, transform=lambda v: v**0.5)
    if enc_net is not None:
        v, eslot = logs.slots(1)
        _encoder_opt.grad_norm_sq_into(enc_net, v)
        logs.defer("gradients/encoder_criticloss_grad_norm", eslot, transform=lambda v: v**0.5)
    elif enc_outs:
        gn = torch.linalg.vector_norm(torch.stack([p.grad.norm() for p in agent.encoder.parameters() if p.grad is not None]))
        logs.put_tensor("gradients/encoder_criticloss_grad_norm", gn)
    else:
        logs["gradients/encoder_criticloss_grad_norm"] = 0.0
    if update_priorities:
        lu.adjust_priorities(logs, rd, agent, buffer)
    return logs.finalize(), replay_dicts


def online_actor_update(buffer, agent, pop, actor_optimizer, log_alphas, batch_size, clip, augmenter, aug_mix,
                        premade_replay_dicts, per):
    """learning.py:344-421 with discrete=True: the critics are evaluated without gradient (:384-387), so the backward
    is the actor's own MLP only."""
    _check(agent)
    lu.pipeline_barrier()
    aa, ca = agent._actor_arena, agent._critic_arena
    dev = aa.device
    L, stream = _lib.lib(), _lib.stream_ptr()
    E, N, B, A = agent.ensemble_size, agent.num_critics, batch_size, agent.act_space_size
    logs = _logs.DeviceLogs(dev)
    loss_all, loss_slot = logs.slots(E)
    opt = _arena.FlatAdam.attach(actor_optimizer, aa)
    for i in range(E):
        if premade_replay_dicts is not None:
            rd = premade_replay_dicts[i]
        else:
            rd = lu.sample_move_and_augment(buffer=buffer, batch_size=B, augmenter=augmenter, aug_mix=aug_mix, per=per)
        o, *_ = rd["primary_batch"]
        with torch.no_grad():  # actor gradients do not train the encoder (learning.py:378-380)
            s_rep = agent.encoder(o)
        X, ld = _rows(s_rep)
        logits, h1, h2 = _forward(aa, i, 1, X, ld, B, keep=True)
        q = _forward(ca, i * N, N, X, ld, B)                                   # [N,B,A]
        popart = agent.popart[i]
        dlogits = torch.empty((1, B, A), dtype=torch.float32, device=dev)
        L.discrete_actor_seed(logits.data_ptr(), q.data_ptr(), N, B, A, log_alphas[i].data_ptr(),
                              popart.state_ptr() if popart else None, int(bool(pop)), E, dlogits.data_ptr(),
                              loss_all[i:i + 1].data_ptr(), stream)
        _ops.mlp_backward(aa, i, 1, X, B, h1, h2, dlogits, ldx=ld, want_dw=True, accumulate=False)
    if clip:
        opt.grad_norm_sq(stream)
    opt.step(stream, max_norm=clip if clip else None)
    member = random.choice(range(E))
    gslot = lu._member_grad_norm_slot(logs, aa, member, member + 1)
    logs.defer("gradients/random_actor_online_grad", gslot, transform=lambda v: v**0.5)
    logs.defer("losses/actor_pg_loss", [loss_slot + i for i in range(E)])
    return logs.finalize()


def alpha_update(buffer, agent, optimizers, batch_size, log_alphas, augmenter, aug_mix, target_entropy,
                 premade_replay_dicts, alpha_state_cls):
    """learning.py:222-263 with discrete=True: logp = sum_a p log p (the policy's negative entropy, :252-253)."""
    _check(agent)
    lu.pipeline_barrier()
    aa = agent._actor_arena
    dev = aa.device
    L, stream = _lib.lib(), _lib.stream_ptr()
    E, B, A = agent.ensemble_size, batch_size, agent.act_space_size
    logs = _logs.DeviceLogs(dev)
    for i in range(E):
        if premade_replay_dicts is not None:
            rd = premade_replay_dicts[i]
        else:
            rd = lu.sample_move_and_augment(buffer=buffer, batch_size=B, augmenter=augmenter, per=False, aug_mix=aug_mix)
        o, *_ = rd["primary_batch"]
        with torch.no_grad():
            s_rep = agent.encoder(o)
        X, ld = _rows(s_rep)
        logits = _forward(aa, i, 1, X, ld, B)
        plogp = torch.empty(B, dtype=torch.float32, device=dev)
        L.discrete_neg_entropy(logits.data_ptr(), B, A, plogp.data_ptr(), stream)
        pg = optimizers[i].param_groups[0]
        st = alpha_state_cls.attach(optimizers[i], log_alphas[i])
        lv, slot = logs.slots(2)
        L.alpha_step(log_alphas[i].data_ptr(), plogp.data_ptr(), B, float(target_entropy), st.state.data_ptr(),
                     st.ctl.data_ptr(), float(pg["lr"]), float(pg["betas"][0]), float(pg["betas"][1]), float(pg["eps"]),
                     lv.data_ptr(), stream)
        st.steps += 1
        st._step_tensor.fill_(st.steps)
        logs.defer(f"losses/alpha_loss_{i}", slot)
        logs.defer(f"alphas/alpha_{i}", slot + 1)
    return logs.finalize()


def _advantage(agent, replay_dict, ensemble_idx, want_priority=False):
    """Indirect advantage of a discrete agent (adv_estimator.py:45-56): A(s,a) = Q_min(s,a) - sum_a' pibar(a'|s) Q_min(s,a')
    with pibar the mean policy of ALL actors and Q_min member ``ensemble_idx``'s critics (+ its PopArt layer).
    Returns (adv [B], mask [B], priority float64 [B] or None) like learning_utils._advantage."""
    _check(agent)
    o, a, *_ = replay_dict["primary_batch"]
    i, E, N, A = ensemble_idx, agent.ensemble_size, agent.num_critics, agent.act_space_size
    L, stream = _lib.lib(), _lib.stream_ptr()
    with torch.no_grad():
        s_rep = agent.encoder(o)
    X, ld = _rows(s_rep)
    B, dev = X.shape[0], X.device
    logits = _forward(agent._actor_arena, 0, E, X, ld, B)                  # [E,B,A]: every actor, one launch
    q = _forward(agent._critic_arena, i * N, N, X, ld, B)                  # [N,B,A]
    act = _actions(a, B)
    popart = agent.popart[i]
    adv = torch.empty(B, dtype=torch.float32, device=dev)
    mask = torch.empty(B, dtype=torch.float32, device=dev)
    prio = torch.empty(B, dtype=torch.float64, device=dev) if want_priority else None
    L.discrete_advantage(logits.data_ptr(), E, q.data_ptr(), N, B, A, act.data_ptr(), popart.state_ptr() if popart else None,
                         adv.data_ptr(), mask.data_ptr(), None if prio is None else prio.data_ptr(), stream)
    return adv, mask, prio


def offline_actor_update(buffer, agent, actor_optimizer, encoder_optimizer, batch_size, actor_clip, update_encoder,
                         encoder_clip, augmenter, actor_lambda, aug_mix, premade_replay_dicts, per, filter_):
    """learning.py:144-219 with discrete=True (learning_utils.py:241-269): advantage-filtered log-likelihood of the data
    actions under the categorical policy."""
    _check(agent)
    if actor_lambda:
        raise NotImplementedError("action invariance regulariser (lambda = 0 in every shipped config) is out of scope")
    lu.pipeline_barrier()
    aa = agent._actor_arena
    dev = aa.device
    L, stream = _lib.lib(), _lib.stream_ptr()
    E, B, A = agent.ensemble_size, batch_size, agent.act_space_size
    logs = _logs.DeviceLogs(dev)
    loss_all, loss_slot = logs.slots(E)
    opt = _arena.FlatAdam.attach(actor_optimizer, aa)
    enc_outs = []
    for i in range(E):
        if premade_replay_dicts is not None:
            rd = premade_replay_dicts[i]
        else:
            rd = lu.sample_move_and_augment(buffer=buffer, batch_size=B, augmenter=augmenter, aug_mix=aug_mix, per=per)
        o, a, *_ = rd["primary_batch"]
        mask = None
        if filter_:
            _, mask, _ = _advantage(agent, rd, i)
            logs.put_tensor("losses/adv_weights_mean", mask.mean())
        if update_encoder:
            s_rep = agent.encoder(o)
        else:
            with torch.no_grad():
                s_rep = agent.encoder(o)
        need_ds = torch.is_tensor(s_rep) and s_rep.requires_grad
        X, ld = _rows(s_rep)
        S = X.shape[1]
        logits, h1, h2 = _forward(aa, i, 1, X, ld, B, keep=True)
        act = _actions(a, B)
        dlogits = torch.empty((1, B, A), dtype=torch.float32, device=dev)
        L.discrete_bc_seed(logits.data_ptr(), act.data_ptr(), None if mask is None else mask.data_ptr(), B, A, E,
                           dlogits.data_ptr(), loss_all[i:i + 1].data_ptr(), stream)
        logs.defer(f"losses/filterd_bc_loss_{i}", loss_slot + i)
        dxg = torch.empty((1, B, S), dtype=torch.float32, device=dev) if need_ds else None
        _ops.mlp_backward(aa, i, 1, X, B, h1, h2, dlogits, ldx=ld, want_dw=True, accumulate=False, dx=dxg, lddx=S)
        if need_ds:
            enc_outs.append((s_rep, dxg[0]))
    encoder_optimizer.zero_grad()
    if enc_outs:
        torch.autograd.backward([s for s, _ in enc_outs], [g for _, g in enc_outs])
    if actor_clip:
        opt.grad_norm_sq(stream)
    enc_net = None
    if update_encoder and enc_outs:
        enc_net = _encoder_opt.fused_step(agent.encoder, encoder_optimizer, encoder_clip)
    if enc_net is None and encoder_clip and enc_outs:
        torch.nn.utils.clip_grad_norm_(agent.encoder.parameters(), encoder_clip)
    opt.step(stream, max_norm=actor_clip if actor_clip else None)
    if enc_net is None and update_encoder and enc_outs:
        encoder_optimizer.step()
    logs.defer("losses/filtered_bc_overall_loss", [loss_slot + i for i in range(E)], transform=lambda v: v / E)
    member = random.choice(range(E))
    gslot = lu._member_grad_norm_slot(logs, aa, member, member + 1)
    logs.defer("gradients/actor_offline_grad_norm", gslot, transform=lambda v: v**0.5)
    if enc_outs:
        gn = torch.linalg.vector_norm(torch.stack([p.grad.norm() for p in agent.encoder.parameters() if p.grad is not None]))
        logs.put_tensor("gradients/encoder_offline_actorloss_grad_norm", gn)
    else:
        logs["gradients/encoder_offline_actorloss_grad_norm"] = 0.0
    if per:
        lu.adjust_priorities(logs, rd, agent, buffer)
    return logs.finalize()
