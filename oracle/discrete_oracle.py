"""CPU restatement of the reference's SAC-Discrete update arithmetic.  TEST INFRASTRUCTURE ONLY (see oracle/__init__).

The discrete-action branches of learning.py / learning_utils.py, in the style of update_oracle.py:
torch-CPU fp32, explicit forward / backward, per-net loops.  Pinned on tests/golden/update_discrete_*.npz, which
tests/golden/make_golden.py generated from the UNMODIFIED reference (tests/test_oracle_golden.py).
All ``file:line`` citations are relative to the reference root.
"""
import torch

from .update_oracle import MLPStack, PopArt, clip_grad_norm, mlp_backward, mlp_forward


class DiscreteOracleAgent:
    """agent.Agent(discrete=True) (agent.py:43-130): actors = DiscreteActor (nets/mlps.py:132-149, S -> H -> H -> A
    logits), critics = DiscreteCritic (nets/mlps.py:170-185, S -> H -> H -> A values), identity encoder."""

    def __init__(self, E, N, S, A, H, popart=False, encoder=None):
        self.E, self.N, self.S, self.A, self.H = E, N, S, A, H
        self.actors = MLPStack(E, S, H, A)
        self.critics = MLPStack(E * N, S, H, A)
        self.popart = [PopArt() if popart else None for _ in range(E)]
        self.encoder = encoder   # a user plugin (torch module on the obs dict, differentiated by autograd) or None = identity

    def encode(self, obs):
        return obs["obs"] if self.encoder is None else self.encoder(obs)

    def clone(self):
        import copy

        c = DiscreteOracleAgent(self.E, self.N, self.S, self.A, self.H)
        c.actors, c.critics = self.actors.clone(), self.critics.clone()
        c.popart = [p.clone() if p is not None else None for p in self.popart]
        c.encoder = copy.deepcopy(self.encoder)
        return c

    def critic_min(self, i, s_rep, nets=None):
        """Critic.forward(s, subset, return_min=True): agent.py:22-38 -> [B, A]."""
        nets = range(self.N) if nets is None else nets
        return torch.stack([mlp_forward(self.critics, i * self.N + k, s_rep)[0] for k in nets], 0).min(0).values


def policy(logits):
    """Categorical(logits=...) (nets/mlps.py:147-148): probs and log_softmax as learning_utils.py:325-326 read them."""
    logp = torch.log_softmax(logits, dim=-1)
    return logp.exp(), logp


def compute_td_target(agent, target, i, batch, subset, hp, log_alpha, logs):
    """learning_utils.py:298-354, discrete branch (:322-328)."""
    o, a, r, o1, d = batch
    popart = agent.popart[i]
    with torch.no_grad():
        s1 = target.encode(o1)
    probs, logp = policy(mlp_forward(agent.actors, i, s1)[0])
    q1 = target.critic_min(i, s1, nets=subset)
    entropy_bonus = log_alpha.exp() * logp
    val_s1 = (probs * (q1 - entropy_bonus)).sum(-1, keepdim=True)
    if popart is not None and hp.get("pop", False):
        val_s1 = popart.forward(val_s1, normalized=False)
    td_target = r + hp["gamma"] * (1.0 - d) * val_s1
    if popart is not None:
        popart.update_stats(td_target)
        td_target = popart.normalize_values(td_target)
    logs[f"td_targets/mean_td_target_{i}"] = td_target.mean().item()
    logs[f"td_targets/std_td_target_{i}"] = td_target.std().item()
    logs[f"td_targets/entropy_bonus_{i}"] = entropy_bonus.mean().item()
    return td_target, s1


def compute_backup_weights(agent, target, batch, hp, logs):
    """learning_utils.py:357-398, sunrise weights on the gathered Q(s, a) (:373-376)."""
    wt, temp = hp.get("weight_type"), hp.get("weight_temp")
    if wt is None or temp is None or agent.E == 1:
        return 1.0
    if wt != "sunrise":
        raise NotImplementedError("discrete softmax weights draw Categorical samples; not restated")
    o, a, *_ = batch
    with torch.no_grad():
        s_t = target.encode(o)
    q_std = torch.stack([target.critic_min(j, s_t).gather(-1, a.long()) for j in range(agent.E)], 0).std(0)
    weights = torch.sigmoid(-q_std * temp) + 0.5
    logs["bellman_weights/mean"] = weights.mean().item()
    logs["bellman_weights/max"] = weights.max().item()
    logs["bellman_weights/min"] = weights.min().item()
    logs["bellman_weights/std"] = weights.std().item()
    return weights


def critic_update(agent, target, batches, subsets, hp, log_alphas, critic_opt, encoder_opt=None):
    """learning.py:18-141 with discrete=True, per=False: every net's Q row is gathered at the taken action (:90-92).
    A trainable encoder receives the summed input gradients of the member's critics through autograd (:121)."""
    E, N = agent.E, agent.N
    logs, aux = {}, dict(td_target=[], weights=[])
    enc_outs = []
    grads = agent.critics.zeros_like()
    loss = 0.0
    scale = 1.0 / (E * N)
    dr3 = hp.get("dr3_coeff", 0.0)
    for i in range(E):
        o, a, r, o1, d = batches[i]
        B = a.shape[0]
        td_target, s1 = compute_td_target(agent, target, i, batches[i], subsets[i], hp, log_alphas[i], logs)
        w = compute_backup_weights(agent, target, batches[i], hp, logs)
        aux["td_target"].append(td_target)
        aux["weights"].append(w)
        s_rep = agent.encode(o)
        needs_enc_grad = agent.encoder is not None and s_rep.requires_grad
        x = s_rep.detach()
        dx_sum = torch.zeros_like(x) if needs_enc_grad else None
        popart = agent.popart[i]
        pop_on = popart is not None and hp.get("pop", False)
        outs = [mlp_forward(agent.critics, i * N + k, x) for k in range(N)]
        if dr3 > 0:   # learning.py:100-108: second forward on s1, both feature sets carry gradient
            f1 = [mlp_forward(agent.critics, i * N + k, s1) for k in range(N)]
            co = torch.stack([(outs[k][2] * f1[k][2]).sum(-1) for k in range(N)], 0).mean()
            logs[f"dr3_dotproduct_{i}"] = co.item()
            loss = loss + dr3 * co
        onehot = torch.zeros(B, agent.A).scatter_(1, a.long(), 1.0)
        for k in range(N):
            q, h1, h2 = outs[k]
            q_sel = q.gather(-1, a.long())
            qp = popart.forward(q_sel) if pop_on else q_sel
            td_error = td_target - qp
            loss = loss + (w * td_error**2).mean()
            dq = (-2.0 * scale / B) * (w * td_error)
            if pop_on:
                dq = dq * popart.w
            extra = (dr3 * scale / (N * B)) * f1[k][2] if dr3 > 0 else None
            dx = mlp_backward(agent.critics, i * N + k, x, h1, h2, dq * onehot, grads, dh2_extra=extra,
                              need_dx=needs_enc_grad)
            if needs_enc_grad:
                dx_sum += dx
            if dr3 > 0:
                _, h1b, h2b = f1[k]
                mlp_backward(agent.critics, i * N + k, s1, h1b, h2b, torch.zeros(B, agent.A), grads,
                             dh2_extra=(dr3 * scale / (N * B)) * h2)
        if needs_enc_grad:
            enc_outs.append((s_rep, dx_sum))
    loss = loss / (E * N)
    if encoder_opt is not None:
        encoder_opt.zero_grad()
    if enc_outs:
        torch.autograd.backward([s_ for s_, _ in enc_outs], [g for _, g in enc_outs])
    glist = grads.tensors()
    if hp.get("critic_clip"):
        clip_grad_norm(glist, hp["critic_clip"])
    if hp.get("encoder_clip") and agent.encoder is not None:
        torch.nn.utils.clip_grad_norm_(agent.encoder.parameters(), hp["encoder_clip"])
    aux["grads"] = grads
    if encoder_opt is not None:
        encoder_opt.step()
    critic_opt.step(glist)
    logs["losses/last_member_critic_td_error"] = td_error.mean().item()
    logs["losses/critic_overall_loss"] = float(loss)
    return logs, aux


def online_actor_update(agent, batches, hp, log_alphas, actor_opt):
    """learning.py:344-421, discrete branch (:382-390): loss = -(1/E) sum_i mean_b sum_a p (Q_min - alpha log p)."""
    E, N = agent.E, agent.N
    logs, aux = {}, {}
    grads = agent.actors.zeros_like()
    loss = 0.0
    for i in range(E):
        with torch.no_grad():
            s = agent.encode(batches[i][0])
        B = s.shape[0]
        popart = agent.popart[i]
        logits, h1, h2 = mlp_forward(agent.actors, i, s)
        probs, logp = policy(logits)
        vals = agent.critic_min(i, s)
        if popart is not None and hp.get("pop", False):
            vals = popart.forward(vals)
        alpha = log_alphas[i].exp()
        g = vals - alpha * logp
        f = (probs * g).sum(-1, keepdim=True)
        loss = loss + f.mean()
        # d f / d logit_k = p_k (g_k - f): the -alpha * sum_a p_a dlogp_a/dz_k term vanishes
        dlogits = (-1.0 / (E * B)) * probs * (g - f)
        mlp_backward(agent.actors, i, s, h1, h2, dlogits, grads)
    loss = -loss / E
    glist = grads.tensors()
    if hp.get("actor_clip"):
        clip_grad_norm(glist, hp["actor_clip"])
    aux["grads"] = grads
    actor_opt.step(glist)
    logs["losses/actor_pg_loss"] = float(loss)
    return logs, aux


def alpha_update(agent, batches, log_alphas, alpha_opts, target_entropy):
    """learning.py:222-263, discrete branch (:252-253): logp = sum_a p log p (the negative entropy)."""
    logs = {}
    for i in range(agent.E):
        with torch.no_grad():
            s = agent.encode(batches[i][0])
        probs, logp = policy(mlp_forward(agent.actors, i, s)[0])
        t = (probs * logp).sum(-1) + target_entropy
        alpha_loss = -(log_alphas[i] * t).mean()
        alpha_opts[i].step([-t.mean().reshape(1)])
        logs[f"losses/alpha_loss_{i}"] = alpha_loss.item()
        logs[f"alphas/alpha_{i}"] = log_alphas[i].exp().item()
    return logs


def advantage(agent, i, s, a):
    """adv_estimator.py:45-56 (discrete 'indirect'): A(s,a) = Q_min(s,a) - sum_a' mean_e pi_e(a'|s) Q_min(s,a'), with
    member i's critics (+ PopArt whenever the member has one: adv_estimator.py:30-35 has no ``pop`` switch)."""
    probs = torch.stack([policy(mlp_forward(agent.actors, e, s)[0])[0] for e in range(agent.E)], 0).mean(0)
    min_q = agent.critic_min(i, s)
    if agent.popart[i] is not None:
        min_q = agent.popart[i].forward(min_q)
    value = (probs * min_q).sum(-1, keepdim=True)
    return min_q.gather(-1, a.long()) - value


def offline_actor_update(agent, batches, hp, actor_opt, filter_=True):
    """learning.py:144-219 with discrete=True, update_encoder=False; learning_utils.py:241-269: masked log-likelihood of
    the data action under Categorical(logits).  d(-mean(mask * log p_a))/dlogit_k = -(mask/B) (1[k=a] - p_k)."""
    E = agent.E
    logs, aux = {}, dict(adv=[])
    grads = agent.actors.zeros_like()
    loss = 0.0
    for i in range(E):
        o, a, *_ = batches[i]
        s = o["obs"]
        B = s.shape[0]
        if filter_:
            adv = advantage(agent, i, s, a)
            mask = (adv >= 0.0).float()
            aux["adv"].append(adv)
            logs["losses/adv_weights_mean"] = mask.mean().item()
        else:
            mask = torch.ones(B, 1)
        logits, h1, h2 = mlp_forward(agent.actors, i, s)
        probs, logp = policy(logits)
        member = -(logp.gather(-1, a.long()) * mask).mean()
        logs[f"losses/filterd_bc_loss_{i}"] = member.item()
        loss = loss + member
        onehot = torch.zeros(B, agent.A).scatter_(1, a.long(), 1.0)
        mlp_backward(agent.actors, i, s, h1, h2, (-1.0 / (E * B)) * mask * (onehot - probs), grads)
    loss = loss / E
    glist = grads.tensors()
    if hp.get("actor_clip"):
        clip_grad_norm(glist, hp["actor_clip"])
    aux["grads"] = grads
    actor_opt.step(glist)
    logs["losses/filtered_bc_overall_loss"] = float(loss)
    return logs, aux


def priorities(agent, member, batch):
    """learning_utils.py:288-295: relu(A) + 1e-4 with one (randomly chosen) member's advantage."""
    o, a, *_ = batch
    return torch.relu(advantage(agent, member, o["obs"], a)).squeeze(1).double() + 1e-4
