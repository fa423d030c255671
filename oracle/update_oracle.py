"""CPU restatement of the reference's update arithmetic.  TEST INFRASTRUCTURE ONLY (see oracle/__init__).

torch-CPU fp32, explicit forward/backward (no autograd on the restated path; a user encoder, which
is a plugin outside the path, is still differentiated by autograd).  Net loops are kept per-net like
the reference (agent.py:34) so that timing this file is a fair stand-in for the reference's CPU cost.
All ``file:line`` citations are relative to the reference root.
"""
import math

import torch
import torch.nn.functional as F

LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))  # torch.distributions.Normal.log_prob constant
LOG2 = math.log(2.0)

PARAM_NAMES = ("W1", "b1", "W2", "b2", "W3", "b3")


# --------------------------------------------------------------------------------------------
# parameter containers
# --------------------------------------------------------------------------------------------
class MLPStack:
    """G stacked 3-Linear ReLU MLPs  D -> H -> H -> O   (nets/mlps.py:113-129 critic, :11-41 / :78-93 actors).

    W1 [G,H,D]  b1 [G,H]  W2 [G,H,H]  b2 [G,H]  W3 [G,O,H]  b3 [G,O]   (nn.Linear weight layout [out,in])
    """

    def __init__(self, G, D, H, O):
        self.G, self.D, self.H, self.O = G, D, H, O
        self.W1 = torch.zeros(G, H, D)
        self.b1 = torch.zeros(G, H)
        self.W2 = torch.zeros(G, H, H)
        self.b2 = torch.zeros(G, H)
        self.W3 = torch.zeros(G, O, H)
        self.b3 = torch.zeros(G, O)

    @classmethod
    def from_modules(cls, modules):
        """modules: list of nn.Module with fc1, fc2 and out|fc3 Linear layers (reference net classes)."""
        last = lambda m: m.out if hasattr(m, "out") else m.fc3
        m0 = modules[0]
        st = cls(len(modules), m0.fc1.in_features, m0.fc1.out_features, last(m0).out_features)
        for g, m in enumerate(modules):
            st.W1[g], st.b1[g] = m.fc1.weight.detach(), m.fc1.bias.detach()
            st.W2[g], st.b2[g] = m.fc2.weight.detach(), m.fc2.bias.detach()
            st.W3[g], st.b3[g] = last(m).weight.detach(), last(m).bias.detach()
        return st

    @classmethod
    def from_arrays(cls, arrs, prefix=""):
        W1 = torch.as_tensor(arrs[prefix + "W1"]).float()
        st = cls(W1.shape[0], W1.shape[2], W1.shape[1], torch.as_tensor(arrs[prefix + "W3"]).shape[1])
        for n in PARAM_NAMES:
            setattr(st, n, torch.as_tensor(arrs[prefix + n]).float().clone())
        return st

    def random_init(self, gen, scale=None):
        """He-style init good enough for parity tests (the reference's orthogonal init is not on the path)."""
        for n in ("W1", "W2", "W3"):
            w = getattr(self, n)
            s = scale if scale is not None else (1.0 / math.sqrt(w.shape[-1]))
            w.copy_(torch.randn(w.shape, generator=gen) * s)
        for n in ("b1", "b2", "b3"):
            b = getattr(self, n)
            b.copy_(torch.randn(b.shape, generator=gen) * 0.05)
        return self

    def tensors(self):
        return [getattr(self, n) for n in PARAM_NAMES]

    def named(self, prefix=""):
        return {prefix + n: getattr(self, n) for n in PARAM_NAMES}

    def clone(self):
        c = MLPStack(self.G, self.D, self.H, self.O)
        for n in PARAM_NAMES:
            setattr(c, n, getattr(self, n).clone())
        return c

    def zeros_like(self):
        return MLPStack(self.G, self.D, self.H, self.O)


def mlp_forward(p, g, x):
    """One net: relu(fc1) -> relu(fc2) -> out.  nets/mlps.py:123-129 (critic on cat(s,a)), :32-35 (actor)."""
    h1 = F.relu(F.linear(x, p.W1[g], p.b1[g]))
    h2 = F.relu(F.linear(h1, p.W2[g], p.b2[g]))
    y = F.linear(h2, p.W3[g], p.b3[g])
    return y, h1, h2


def mlp_backward(p, g, x, h1, h2, dy, grads, dh2_extra=None, need_dx=False, need_dw=True):
    """Explicit backward of mlp_forward (what autograd does for learning.py:121 / :411).

    Accumulates into ``grads`` (an MLPStack of zeros) like autograd accumulates into .grad.
    dh2_extra: extra gradient on the post-ReLU features (DR3 term, learning.py:100-108).
    """
    dh2 = dy @ p.W3[g]
    if dh2_extra is not None:
        dh2 = dh2 + dh2_extra
    dz2 = dh2 * (h2 > 0).float()
    dh1 = dz2 @ p.W2[g]
    dz1 = dh1 * (h1 > 0).float()
    if need_dw:
        grads.W3[g] += dy.t() @ h2
        grads.b3[g] += dy.sum(0)
        grads.W2[g] += dz2.t() @ h1
        grads.b2[g] += dz2.sum(0)
        grads.W1[g] += dz1.t() @ x
        grads.b1[g] += dz1.sum(0)
    return (dz1 @ p.W1[g]) if need_dx else None


# --------------------------------------------------------------------------------------------
# policy heads  (nets/distributions.py)
# --------------------------------------------------------------------------------------------
def squash_log_std(raw, lo, hi):
    """nets/distributions.py:11-12."""
    t = torch.tanh(raw)
    return lo + 0.5 * (hi - lo) * (t + 1), t


def tanh_normal_sample(out, eps, lo, hi):
    """create_tanh_normal + (r)sample + log_prob of the cached sample.

    nets/distributions.py:9-15 (mu/log_std split, squashing), :64-104 (TanhTransform / SquashedNormal),
    torch Normal.log_prob, TransformedDistribution.log_prob with the cache hit (no clamp).
    Returns a [B,A], logp [B,1] and a cache for the backward.
    """
    A = out.shape[-1] // 2
    mu, raw = out[..., :A], out[..., A:]
    log_std, t_raw = squash_log_std(raw, lo, hi)
    std = log_std.exp()
    x = mu + eps * std
    a = torch.tanh(x)
    ladj = 2.0 * (LOG2 - x - F.softplus(-2.0 * x))
    nlp = -((x - mu) ** 2) / (2 * std**2) - std.log() - LOG_SQRT_2PI
    logp = ((0.0 - ladj) + nlp).sum(-1, keepdim=True)
    return a, logp, dict(eps=eps, std=std, a=a, t_raw=t_raw, lo=lo, hi=hi)


def tanh_normal_sample_backward(cache, da, dlogp):
    """Gradient of (a, logp) wrt the actor output [mu | raw_log_std]; rsample path (learning.py:392-399).

    x = mu + eps*std, a = tanh x, logp = sum_j [N(x_j) - ladj(x_j)].
    d logp/d x (via ladj) = 2 tanh x; the Normal terms cancel through x/mu and leave -1/std on std.
    """
    eps, std, a, t_raw = cache["eps"], cache["std"], cache["a"], cache["t_raw"]
    dx = 2.0 * a * dlogp
    if da is not None:
        dx = dx + da * (1.0 - a * a)
    dmu = dx
    dlog_std = dx * eps * std - dlogp
    draw = dlog_std * (0.5 * (cache["hi"] - cache["lo"])) * (1.0 - t_raw * t_raw)
    return torch.cat([dmu, draw], dim=-1)


def tanh_normal_logprob_data(out, a_data, lo, hi):
    """log_prob of a *dataset* action: cache miss -> atanh(clamp(a, +-0.99)).  nets/distributions.py:76-87."""
    A = out.shape[-1] // 2
    mu, raw = out[..., :A], out[..., A:]
    log_std, t_raw = squash_log_std(raw, lo, hi)
    std = log_std.exp()
    y = a_data.clamp(-0.99, 0.99)
    x = 0.5 * (y.log1p() - (-y).log1p())
    ladj = 2.0 * (LOG2 - x - F.softplus(-2.0 * x))
    nlp = -((x - mu) ** 2) / (2 * std**2) - std.log() - LOG_SQRT_2PI
    logp = ((0.0 - ladj) + nlp).sum(-1, keepdim=True)
    return logp, dict(x=x, mu=mu, std=std, t_raw=t_raw, lo=lo, hi=hi)


def tanh_normal_logprob_data_backward(cache, dlogp):
    """x is data (constant): d/dmu = (x-mu)/var, d/dstd = (x-mu)^2/std^3 - 1/std."""
    x, mu, std, t_raw = cache["x"], cache["mu"], cache["std"], cache["t_raw"]
    dmu = dlogp * (x - mu) / (std * std)
    dstd = dlogp * ((x - mu) ** 2 / (std**3) - 1.0 / std)
    draw = dstd * std * (0.5 * (cache["hi"] - cache["lo"])) * (1.0 - t_raw * t_raw)
    return torch.cat([dmu, draw], dim=-1)


def gaussian_noise_clamp(a, noise_std_normal, sigma, clip, low, high, eps=1e-6):
    """GaussianExplorationNoise.sample torch branch, learning_utils.py:48-59.  Forward value is the clamped
    action; the straight-through trick makes d out / d a = 1."""
    noise = sigma * noise_std_normal
    if clip is not None:
        noise = noise.clamp(-clip, clip)
    noisy = a + noise
    return noisy.clamp(low + eps, high - eps)


# --------------------------------------------------------------------------------------------
# PopArt  (popart.py:8-59)
# --------------------------------------------------------------------------------------------
class PopArt:
    def __init__(self, beta=1e-4, min_steps=1000, init_nu=0):
        self.mu = torch.zeros(1)
        self.nu = torch.ones(1) * init_nu
        self.beta = beta
        self.w = torch.ones(1)
        self.b = torch.zeros(1)
        self.t = 1
        self.stable = False
        self.min_steps = min_steps

    @property
    def sigma(self):  # popart.py:21-23
        return (torch.sqrt(self.nu - self.mu**2) + 1e-5).clamp(1e-4, 1e6)

    def normalize_values(self, val):  # popart.py:25-26
        return (val - self.mu) / self.sigma

    def update_stats(self, val):  # popart.py:35-52
        self.t += 1
        old_sigma = self.sigma
        old_mu = self.mu
        beta_t = self.beta / (1.0 - (1.0 - self.beta) ** self.t)
        self.mu = (1.0 - beta_t) * self.mu + beta_t * val.mean()
        self.nu = (1.0 - beta_t) * self.nu + (beta_t * (val**2).mean())
        self.stable = bool((self.t > self.min_steps) and (((1.0 - old_sigma) / self.sigma) <= 0.1))
        if self.stable:
            self.w = self.w * (old_sigma / self.sigma)
            self.b = (old_sigma * self.b + old_mu - self.mu) / (self.sigma)

    def forward(self, x, normalized=True):  # popart.py:54-59
        out = (self.w * x) + self.b
        return out if normalized else (self.sigma * out) + self.mu

    def clone(self):
        c = PopArt(self.beta, self.min_steps)
        c.mu, c.nu, c.w, c.b = self.mu.clone(), self.nu.clone(), self.w.clone(), self.b.clone()
        c.t, c.stable = self.t, self.stable
        return c


# --------------------------------------------------------------------------------------------
# optimiser / target update
# --------------------------------------------------------------------------------------------
class Adam:
    """torch.optim.Adam, single-tensor non-amsgrad path (torch/optim/adam.py _single_tensor_adam) as the
    reference configures it (main.py:188-239): coupled L2 weight decay, eps 1e-8."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = list(params)
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.t = 0

    def step(self, grads):
        self.t += 1
        b1, b2 = self.betas
        bc1 = 1 - b1**self.t
        bc2 = 1 - b2**self.t
        step_size = self.lr / bc1
        bc2_sqrt = math.sqrt(bc2)
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            if g is None:
                continue
            if self.wd != 0:
                g = g + self.wd * p
            m.lerp_(g, 1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (v.sqrt() / bc2_sqrt).add_(self.eps)
            p.addcdiv_(m, denom, value=-step_size)


def clip_grad_norm(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ (L2, error_if_nonfinite=False): learning.py:122-128."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads]))
    coef = (max_norm / (total + 1e-6)).clamp(max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


def soft_update(target_tensors, source_tensors, tau):
    """learning_utils.py:160-162: t <- t*(1-tau) + p*tau, three separately rounded fp32 ops."""
    for t, s in zip(target_tensors, source_tensors):
        t.copy_(t * (1.0 - tau) + s * tau)


def hard_update(target_tensors, source_tensors):
    """learning_utils.py:165-167."""
    for t, s in zip(target_tensors, source_tensors):
        t.copy_(s)


def grad_norm(tensors):
    """learning_utils.py:95-106 get_grad_norm."""
    tot = 0.0
    for g in tensors:
        if g is None:
            continue
        tot += g.norm(2).item() ** 2
    return tot**0.5


# --------------------------------------------------------------------------------------------
# agent container + the update functions
# --------------------------------------------------------------------------------------------
class OracleAgent:
    """Mirror of agent.Agent's learnable state (agent.py:43-130), continuous actions only.

    actors : MLPStack  G=E, D=S, O=2A (stochastic) or A (deterministic)
    critics: MLPStack  G=E*N (member-major: net g = i*N + k), D=S+A, O=1
    popart : list of PopArt or None per member
    encoder: torch nn.Module taking the obs dict (user plugin, differentiated by autograd) or None for
             the identity-on-obs["obs"] encoder of experiments/gym/train_gym.py:18-28.
    """

    def __init__(self, E, N, S, A, H, deterministic=False, log_std_low=-10.0, log_std_high=2.0,
                 popart=False, encoder=None):
        self.E, self.N, self.S, self.A, self.H = E, N, S, A, H
        self.deterministic = deterministic
        self.lo, self.hi = log_std_low, log_std_high
        self.actors = MLPStack(E, S, H, A if deterministic else 2 * A)
        self.critics = MLPStack(E * N, S + A, H, 1)
        self.popart = [PopArt() if popart else None for _ in range(E)]
        self.encoder = encoder

    def encode(self, obs):
        return obs["obs"] if self.encoder is None else self.encoder(obs)

    def clone(self):
        import copy

        c = OracleAgent(self.E, self.N, self.S, self.A, self.H, self.deterministic, self.lo, self.hi)
        c.actors, c.critics = self.actors.clone(), self.critics.clone()
        c.popart = [p.clone() if p is not None else None for p in self.popart]
        c.encoder = copy.deepcopy(self.encoder)
        return c

    # Critic.forward (agent.py:22-40)
    def critic_min(self, i, s_rep, a, nets=None):
        x = torch.cat((s_rep, a), dim=-1)
        nets = range(self.N) if nets is None else nets
        preds = [mlp_forward(self.critics, i * self.N + k, x)[0] for k in nets]
        return torch.stack(preds, 0).min(0).values

    def actor_out(self, i, s_rep):
        return mlp_forward(self.actors, i, s_rep)


def actor_sample(agent, i, s_rep, eps):
    """actor(s).sample() and its log-prob (learning_utils.py:321-338).  Deterministic actor: tanh(out),
    'sample' = loc (nets/distributions.py:107-114, nets/mlps.py:88-93)."""
    out, h1, h2 = agent.actor_out(i, s_rep)
    if agent.deterministic:
        # Normal(loc, 1e-4).log_prob(loc), summed over A: a constant
        logp = torch.full((s_rep.shape[0], 1), agent.A * (0.0 - math.log(1e-4) - LOG_SQRT_2PI))
        return torch.tanh(out), logp, dict(out=out, h1=h1, h2=h2)
    a, logp, cache = tanh_normal_sample(out, eps, agent.lo, agent.hi)
    cache.update(h1=h1, h2=h2, out=out)
    return a, logp, cache


def compute_td_target(agent, target, i, batch, rand, hp, log_alpha, logs):
    """learning_utils.py:298-354 (continuous branch)."""
    o, a, r, o1, d = batch
    popart = agent.popart[i]
    with torch.no_grad():
        s1_rep = target.encode(o1)
        a_s1, logp, _ = actor_sample(agent, i, s1_rep, rand.get("eps"))
        if hp.get("noise_sigma") is not None:
            a_s1 = gaussian_noise_clamp(a_s1, rand["noise"], hp["noise_sigma"], hp.get("noise_clip"), -1.0, 1.0)
            entropy_bonus = torch.zeros(1)
        else:
            entropy_bonus = log_alpha.exp() * logp
        s1_q = target.critic_min(i, s1_rep, a_s1, nets=rand["subset"])
        val_s1 = s1_q - entropy_bonus
        if popart is not None and hp.get("pop", False):
            val_s1 = popart.forward(val_s1, normalized=False)
        td_target = r + hp["gamma"] * (1.0 - d) * val_s1
        if popart is not None:
            popart.update_stats(td_target)
            td_target = popart.normalize_values(td_target)
    logs[f"td_targets/mean_td_target_{i}"] = td_target.mean().item()
    logs[f"td_targets/std_td_target_{i}"] = td_target.std().item()
    logs[f"td_targets/entropy_bonus_{i}"] = entropy_bonus.mean().item()
    return td_target, (s1_rep, a_s1)


def compute_backup_weights(agent, target, batch, rand, hp, logs):
    """learning_utils.py:357-398."""
    wt, temp = hp.get("weight_type"), hp.get("weight_temp")
    if wt is None or temp is None or agent.E == 1:
        return 1.0
    o, a, _, o1, _ = batch
    B = a.shape[0]
    with torch.no_grad():
        if wt == "sunrise":
            s_rep = target.encode(o)
            q_std = torch.stack([target.critic_min(j, s_rep, a) for j in range(agent.E)], 0).std(0)
            weights = torch.sigmoid(-q_std * temp) + 0.5
        elif wt == "softmax":
            s1_rep = target.encode(o1)
            q1s = []
            for j in range(agent.E):
                a1, _, _ = actor_sample(agent, j, s1_rep, rand["weight_eps"][j] if not agent.deterministic else None)
                q1s.append(agent.critic_min(j, s1_rep, a1))
            q_std = torch.stack(q1s, 0).std(0)
            weights = B * F.softmax(-q_std * temp, dim=0)
        else:
            raise ValueError(wt)
    logs["bellman_weights/mean"] = weights.mean().item()
    logs["bellman_weights/max"] = weights.max().item()
    logs["bellman_weights/min"] = weights.min().item()
    logs["bellman_weights/std"] = weights.std().item()
    return weights


def critic_update(agent, target, batches, rands, hp, log_alphas, critic_opt, encoder_opt=None):
    """learning.py:18-141 (continuous, per=False).  batches[i] = (o, a, r, o1, d) float tensors, i.e. the
    'primary_batch' of learning_utils.py:208-214; rands[i] = dict(eps, subset, noise, weight_eps).
    Returns (logs, aux) where aux carries every intermediate the parity tests compare."""
    E, N = agent.E, agent.N
    logs, aux = {}, dict(td_target=[], weights=[], q_preds=[], s_rep_grad=[])
    grads = agent.critics.zeros_like()
    loss = 0.0
    enc_outs = []
    scale = 1.0 / (E * N)
    for i in range(E):
        batch = batches[i]
        o, a, r, o1, d = batch
        B = a.shape[0]
        td_target, (s1, a1) = compute_td_target(agent, target, i, batch, rands[i], hp, log_alphas[i], logs)
        w = compute_backup_weights(agent, target, batch, rands[i], hp, logs)
        aux["td_target"].append(td_target)
        aux["weights"].append(w)
        s_rep = agent.encode(o)
        needs_enc_grad = agent.encoder is not None and s_rep.requires_grad
        s_det = s_rep.detach()
        x = torch.cat((s_det, a), dim=-1)
        popart = agent.popart[i]
        pop_on = popart is not None and hp.get("pop", False)
        dx_sum = torch.zeros_like(x) if needs_enc_grad else None
        qs, feats, cached = [], [], []
        for k in range(N):
            g = i * N + k
            q, h1, h2 = mlp_forward(agent.critics, g, x)
            qs.append(q)
            feats.append(h2)
            cached.append((h1, h2))
        dr3 = hp.get("dr3_coeff", 0.0)
        if dr3 > 0:
            # learning.py:100-108 : second forward on (s1, a1); both feature sets carry gradient
            x1 = torch.cat((s1, a1), dim=-1)
            f1 = [mlp_forward(agent.critics, i * N + k, x1) for k in range(N)]
            co = torch.stack([(feats[k] * f1[k][2]).sum(-1) for k in range(N)], 0).mean()
            logs[f"dr3_dotproduct_{i}"] = co.item()
            loss = loss + dr3 * co
        for k in range(N):
            g = i * N + k
            q = qs[k]
            qp = popart.forward(q) if pop_on else q
            td_error = td_target - qp
            loss = loss + (w * 1.0 * td_error**2).mean()
            dq = (-2.0 * scale / B) * (w * td_error)
            if pop_on:
                dq = dq * popart.w
            h1, h2 = cached[k]
            extra = (dr3 * scale / (N * B)) * f1[k][2] if dr3 > 0 else None
            dx = mlp_backward(agent.critics, g, x, h1, h2, dq, grads, dh2_extra=extra, need_dx=needs_enc_grad)
            if dr3 > 0:
                _, h1b, h2b = f1[k]
                mlp_backward(agent.critics, g, x1, h1b, h2b, torch.zeros_like(q), grads,
                             dh2_extra=(dr3 * scale / (N * B)) * h2, need_dx=False)
            if needs_enc_grad:
                dx_sum += dx
        aux["q_preds"].append(torch.stack(qs, 0))
        aux["last_td_error"] = td_error
        if needs_enc_grad:
            enc_outs.append((s_rep, dx_sum[:, : agent.S]))
            aux["s_rep_grad"].append(dx_sum[:, : agent.S].clone())
    loss = loss / (E * N)
    if encoder_opt is not None:
        encoder_opt.zero_grad()
    if enc_outs:
        torch.autograd.backward([s for s, _ in enc_outs], [g for _, g in enc_outs])
    glist = grads.tensors()
    if hp.get("critic_clip"):
        aux["critic_grad_norm"] = clip_grad_norm(glist, hp["critic_clip"])
    if hp.get("encoder_clip") and agent.encoder is not None:
        torch.nn.utils.clip_grad_norm_(agent.encoder.parameters(), hp["encoder_clip"])
    aux["grads"] = grads
    if encoder_opt is not None:
        encoder_opt.step()
    critic_opt.step(glist)
    logs["losses/last_member_critic_td_error"] = td_error.mean().item()
    logs["losses/critic_overall_loss"] = float(loss)
    return logs, aux


def online_actor_update(agent, batches, rands, hp, log_alphas, actor_opt):
    """learning.py:344-421 (continuous, use_baseline=False).  rands[i] = dict(eps [B,A], noise [B,A])."""
    E, N = agent.E, agent.N
    logs, aux = {}, {}
    grads = agent.actors.zeros_like()
    loss = 0.0
    for i in range(E):
        o = batches[i][0]
        with torch.no_grad():
            s_rep = agent.encode(o)
        B = s_rep.shape[0]
        popart = agent.popart[i]
        pop_on = popart is not None and hp.get("pop", False)
        a, logp, cache = actor_sample(agent, i, s_rep, rands[i].get("eps"))
        a_pre = a
        if agent.deterministic and rands[i].get("eps") is not None:
            # ContinuousDeterministic is Normal(loc, 1e-4) (nets/distributions.py:107-114): rsample() (learning.py:392)
            # returns loc + 1e-4*eps, and log_prob of that sample is const - eps^2/2 per dimension
            eps = rands[i]["eps"]
            a = a + 1e-4 * eps
            logp = (-(eps**2) / 2 - math.log(1e-4) - LOG_SQRT_2PI).sum(-1, keepdim=True)
        if hp.get("noise_sigma") is not None:
            a = gaussian_noise_clamp(a, rands[i]["noise"], hp["noise_sigma"], hp.get("noise_clip"), -1.0, 1.0)
            entropy = torch.zeros(1)
            alpha = 0.0
        else:
            alpha = log_alphas[i].exp().item()
            entropy = log_alphas[i].exp() * logp
        x = torch.cat((s_rep, a), dim=-1)
        outs = [mlp_forward(agent.critics, i * N + k, x) for k in range(N)]
        qstack = torch.stack([o_[0] for o_ in outs], 0)
        vals, arg = qstack.min(0)
        if pop_on:
            vals = popart.forward(vals)
        loss = loss + (vals - entropy).mean()
        # backward:  L = -(1/E) sum_i mean(vals - entropy)
        dvals = torch.full_like(vals, -1.0 / (E * B))
        if pop_on:
            dvals = dvals * popart.w
        da = torch.zeros(B, agent.A)
        dummy = agent.critics.zeros_like()
        for k in range(N):
            dq = dvals * (arg == k).float()
            _, h1, h2 = outs[k]
            dx = mlp_backward(agent.critics, i * N + k, x, h1, h2, dq, dummy, need_dx=True, need_dw=False)
            da += dx[:, agent.S:]
        if agent.deterministic:
            dout = da * (1.0 - a_pre * a_pre)  # straight-through clamp: d a / d a_pre = 1 (learning_utils.py:57-59)
        else:
            dlogp = torch.full((B, 1), alpha / (E * B))
            dout = tanh_normal_sample_backward(cache, da, dlogp)
        mlp_backward(agent.actors, i, s_rep, cache["h1"], cache["h2"], dout, grads, need_dx=False)
        aux.setdefault("actions", []).append(a)
        aux.setdefault("vals", []).append(vals)
    loss = -loss / E
    glist = grads.tensors()
    if hp.get("actor_clip"):
        clip_grad_norm(glist, hp["actor_clip"])
    aux["grads"] = grads
    actor_opt.step(glist)
    logs["losses/actor_pg_loss"] = float(loss)
    return logs, aux


def alpha_update(agent, batches, rands, log_alphas, alpha_opts, target_entropy):
    """learning.py:222-263 (continuous).  rands[i] = dict(eps)."""
    logs = {}
    for i in range(agent.E):
        o = batches[i][0]
        with torch.no_grad():
            s_rep = agent.encode(o)
            _, logp, _ = actor_sample(agent, i, s_rep, rands[i].get("eps"))
        t = (logp + target_entropy)
        alpha_loss = -(log_alphas[i] * t).mean()
        grad = -t.mean().reshape(1)
        alpha_opts[i].step([grad])
        logs[f"losses/alpha_loss_{i}"] = alpha_loss.item()
        logs[f"alphas/alpha_{i}"] = log_alphas[i].exp().item()
    return logs


def advantage(agent, i, o, a, eps_list, method="mean"):
    """adv_estimator.py:58-79 continuous_forward, method 'mean' | 'max' / n = len(eps_list) policy samples."""
    with torch.no_grad():
        s_rep = agent.encode(o)
        popart = agent.popart[i]
        pop = (lambda q: popart.forward(q)) if popart is not None else (lambda q: q)
        qs = []
        for eps in eps_list:
            act, _, _ = actor_sample(agent, i, s_rep, eps)
            qs.append(pop(agent.critic_min(i, s_rep, act)))
        value = torch.stack(qs, 0).mean(0) if method == "mean" else torch.stack(qs, 0).max(0).values
        q = pop(agent.critic_min(i, s_rep, a))
    return q - value


def offline_actor_update(agent, batches, rands, hp, actor_opt):
    """learning.py:144-219 + learning_utils.py:241-269 (continuous, update_encoder=False path for the
    restated part).  rands[i] = dict(adv_eps=[4 x [B,A]])."""
    E = agent.E
    logs, aux = {}, {}
    grads = agent.actors.zeros_like()
    loss = 0.0
    for i in range(E):
        o, a = batches[i][0], batches[i][1]
        B = a.shape[0]
        if hp.get("filter", True):
            adv = advantage(agent, i, o, a, rands[i]["adv_eps"])
            mask = (adv >= 0.0).float()
            logs["losses/adv_weights_mean"] = mask.mean().item()
        else:
            mask = torch.ones(B, 1)
        with torch.no_grad():
            s_rep = agent.encode(o)
        out, h1, h2 = agent.actor_out(i, s_rep)
        logp, cache = tanh_normal_logprob_data(out, a, agent.lo, agent.hi)
        member_loss = -(logp * mask).mean()
        logs[f"losses/filterd_bc_loss_{i}"] = member_loss.item()
        loss = loss + member_loss
        dlogp = -(mask / B) / E
        dout = tanh_normal_logprob_data_backward(cache, dlogp)
        mlp_backward(agent.actors, i, s_rep, h1, h2, dout, grads)
        aux.setdefault("adv_mask", []).append(mask)
    loss = loss / E
    glist = grads.tensors()
    if hp.get("actor_clip"):
        clip_grad_norm(glist, hp["actor_clip"])
    aux["grads"] = grads
    actor_opt.step(glist)
    logs["losses/filtered_bc_overall_loss"] = float(loss)
    return logs, aux
