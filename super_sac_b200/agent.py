"""Agent / Critic containers (reference agent.py:13-202), re-laid-out for one-launch ensembles.

Same constructor, attributes and methods as the reference ``Agent`` so that ``main.super_sac`` and user scripts
keep working: ``.encoder .actors .critics .popart .ensemble .adv_estimator .inverse_model .contrastive_model``,
``.to .train .eval .save .load .forward .sample_action``; ``Critic.forward(*args, subset=None, return_min=True)``,
``Critic.features``, ``Critic.nets``.  What changed is where the numbers live: all E*N critic nets share one
flat fp32 arena (and all E actors another), see ``_arena.py``; the update functions in ``learning.py`` run them as
grouped launches, and ``Critic.forward`` is an autograd function over the same kernels.
"""
import copy
import os
import random

import torch
from torch import nn

from . import _arena, _ops, _rng, adv_estimator, graphed, nets, popart


class _EnsembleMLPFn(torch.autograd.Function):
    """y[g] = MLP_g(x) for a group of nets living in an arena.  Backward returns dL/dx and *accumulates* the
    parameter gradients into the arena's grad buffer (the ``.grad`` of the nn.Parameter views)."""

    @staticmethod
    def forward(ctx, x, arena, g0, G, net_index, want_param_grad):
        B = x.shape[0]
        x = x.contiguous()
        h1 = torch.empty((G, B, arena.H), dtype=torch.float32, device=x.device)
        h2 = torch.empty_like(h1)
        y = torch.empty((G, B, arena.O), dtype=torch.float32, device=x.device)
        _ops.mlp_forward(arena, g0, G, x, B, h1, h2, y, net_index=net_index)
        ctx.save_for_backward(x, h1, h2)
        ctx.arena, ctx.g0, ctx.G, ctx.net_index, ctx.want_param_grad = arena, g0, G, net_index, want_param_grad
        ctx.mark_non_differentiable(h2)
        return y, h2

    @staticmethod
    def backward(ctx, dy, _dh2):
        x, h1, h2 = ctx.saved_tensors
        arena, B = ctx.arena, x.shape[0]
        need_dx = ctx.needs_input_grad[0]
        dxg = torch.empty((ctx.G, B, arena.D), dtype=torch.float32, device=x.device) if need_dx else None
        want_dw = ctx.want_param_grad and ctx.net_index is None
        _ops.mlp_backward(arena, ctx.g0, ctx.G, x, B, h1, h2, dy.contiguous(), want_dw=want_dw, accumulate=True,
                          dx=dxg, lddx=arena.D, net_index=ctx.net_index)
        dx = dxg.sum(0) if need_dx else None
        return dx, None, None, None, None, None


class Critic(nn.Module):
    """N critic networks evaluated together (reference agent.py:13-40)."""

    def __init__(self, critic_network_cls, critic_kwargs, num_critics):
        super().__init__()
        self.nets = nn.ModuleList([critic_network_cls(**critic_kwargs) for _ in range(num_critics)])
        self.features = None
        self.num_critics = num_critics
        if not all(_arena.supported_mlp(n) for n in self.nets):
            raise NotImplementedError(
                f"{critic_network_cls.__name__}: the fused critic path implements fc1/fc2/out 3-Linear ReLU MLPs "
                "(nets.mlps.ContinuousCritic); other critic architectures are out of scope")
        self._arena, self._g0 = None, 0
        self._own_arena()

    # arena plumbing ---------------------------------------------------------------------------
    def _own_arena(self):
        n0 = self.nets[0]
        dev = n0.fc1.weight.device
        arena = _arena.MLPArena(self.num_critics, n0.fc1.in_features, n0.fc1.out_features, n0.out.out_features, dev)
        arena.bind(list(self.nets))
        self._arena, self._g0 = arena, 0

    def _adopt(self, arena, g0):
        """Called by Agent._pack: this member's nets now live at [g0, g0+N) of a shared arena."""
        self._arena, self._g0 = arena, g0

    def _apply(self, fn, recurse=True):
        # nn.Module.to()/cuda()/float(): move the arena as a whole and re-point the parameter views
        probe = fn(torch.empty(0, dtype=torch.float32, device=self._arena.device))
        if self._g0 == 0 and self._arena.G == self.num_critics:
            self._arena.to(probe.device)
        elif probe.device != self._arena.device:
            self._own_arena_from_current(probe.device)
        return self

    def _own_arena_from_current(self, device):
        n0 = self.nets[0]
        arena = _arena.MLPArena(self.num_critics, n0.fc1.in_features, n0.fc1.out_features, n0.out.out_features, device)
        arena.bind(list(self.nets))
        self._arena, self._g0 = arena, 0

    def __deepcopy__(self, memo):
        new = Critic.__new__(Critic)
        nn.Module.__init__(new)
        new.nets = copy.deepcopy(self.nets, memo)  # Parameter.__deepcopy__ clones -> un-aliased from our arena
        new.features = None
        new.num_critics = self.num_critics
        new.training = self.training
        new._arena, new._g0 = None, 0
        new._own_arena()
        return new

    # reference call surface ---------------------------------------------------------------------
    def forward(self, *args, subset=None, return_min=True):
        x = torch.cat(args, dim=-1) if len(args) > 1 else args[0]
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1]).float()
        _ops.check_cuda(x2)
        net_index, G = None, self.num_critics
        if subset is not None:
            assert 0 < subset <= self.num_critics
            net_index = torch.empty(subset, dtype=torch.int32, device=x2.device)
            _rng.source().subsets(net_index, self.num_critics, subset)
            G = subset
        y, h2 = _EnsembleMLPFn.apply(x2, self._arena, self._g0, G, net_index, torch.is_grad_enabled())
        self.features = h2.reshape((G,) + tuple(lead) + (h2.shape[-1],))
        y = y.reshape((G,) + tuple(lead) + (y.shape[-1],))
        if return_min:
            return y.min(0).values
        return tuple(y.unbind(0))


class Agent:
    """Learnable state of the (ensemble) actor-critic agent; constructor as reference agent.py:45-61."""

    def __init__(self, act_space_size, encoder, actor_network_cls, critic_network_cls, discrete=False, ensemble_size=3,
                 num_critics=2, ucb_bonus=0.0, hidden_size=256, auto_rescale_targets=True, log_std_low=-10.0,
                 log_std_high=2.0, adv_method=None, beta_dist=False):
        assert hasattr(encoder, "embedding_dim")
        if beta_dist:
            raise NotImplementedError("beta-distribution policies are out of scope (unused by every shipped config)")
        actor_kwargs = dict(state_size=encoder.embedding_dim, action_size=act_space_size, hidden_size=hidden_size)
        if not discrete:   # agent.py:76-83
            actor_kwargs.update(log_std_low=log_std_low, log_std_high=log_std_high, dist_impl="pyd")
        critic_kwargs = dict(state_size=encoder.embedding_dim, action_size=act_space_size, hidden_size=hidden_size)

        self.encoder = encoder
        self.actors = [actor_network_cls(**actor_kwargs) for _ in range(ensemble_size)]
        self.critics = [Critic(critic_network_cls, critic_kwargs, num_critics) for _ in range(ensemble_size)]
        if not all(_arena.supported_mlp(a) for a in self.actors):
            raise NotImplementedError(f"{actor_network_cls.__name__}: the fused actor path implements 3-Linear ReLU MLPs")
        self.ensemble_size = ensemble_size
        self.num_critics = num_critics
        self.act_space_size = act_space_size
        self.log_std_low, self.log_std_high = log_std_low, log_std_high
        self.deterministic = getattr(self.actors[0], "dist_impl", "pyd") == "deterministic"
        self.popart = [popart.PopArtLayer() if auto_rescale_targets else False for _ in range(ensemble_size)]
        if discrete:   # agent.py:103-113
            self.adv_estimator = adv_estimator.AdvantageEstimator(
                encoder=self.encoder, actors=self.actors, critics=self.critics, popart=self.popart, discrete=True,
                discrete_method=adv_method if adv_method else "indirect")
            self.inverse_model = nets.mlps.DiscreteInverseModel(**actor_kwargs)
            c0, a0 = self.critics[0].nets[0], self.actors[0]
            if (c0.fc1.in_features != encoder.embedding_dim or c0.out.out_features != act_space_size
                    or _arena._last_linear(a0).out_features != act_space_size):
                raise NotImplementedError("discrete agents need S -> H -> H -> A actor and critic networks "
                                          "(nets.mlps.DiscreteActor / DiscreteCritic)")
        else:
            self.adv_estimator = adv_estimator.AdvantageEstimator(
                encoder=self.encoder, actors=self.actors, critics=self.critics, popart=self.popart, discrete=False,
                continuous_method=adv_method if adv_method else "mean")
            self.inverse_model = nets.mlps.ContinuousInverseModel(**actor_kwargs)
        self.adv_estimator.bind(self)
        self.contrastive_model = nets.mlps.ContrastiveModel(state_size=encoder.embedding_dim, hidden_size=hidden_size)
        self.discrete = bool(discrete)
        self.ucb_bonus = ucb_bonus
        self._critic_arena = self._actor_arena = None
        self._pack(self.critics[0]._arena.device)

    # flat layout ------------------------------------------------------------------------------
    def _pack(self, device):
        """(Re)build the two arenas from the modules' current values and bind every Parameter to them."""
        E, N = self.ensemble_size, self.num_critics
        c0 = self.critics[0].nets[0]
        ca = _arena.MLPArena(E * N, c0.fc1.in_features, c0.fc1.out_features, c0.out.out_features, device)
        ca.bind([net for c in self.critics for net in c.nets])
        for i, c in enumerate(self.critics):
            c._adopt(ca, i * N)
        a0 = self.actors[0]
        aa = _arena.MLPArena(E, a0.fc1.in_features, a0.fc1.out_features, _arena._last_linear(a0).out_features, device)
        aa.bind(list(self.actors))
        self._critic_arena, self._actor_arena = ca, aa

    @property
    def ensemble(self):
        return zip(self.actors, self.critics)

    def to(self, device):
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.encoder = self.encoder.to(device)
        for i, p in enumerate(self.popart):
            if p:
                self.popart[i] = p.to(device)
        self.inverse_model = self.inverse_model.to(device)
        self.contrastive_model = self.contrastive_model.to(device)
        self._critic_arena.to(device)
        self._actor_arena.to(device)

    def __deepcopy__(self, memo):
        new = Agent.__new__(Agent)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_critic_arena", "_actor_arena", "adv_estimator"):
                continue
            setattr(new, k, copy.deepcopy(v, memo))
        new.adv_estimator = adv_estimator.AdvantageEstimator(
            encoder=new.encoder, actors=new.actors, critics=new.critics, popart=new.popart, discrete=self.discrete,
            discrete_method=self.adv_estimator.discrete_method, continuous_method=self.adv_estimator.cont_method)
        new.adv_estimator.bind(new)
        new._pack(self._critic_arena.device)
        return new

    def _modules(self):
        yield self.encoder
        for p in self.popart:
            if p:
                yield p
        yield from self.critics
        yield from self.actors
        yield self.inverse_model
        yield self.contrastive_model

    def eval(self):
        for m in self._modules():
            m.eval()

    def train(self):
        for m in self._modules():
            m.train()

    # checkpoints: same file names / keys as reference agent.py:172-202 ---------------------------
    def _files(self):
        yield "encoder.pt", self.encoder
        for i, p in enumerate(self.popart):
            if p:
                yield f"popart{i}.pt", p
        for i, c in enumerate(self.critics):
            yield f"critic{i}.pt", c
        for i, a in enumerate(self.actors):
            yield f"actor{i}.pt", a
        yield "inverse.pt", self.inverse_model
        yield "contrastive.pt", self.contrastive_model

    def save(self, path):
        graphed.join()
        for name, module in self._files():
            torch.save(module.state_dict(), os.path.join(path, name))

    def load(self, path):
        graphed.join()
        dev = self._critic_arena.device
        for name, module in self._files():
            module.load_state_dict(torch.load(os.path.join(path, name), map_location=dev))

    # acting path (B = num_envs; outside the update hot path) -----------------------
    def _process_obs(self, obs, num_envs=1):
        dev = self._critic_arena.device
        out = {}
        for k, v in obs.items():
            t = torch.from_numpy(v)
            out[k] = (t.unsqueeze(0) if num_envs == 1 else t).float().to(dev)
        return out

    def _process_act(self, act, num_envs=1):
        act = act.squeeze(0) if num_envs == 1 else act
        if self.discrete:   # action indices (agent.py:323-327)
            return act.cpu().numpy()
        return act.clamp(-1.0, 1.0).cpu().numpy()

    # discrete acting (agent.py:204-221, :262-320): B = num_envs rows through the module views of the arenas
    def _discrete_forward(self, s_rep):
        probs = torch.stack([actor(s_rep).probs for actor in self.actors], dim=0).mean(0)
        return torch.argmax(probs, dim=-1, keepdim=True)

    def _discrete_sample(self, s_rep, num_envs):
        if self.ucb_bonus > 0:
            dists = [actor(s_rep) for actor in self.actors]
            cands = torch.stack([d.sample() for d in dists], dim=0).unsqueeze(-1)          # [E, envs, 1]
            act_dist = random.choice(dists)
            q = torch.stack([critic(s_rep) for critic in self.critics], dim=0)              # [E_c, envs, A]
            acts = []
            for e in range(cands.shape[1]):   # per environment, as agent.py:293-306
                c_e = cands[:, e, 0]                                                        # [E]
                q_e = q[:, e, :][:, c_e]                                                    # [E_c, E]
                ucb = q_e.mean(0) + self.ucb_bonus * q_e.std(0)
                acts.append(cands[torch.argmax(ucb), e])
            return torch.stack(acts, dim=0), act_dist
        act_dist = random.choice(self.actors)(s_rep)
        return act_dist.sample().unsqueeze(-1), act_dist

    def _encode(self, obs, rolling):
        return self.encoder.forward_rolling(obs) if rolling else self.encoder(obs)

    def _policy_rows(self, s_rep, i, eps):
        """Actor ``i`` on the encoded observations [envs, S] through the kernels: a = tanh(mu + eps * std) (stochastic;
        eps = 0 gives the mean action tanh(mu)) or tanh(out) (deterministic).  One launch (ssac_policy_rows for the
        2 x 256-class networks, the tensor-core forward otherwise)."""
        from . import learning_utils as lu

        S, A = self._actor_arena.D, self.act_space_size
        B = s_rep.shape[0]
        X = torch.empty((B, S + A), dtype=torch.float32, device=s_rep.device)
        X[:, :S].copy_(s_rep)
        lu._policy_sample(self, i, X, B, S, A, None, None, rsample=False, eps=eps, keep=False)
        return X[:, S:]

    def _kernel_path(self, s_rep):
        return torch.is_tensor(s_rep) and s_rep.is_cuda and s_rep.dim() == 2 and s_rep.shape[1] == self._actor_arena.D

    def forward(self, state, from_cpu=True, num_envs=1, rolling=False):
        """Greedy action: the mean over the ensemble of each actor's mean action (reference agent.py:204-226)."""
        graphed.join()
        if from_cpu:
            state = self._process_obs(state, num_envs)
        self.eval()
        with torch.no_grad():
            s_rep = self._encode(state, rolling)
            if self.discrete:
                act = self._discrete_forward(s_rep)
            elif self._kernel_path(s_rep):
                zero = None if self.deterministic else torch.zeros((s_rep.shape[0], self.act_space_size), dtype=torch.float32,
                                                                   device=s_rep.device)
                acts = [self._policy_rows(s_rep, i, zero) for i in range(self.ensemble_size)]
                act = acts[0] if len(acts) == 1 else torch.stack(acts, dim=0).mean(0)
            else:
                act = torch.stack([actor(s_rep).mean for actor in self.actors], dim=0).mean(0)
        self.train()
        return self._process_act(act, num_envs) if from_cpu else act

    def discrete_forward(self, obs, from_cpu=True, num_envs=1, rolling=False):
        """Reference agent.py:204-221 (``forward`` dispatches here for a discrete agent)."""
        assert self.discrete
        return self.forward(obs, from_cpu=from_cpu, num_envs=num_envs, rolling=rolling)

    def continuous_forward(self, obs, from_cpu=True, num_envs=1, rolling=False):
        """Reference agent.py:223-237."""
        assert not self.discrete
        return self.forward(obs, from_cpu=from_cpu, num_envs=num_envs, rolling=rolling)

    def sample_action(self, obs, from_cpu=True, num_envs=1, return_dist=False, rolling=False):
        """Exploration action (reference agent.py:228-327): a sample of a randomly chosen actor, or -- with
        ``ucb_bonus > 0`` -- SUNRISE's UCB choice among every actor's proposal, scored by every member's critics.  On the
        device both run through the grouped kernels at B = num_envs (one launch per actor proposal, ONE launch for all
        E x N critics on all E x envs candidates); ``return_dist`` needs the torch distribution object and keeps the
        nn.Module path."""
        graphed.join()
        if from_cpu:
            obs = self._process_obs(obs, num_envs)
        with torch.no_grad():
            s_rep = self._encode(obs, rolling)
            kernels = self._kernel_path(s_rep) and not return_dist
            act_dist = None
            if self.discrete:
                act, act_dist = self._discrete_sample(s_rep, num_envs)
            elif kernels:
                from . import learning_utils as lu

                E, N, A = self.ensemble_size, self.num_critics, self.act_space_size
                envs = s_rep.shape[0]
                if self.ucb_bonus > 0:
                    random.choice(range(E))   # the reference draws its (unused here) act_dist member: keep the RNG stream
                    cands = torch.stack([self._policy_rows(s_rep, i, None) for i in range(E)], dim=0)       # [E, envs, A]
                    X = torch.cat((s_rep.unsqueeze(0).expand(E, envs, s_rep.shape[1]), cands), dim=-1)      # [E, envs, S+A]
                    X = X.reshape(E * envs, -1).contiguous()
                    q = lu._critic_values(self, 0, E * N, X, E * envs)                                      # [E*N, E*envs, 1]
                    q = q.view(E, N, E, envs).min(1).values                                                 # [E_c, E, envs]
                    ucb = q.mean(0) + self.ucb_bonus * q.std(0)                                             # [E, envs]
                    best = ucb.argmax(0)
                    act = cands[best, torch.arange(envs, device=cands.device)]
                else:
                    act = self._policy_rows(s_rep, random.choice(range(E)), None)
            elif self.ucb_bonus > 0:
                # SUNRISE UCB exploration: every actor proposes, every critic scores, pick argmax(mean + c*std)
                dists = [actor(s_rep) for actor in self.actors]
                cands = torch.stack([d.sample() for d in dists], dim=0)          # [E, envs, A]
                act_dist = random.choice(dists)
                reps = s_rep.unsqueeze(0).expand(cands.shape[0], *s_rep.shape)   # [E, envs, S]
                q = torch.stack([c(reps, cands) for c in self.critics], dim=0)   # [E_c, E, envs, 1]
                ucb = q.mean(0) + self.ucb_bonus * q.std(0)                      # [E, envs, 1]
                best = ucb.squeeze(-1).argmax(0)                                 # [envs]
                act = cands[best, torch.arange(cands.shape[1], device=cands.device)]
            else:
                act_dist = random.choice(self.actors)(s_rep)
                act = act_dist.sample()
        if from_cpu:
            act = self._process_act(act, num_envs)
        return (act, act_dist) if return_dist else act
