"""Ensemble sharding across the GPUs of one node: one process per GPU, `torch.distributed` plumbing.

REDQ-style critics shard naturally: each critic's forward / backward / Adam / Polyak is independent given the TD
target.  With sharding enabled, rank r builds its ``Agent`` with only its ``n_local`` critics (``local_range``), every
rank keeps a replica of the actor, the temperature and the batch (identical Philox seed => identical indices, eps and
subset on every rank, no broadcast), and the update functions exchange exactly two small tensors per phase:

  critic update : all-gather of the target critics' Q(s1, a1)  [N, B]  -> subset-min on every rank (same TD target)
  actor update  : all-gather of Q(s, pi(s))                     [N, B]  -> arg-min routing;
                  all-reduce(sum) of dL/da                       [B, A]  -> replicated actor backward + Adam

Payloads are <= 10 KB, so the exchange is pure latency (NCCL LL over NVLink); at B = 256 a sharded REDQ update is
expected to be slower than the single-GPU one -- the numbers are reported as measured next to replica-mode throughput.
"""
import torch
import torch.distributed as dist

_state = {"on": False, "group": None, "world": 1, "rank": 0, "n_global": None}


def local_range(n_global, world, rank):
    """Contiguous block of nets owned by ``rank``: sizes differ by at most one, the first ``n % world`` ranks get the
    extra net (10 nets over 8 ranks -> 2,2,1,1,1,1,1,1)."""
    base, extra = divmod(n_global, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def enable_critic_sharding(num_critics_global, group=None):
    """Call after ``dist.init_process_group``; returns (lo, hi), the global ids of this rank's critics.  Build the Agent
    with ``num_critics = hi - lo``."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if num_critics_global < world:
        raise ValueError(f"{num_critics_global} critics cannot be sharded over {world} ranks")
    _state.update(on=True, group=group, world=world, rank=rank, n_global=int(num_critics_global))
    return local_range(num_critics_global, world, rank)


def disable():
    _state.update(on=False, group=None, world=1, rank=0, n_global=None)
    disable_member_sharding()


def is_sharded():
    return _state["on"]


def n_global():
    return _state["n_global"]


def my_range():
    return local_range(_state["n_global"], _state["world"], _state["rank"])


def all_gather_q(q_local):
    """q_local [n_local, B, ...] -> [N_global, B, ...] in global net order (uneven shards are padded on the wire)."""
    return _all_gather_rows(q_local, _state["n_global"], _state["world"], _state["group"])


def _all_gather_rows(q_local, n, world, group):
    n_max = -(-n // world)
    pad = torch.zeros((n_max,) + tuple(q_local.shape[1:]), dtype=q_local.dtype, device=q_local.device)
    pad[: q_local.shape[0]].copy_(q_local)
    out = torch.empty((world * n_max,) + tuple(q_local.shape[1:]), dtype=q_local.dtype, device=q_local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    if n == world * n_max:
        return out
    rows = []
    for r in range(world):
        lo, hi = local_range(n, world, r)
        rows.append(out[r * n_max: r * n_max + (hi - lo)])
    return torch.cat(rows, dim=0)


def all_reduce_sum_(t):
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_state["group"])
    return t


# ---- SUNRISE-style ensembles: whole members sharded over the ranks ---------------------------------
# Members are independent learners (own actor, critics, batch); rank r builds its Agent with only its block of members
# and seeds its own Philox stream.  The single coupling is the SUNRISE weight, a statistic over ALL members' target
# critics evaluated on each member's batch: one all-gather of the local batches [E_local, B, S+A] and one all-gather of
# the local target critics' values on every batch [E_local*N, E_global, B] per update.  Loss normalisation keeps the
# global ensemble size.
_mstate = {"on": False, "group": None, "world": 1, "rank": 0, "e_global": None}


def enable_member_sharding(ensemble_size_global, group=None):
    """Call after ``dist.init_process_group``; returns (lo, hi), the global ids of this rank's members.  Build the Agent
    with ``ensemble_size = hi - lo``."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if ensemble_size_global < world:
        raise ValueError(f"{ensemble_size_global} members cannot be sharded over {world} ranks")
    _mstate.update(on=True, group=group, world=world, rank=rank, e_global=int(ensemble_size_global))
    return local_range(ensemble_size_global, world, rank)


def disable_member_sharding():
    _mstate.update(on=False, group=None, world=1, rank=0, e_global=None)


def members_sharded():
    return _mstate["on"]


def members_global(e_local):
    """The ensemble size the losses are normalised with."""
    return _mstate["e_global"] if _mstate["on"] else e_local


def my_members():
    return local_range(_mstate["e_global"], _mstate["world"], _mstate["rank"])


def all_gather_members(x_local):
    """x_local [E_local * k, ...] (k rows per member, member-major) -> [E_global * k, ...] in global member order."""
    e_local = my_members()[1] - my_members()[0]
    k = x_local.shape[0] // e_local
    x = x_local.reshape((e_local, k) + tuple(x_local.shape[1:]))
    out = _all_gather_rows(x.contiguous(), _mstate["e_global"], _mstate["world"], _mstate["group"])
    return out.reshape((_mstate["e_global"] * k,) + tuple(x_local.shape[1:]))


def all_reduce_members_(t):
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_mstate["group"])
    return t
