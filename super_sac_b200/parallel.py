"""Ensemble sharding across the GPUs of one node: one process per GPU, `torch.distributed` plumbing.

REDQ-style critics shard naturally: each critic's forward / backward / Adam / Polyak is independent given the TD
target.  With sharding enabled, rank r builds its ``Agent`` with only its ``n_local`` critics (``local_range``), every
rank keeps a replica of the actor, the temperature and the batch (identical Philox seed => identical indices, eps and
subset on every rank, no broadcast), and the update functions exchange exactly two small tensors per phase:

  critic update : all-gather of the target critics' Q(s1, a1)  [N, B]  -> subset-min on every rank (same TD target)
  actor update  : all-gather of Q(s, pi(s))                     [N, B]  -> arg-min routing;
                  all-reduce(sum) of dL/da                       [B, A]  -> replicated actor backward + Adam

Payloads are <= 10 KB, so the exchange is pure latency.  On CUDA devices with peer access the exchange does not go
through NCCL at all: every rank stores its rows straight into all peers' symmetric buffers over NVLink and raises a
signal, the consumer spins on its signals (ssac_peer_put / ssac_peer_wait, csrc/ssac_peer.cu; the buffers' peer mappings
come from torch.distributed._symmetric_memory).  ``torch.distributed`` collectives remain as the fallback (CPU / gloo
tests, boxes without peer access) -- ``exchange_name()`` says which one is live.
"""
import os

import torch
import torch.distributed as dist

_state = {"on": False, "group": None, "world": 1, "rank": 0, "n_global": None}


# ---- peer-memory exchange sites ---------------------------------------------------------------------------------------
_peer = {"sites": {}, "why_not": None, "disabled": os.environ.get("SSAC_PEER_EXCHANGE", "1") == "0"}


class _PeerSite:
    """One exchange site: a symmetric buffer of two halves (epoch parity), a symmetric signal row and a device epoch."""

    def __init__(self, half_bytes, group, device):
        import torch.distributed._symmetric_memory as symm

        from . import _lib

        self.lib = _lib
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.half = (int(half_bytes) + 15) // 16 * 16
        name = self.group.group_name
        try:
            symm.enable_symm_mem_for_group(name)
        except Exception:  # noqa: BLE001  (newer torch enables groups lazily)
            pass
        self.buf = symm.empty(2 * self.half, dtype=torch.uint8, device=device)
        self.sig = symm.empty(64, dtype=torch.int32, device=device)
        self.buf.zero_()
        self.sig.zero_()
        self._hb = symm.rendezvous(self.buf, name)
        self._hs = symm.rendezvous(self.sig, name)
        self.buf_ptrs = torch.tensor([int(p) for p in self._hb.buffer_ptrs], dtype=torch.int64, device=device)
        self.sig_ptrs = torch.tensor([int(p) for p in self._hs.buffer_ptrs], dtype=torch.int64, device=device)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)   # every rank's buffers are zeroed and mapped before the first put

    def put(self, src, dst_off):
        src = src.contiguous()
        self.lib.lib().peer_put(src.data_ptr(), src.numel() * src.element_size(), int(dst_off), self.half, self.buf_ptrs.data_ptr(),
                                self.sig_ptrs.data_ptr(), self.rank, self.world, self.epoch.data_ptr(), self.lib.stream_ptr())

    def wait(self, out, row_index=None, total_bytes=None):
        """Gather the exchanged buffer into ``out`` -- all of it, or (``row_index``: device int32) only those rows."""
        nbytes = out.numel() * out.element_size() if row_index is None else int(total_bytes)
        n_rows = 0 if row_index is None else row_index.numel()
        self.lib.lib().peer_wait(self.buf.data_ptr(), self.half, nbytes, self.sig.data_ptr(), self.world, self.epoch.data_ptr(),
                                 out.data_ptr(), None if row_index is None else row_index.data_ptr(), n_rows,
                                 0 if row_index is None else out[0].numel() * out.element_size(), self.lib.stream_ptr())


def _site(name, nbytes, group, device):
    """The peer-exchange site ``name`` (created on first use, outside any stream capture), or None when peer exchange is
    not available -- the reason is kept for ``exchange_name()``."""
    if _peer["disabled"] or _peer["why_not"] is not None or device.type != "cuda":
        return None
    key = (name, int(nbytes), id(group))
    site = _peer["sites"].get(key)
    if site is None:
        if torch.cuda.is_current_stream_capturing():
            return None
        try:
            site = _peer["sites"][key] = _PeerSite(nbytes, group, device)
        except Exception as e:  # noqa: BLE001
            _peer["why_not"] = f"{type(e).__name__}: {str(e)[:120]}"
            return None
    return site


def peer_exchange_ready():
    """True when the exchanges run over peer memory (kernels on the caller's current stream), not NCCL collectives."""
    return bool(_peer["sites"]) and _peer["why_not"] is None and not _peer["disabled"]


def exchange_name():
    if _peer["sites"] and _peer["why_not"] is None:
        return "NVLink peer stores + signals (ssac_peer_put / ssac_peer_wait over symmetric memory)"
    why = "disabled" if _peer["disabled"] else (_peer["why_not"] or "not used yet / CPU tensors")
    return f"torch.distributed collectives (peer exchange unavailable: {why})"


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


def all_gather_q(q_local, site="q", select=None):
    """q_local [n_local, B, ...] -> [N_global, B, ...] in global net order (or only the rows ``select``).  ``site`` names
    the call site: every site has its own symmetric buffer and signals."""
    return _all_gather_rows(q_local, _state["n_global"], _state["world"], _state["group"], site_name=site, select=select)


def _all_gather_rows(q_local, n, world, group, site_name="rows", select=None):
    """``select`` (device int32 [m]): return only those global rows (cheaper than gathering everything and indexing)."""
    rank = dist.get_rank(group)
    row = q_local[0].numel() * q_local.element_size() if q_local.shape[0] else 0
    site = _site(site_name, n * row, group, q_local.device) if row else None
    if site is not None:
        site.put(q_local, local_range(n, world, rank)[0] * row)
        if select is not None:
            out = torch.empty((select.numel(),) + tuple(q_local.shape[1:]), dtype=q_local.dtype, device=q_local.device)
            site.wait(out, row_index=select, total_bytes=n * row)
        else:
            out = torch.empty((n,) + tuple(q_local.shape[1:]), dtype=q_local.dtype, device=q_local.device)
            site.wait(out)
        return out
    n_max = -(-n // world)
    pad = torch.zeros((n_max,) + tuple(q_local.shape[1:]), dtype=q_local.dtype, device=q_local.device)
    pad[: q_local.shape[0]].copy_(q_local)
    out = torch.empty((world * n_max,) + tuple(q_local.shape[1:]), dtype=q_local.dtype, device=q_local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    if n == world * n_max:
        return out if select is None else out.index_select(0, select.long())
    rows = []
    for r in range(world):
        lo, hi = local_range(n, world, r)
        rows.append(out[r * n_max: r * n_max + (hi - lo)])
    out = torch.cat(rows, dim=0)
    return out if select is None else out.index_select(0, select.long())


def _all_reduce_sum_(t, world, group, site_name):
    site = _site(site_name, world * t.numel() * t.element_size(), group, t.device)
    if site is not None:
        # every rank lands its partial in slot `rank` of every peer, then sums the slots in rank order: the same
        # arithmetic on every rank (replicas stay bit-identical), no ring order to depend on
        parts = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        site.put(t, dist.get_rank(group) * t.numel() * t.element_size())
        site.wait(parts)
        t.copy_(parts.sum(0))
        return t
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def all_reduce_sum_(t, site="sum"):
    return _all_reduce_sum_(t, _state["world"], _state["group"], site)


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
    out = _all_gather_rows(x.contiguous(), _mstate["e_global"], _mstate["world"], _mstate["group"],
                           site_name="members%d" % (x[0].numel() if x.shape[0] else 0))
    return out.reshape((_mstate["e_global"] * k,) + tuple(x_local.shape[1:]))


def all_reduce_members_(t):
    return _all_reduce_sum_(t, _mstate["world"], _mstate["group"], "members_sum")
