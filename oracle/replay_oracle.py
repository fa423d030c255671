"""CPU restatement of the reference replay buffer (numpy).  TEST INFRASTRUCTURE ONLY (see oracle/__init__).

Follows replay.py:10-190 (ring storage, uniform and prioritised sampling) and replay.py:207-353
(float64 sum / min segment trees).  Integer / byte work: the CUDA path must match this bit-for-bit.
"""
import numpy as np


class RingStorage:
    """replay.py:10-95  ReplayBufferStorage: pre-allocated ring, dtype per obs key from the first push."""

    def __init__(self, size, state_example, act_example):
        self.size = size
        self.action = np.zeros((size,) + act_example.shape, dtype=np.float32)
        self.reward = np.zeros((size, 1), dtype=np.float32)
        self.done = np.zeros((size, 1), dtype=np.uint8)
        self.s = {k: np.zeros((size,) + v.shape, dtype=v.dtype) for k, v in state_example.items()}
        self.s1 = {k: np.zeros((size,) + v.shape, dtype=v.dtype) for k, v in state_example.items()}
        self.next_idx = 0
        self.filled = 0

    def __len__(self):
        return self.filled

    def add(self, s, a, r, s1, d):
        """replay.py:32-61 (batched or single transition, wrap-around with modulo)."""
        a = np.asarray(a)
        if a.ndim > 1:
            n = len(a)
        else:
            n = 1
            r, d = np.array(r), np.array(d)
        R = np.arange(self.next_idx, self.next_idx + n) % self.size
        for k in s:
            self.s[k][R] = np.asarray(s[k]).astype(self.s[k].dtype)
        for k in s1:
            self.s1[k][R] = np.asarray(s1[k]).astype(self.s1[k].dtype)
        self.action[R] = a.astype(np.float32)
        self.reward[R] = np.asarray(r).astype(np.float32)
        self.done[R] = np.asarray(d).astype(np.uint8)
        self.filled = min(max(self.next_idx + n, self.filled), self.size)
        self.next_idx = (self.next_idx + n) % self.size
        return R

    def gather(self, idx):
        """replay.py:66-84: fancy-index gather; action gets a trailing dim if 1-D."""
        idx = np.asarray(idx)
        s = {k: v[idx] for k, v in self.s.items()}
        s1 = {k: v[idx] for k, v in self.s1.items()}
        a = self.action[idx]
        if a.ndim < 2:
            a = a[:, None]
        return s, a, self.reward[idx], s1, self.done[idx]


def _unique_sorted(x):
    """replay.py:193-204."""
    if len(x) == 1:
        return x
    return x[np.append(x[1:] != x[:-1], True)]


class SegmentTree:
    """replay.py:207-282 with the numpy float64 value array of :290 / :344."""

    def __init__(self, capacity, op, neutral):
        assert capacity > 0 and capacity & (capacity - 1) == 0
        self.capacity = capacity
        self.value = np.full(2 * capacity, neutral, dtype=np.float64)
        self.op = op

    def _reduce(self, start, end, node, lo, hi):
        if start == lo and end == hi:
            return self.value[node]
        mid = (lo + hi) // 2
        if end <= mid:
            return self._reduce(start, end, 2 * node, lo, mid)
        if mid + 1 <= start:
            return self._reduce(start, end, 2 * node + 1, mid + 1, hi)
        return self.op(self._reduce(start, mid, 2 * node, lo, mid), self._reduce(mid + 1, end, 2 * node + 1, mid + 1, hi))

    def reduce(self, start=0, end=None):
        """replay.py:247-261 (end is exclusive like a slice; negative wraps)."""
        if end is None:
            end = self.capacity
        if end < 0:
            end += self.capacity
        end -= 1
        return self._reduce(start, end, 1, 0, self.capacity - 1)

    def set(self, idx, val):
        """replay.py:263-277: leaf write (numpy fancy assignment: last write wins on duplicates) then a
        level-by-level recompute of the touched parents.  Note the reference de-duplicates with a
        *sorted-array* unique on possibly unsorted indices; recomputing a parent twice is idempotent."""
        idxs = np.atleast_1d(np.asarray(idx)) + self.capacity
        self.value[idxs] = val
        idxs = _unique_sorted(idxs // 2)
        while len(idxs) > 1 or idxs[0] > 0:
            self.value[idxs] = self.op(self.value[2 * idxs], self.value[2 * idxs + 1])
            idxs = _unique_sorted(idxs // 2)

    def get(self, idx):
        return self.value[self.capacity + np.asarray(idx)]

    def find_prefixsum_idx(self, prefixsum):
        """replay.py:301-336: descend from the root; go left when value[left] > prefix, else subtract."""
        prefixsum = np.array(prefixsum, dtype=np.float64, copy=True)
        idx = np.ones(len(prefixsum), dtype=np.int64)
        cont = np.ones(len(prefixsum), dtype=bool)
        while np.any(cont):
            idx[cont] = 2 * idx[cont]
            new = np.where(self.value[idx] <= prefixsum, prefixsum - self.value[idx], prefixsum)
            idx = np.where(np.logical_or(self.value[idx] > prefixsum, np.logical_not(cont)), idx, idx + 1)
            prefixsum = new
            cont = idx < self.capacity
        return idx - self.capacity


class ReplayOracle:
    """replay.py:98-190 ReplayBuffer (uniform + PER)."""

    def __init__(self, size, alpha=0.6, beta=1.0):
        self.maxsize = size
        self.alpha, self.beta = alpha, beta
        cap = 1
        while cap < size:
            cap *= 2
        self.it_sum = SegmentTree(cap, np.add, 0.0)
        self.it_min = SegmentTree(cap, np.minimum, float("inf"))
        self.max_priority = 1.0
        self.storage = None

    def __len__(self):
        return len(self.storage) if self.storage is not None else 0

    def push(self, s, a, r, s1, d, priorities=None):
        """replay.py:106-119, :156-161."""
        a = np.asarray(a)
        if self.storage is None:
            if a.ndim > 1:
                ex_a, ex_s = a[0], {k: np.asarray(v)[0] for k, v in s.items()}
            else:
                ex_a, ex_s = a, {k: np.asarray(v) for k, v in s.items()}
            self.storage = RingStorage(self.maxsize, ex_s, ex_a)
        R = self.storage.add(s, a, r, s1, d)
        if priorities is None:
            priorities = self.max_priority
        self.it_sum.set(R, priorities**self.alpha)
        self.it_min.set(R, priorities**self.alpha)
        return R

    def load_experience(self, s, a, r, s1, d):
        """replay.py:131-137."""
        r, d = np.asarray(r), np.asarray(d)
        if r.ndim < 2:
            r = r[:, None]
        if d.ndim < 2:
            d = d[:, None]
        return self.push(s, a, r, s1, d)

    def sample_uniform(self, idx):
        """replay.py:121-126, :179-181 with the torch.randint draw supplied by the caller."""
        return self.storage.gather(idx), np.asarray(idx)

    def sample(self, uniform01):
        """replay.py:163-177 with the np.random.random draw supplied by the caller."""
        n = len(self.storage)
        total = self.it_sum.reduce(0, n - 1)
        mass = np.asarray(uniform01, dtype=np.float64) * total
        idxes = self.it_sum.find_prefixsum_idx(mass)
        p_min = self.it_min.reduce() / self.it_sum.reduce()
        max_weight = (p_min * n) ** (-self.beta)
        p_sample = self.it_sum.get(idxes) / self.it_sum.reduce()
        weights = (p_sample * n) ** (-self.beta) / max_weight
        return self.storage.gather(idxes), weights, idxes

    def update_priorities(self, idxes, priorities):
        """replay.py:183-190."""
        idxes, priorities = np.asarray(idxes), np.asarray(priorities)
        assert len(idxes) == len(priorities)
        assert np.min(priorities) > 0 and np.min(idxes) >= 0 and np.max(idxes) < len(self.storage)
        self.it_sum.set(idxes, priorities**self.alpha)
        self.it_min.set(idxes, priorities**self.alpha)
        self.max_priority = max(self.max_priority, np.max(priorities))
