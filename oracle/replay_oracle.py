"""CPU restatement of the reference's replay buffers — TEST INFRASTRUCTURE (only tests/ may import it).

Follows src/model/replay_buffer.py (ReplayBuffer :7-73, PrioritizedReplayBuffer :76-209) and
src/model/common/segment_tree.py (:4-145) operation by operation, with the stdlib `random` draws replaced by
injected uniforms: randint(0, n - 1) -> int(u * n), random() -> u.  tests/test_replay.py pins it against the
reference classes themselves (imported from /root/reference/src, `random` monkeypatched the same way)."""
import numpy as np


class ReplayOracle(object):
    def __init__(self, size, prioritized=False, alpha=0.6, beta=0.4):
        self.storage, self.maxsize, self.next_idx = [], size, 0
        self.prioritized, self.alpha, self.beta = prioritized, alpha, beta
        cap = 1
        while cap < size:
            cap *= 2
        self.cap = cap
        self.sum = [0.0] * (2 * cap)                 # SumSegmentTree  (segment_tree.py:90-96)
        self.min = [float("inf")] * (2 * cap)        # MinSegmentTree  (:129-135)
        self.max_priority = 1.0

    def __len__(self):
        return len(self.storage)

    def _set(self, idx, val):                        # SegmentTree.__setitem__ (:76-87)
        i = idx + self.cap
        self.sum[i] = val
        self.min[i] = val
        i //= 2
        while i >= 1:
            self.sum[i] = self.sum[2 * i] + self.sum[2 * i + 1]
            self.min[i] = min(self.min[2 * i], self.min[2 * i + 1])
            i //= 2

    def add(self, obs_t, action, reward, obs_tp1, done):   # replay_buffer.py:24-31, :108-113
        idx = self.next_idx
        data = (obs_t, action, reward, obs_tp1, done)
        if self.next_idx >= len(self.storage):
            self.storage.append(data)
        else:
            self.storage[self.next_idx] = data
        self.next_idx = (self.next_idx + 1) % self.maxsize
        if self.prioritized:
            self._set(idx, self.max_priority ** self.alpha)

    def _reduce(self, start, end, node, ns, ne):     # _reduce_helper (:38-53)
        if start == ns and end == ne:
            return self.sum[node]
        mid = (ns + ne) // 2
        if end <= mid:
            return self._reduce(start, end, 2 * node, ns, mid)
        if mid + 1 <= start:
            return self._reduce(start, end, 2 * node + 1, mid + 1, ne)
        return self._reduce(start, mid, 2 * node, ns, mid) + self._reduce(mid + 1, end, 2 * node + 1, mid + 1, ne)

    def sum_range(self, start, end):                 # SumSegmentTree.sum -> reduce (:55-74): `end -= 1`
        return self._reduce(start, end - 1, 1, 0, self.cap - 1)

    def find_prefixsum_idx(self, prefixsum):         # (:107-126)
        i = 1
        while i < self.cap:
            if self.sum[2 * i] > prefixsum:
                i = 2 * i
            else:
                prefixsum -= self.sum[2 * i]
                i = 2 * i + 1
        return i - self.cap

    def encode(self, idxes):                         # _encode_sample (:33-44)
        cols = list(zip(*[self.storage[i] for i in idxes]))
        return (np.array(cols[0]), np.array(cols[1]), np.array(cols[2]), np.array(cols[3]), np.array(cols[4]))

    def sample(self, uniforms):
        n = len(self.storage)
        if not self.prioritized:                     # :46-67
            idxes = [min(int(u * n), n - 1) for u in uniforms]
            return self.encode(idxes) + (idxes,)
        idxes = [self.find_prefixsum_idx(u * self.sum_range(0, n - 1)) for u in uniforms]   # :113-120
        p_min = self.min[1] / self.sum[1]            # :158-159 (min() and sum() over the whole tree)
        max_weight = (p_min * n) ** (-self.beta)
        weights = [((self.sum[self.cap + i] / self.sum[1]) * n) ** (-self.beta) / max_weight for i in idxes]
        return self.encode(idxes) + (np.array(weights), idxes)

    def update_priorities(self, idxes, priorities):  # :173-195
        if not self.prioritized:
            return
        for i, p in zip(idxes, priorities):
            self._set(int(i), float(p) ** self.alpha)
            self.max_priority = max(self.max_priority, float(p))
