"""Uniform replay ring buffer with the interface the reference's callers use (base/core/replay_memory.py:13-100:
add / add_content_of / get_latest / sample / reset / len), stored as preallocated numpy arrays instead of a list of
namedtuples.  Host-side bookkeeping of the RL half; not on the GPU hot path."""
import random
from collections import namedtuple

import numpy as np
import torch

Transition = namedtuple('Transition', ('state', 'action', 'next_state', 'reward', 'done'))


class ReplayMemory:
    def __init__(self, capacity, device):
        self.capacity, self.device = int(capacity), device
        self._store = None          # dict of field -> [capacity, dim] float32
        self._count = 0             # transitions ever written
        self.position = 0

    def reset(self):
        self._count, self.position = 0, 0

    def __len__(self):
        return min(self._count, self.capacity)

    def add(self, state, action, next_state, reward, done):
        row = [np.asarray(v, dtype=np.float32).reshape(-1) for v in (state, action, next_state, reward, done)]
        if self._store is None:
            self._store = {f: np.zeros((self.capacity, r.shape[0]), dtype=np.float32) for f, r in zip(Transition._fields, row)}
        for f, r in zip(Transition._fields, row):
            self._store[f][self.position] = r
        self.position = (self.position + 1) % self.capacity
        self._count += 1

    def add_batch(self, states, actions, next_states, rewards, dones):
        """vectorised `add` of n transitions in order (what a traced GPU episode delivers): same ring semantics."""
        cols = [np.asarray(v, dtype=np.float32).reshape(len(rewards), -1) for v in (states, actions, next_states, rewards, dones)]
        n = cols[0].shape[0]
        if n == 0:
            return
        if self._store is None:
            self._store = {f: np.zeros((self.capacity, c.shape[1]), dtype=np.float32) for f, c in zip(Transition._fields, cols)}
        idx = (self.position + np.arange(n)) % self.capacity
        if n > self.capacity:                       # only the last `capacity` rows survive, at the slots they would land in
            idx, cols = idx[-self.capacity:], [c[-self.capacity:] for c in cols]
        for f, c in zip(Transition._fields, cols):
            self._store[f][idx] = c
        self.position = int((self.position + n) % self.capacity)
        self._count += n

    def _chronological(self):
        n = len(self)
        if self._count <= self.capacity:
            return np.arange(n)
        return (np.arange(n) + self.position) % self.capacity

    def get_latest(self, latest):
        order = self._chronological()[-int(latest):]
        return [Transition(*(self._store[f][i:i + 1].copy() for f in Transition._fields)) for i in order]

    def add_content_of(self, other):
        if other._store is None:
            return
        for i in other._chronological()[-self.capacity:]:
            self.add(*(other._store[f][i] for f in Transition._fields))

    def sample(self, batch_size):
        # same stdlib-`random` consumption as the reference's random.sample(self.memory, batch_size) (:66), so that the
        # stream position seen by the next SSNE.epoch does not depend on which buffer implementation is used
        order = self._chronological() if self._count > self.capacity else None
        pick = np.asarray(random.sample(range(len(self)), batch_size))
        return tuple(torch.from_numpy(self._store[f][pick]).to(self.device) for f in Transition._fields)
