"""Minimal mirror of base/core/replay_memory.py:13-100 (uniform ring buffer of (state, action, next_state, reward, done))."""
import random
from collections import namedtuple

import numpy as np
import torch

Transition = namedtuple('Transition', ('state', 'action', 'next_state', 'reward', 'done'))


class ReplayMemory:
    def __init__(self, capacity, device):
        self.device, self.capacity = device, capacity
        self.memory, self.position = [], 0

    def reset(self):
        self.memory, self.position = [], 0

    def add(self, *args):
        if len(self.memory) < self.capacity:
            self.memory.append(None)
        reshaped = [np.reshape(np.asarray(a, dtype=np.float32), (1, -1)) for a in args]
        self.memory[self.position] = Transition(*reshaped)
        self.position = (self.position + 1) % self.capacity

    def add_content_of(self, other):
        latest = other.get_latest(self.capacity)
        for t in latest:
            self.add(*t)

    def get_latest(self, latest):
        if self.capacity < latest:
            latest_trans = self.memory[self.position:].copy() + self.memory[:self.position].copy()
        elif len(self.memory) < self.capacity:
            latest_trans = self.memory[-latest:].copy()
        elif self.position >= latest:
            latest_trans = self.memory[:self.position][-latest:].copy()
        else:
            latest_trans = self.memory[-latest + self.position:].copy() + self.memory[:self.position].copy()
        return latest_trans

    def sample(self, batch_size):
        transitions = random.sample(self.memory, batch_size)
        batch = Transition(*zip(*transitions))
        f = lambda xs: torch.FloatTensor(np.concatenate(xs)).to(self.device)
        return f(batch.state), f(batch.action), f(batch.next_state), f(batch.reward), f(batch.done)

    def __len__(self):
        return len(self.memory)
