"""Oracle restatement of the reference actor (checker only).

base/core/genetic_agent.py:69-109 (Actor: Linear(S,h), act, L x [Linear(h,h), LayerNorm(h), act], Linear(h,A), Tanh)
base/core/mod_utils.py:39-50 (LayerNorm: unbiased std, eps added to std), :14-18 (activations; 'relu' -> LeakyReLU).
Flat genome layout = order of nn.Module.parameters() (what SSNE.clone / crossover iterate over):
  net.0.weight[h,S] net.0.bias[h]  { net.k.weight[h,h] net.k.bias[h] net.k+1.gamma[h] net.k+1.beta[h] } x L
  net.last.weight[A,h] net.last.bias[A]
"""
import numpy as np
import torch
import torch.nn as nn


class LayerNorm(nn.Module):
    def __init__(self, features, eps=1e-6):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(features))
        self.beta = nn.Parameter(torch.zeros(features))
        self.eps = eps

    def forward(self, x):
        mean = x.mean(-1, keepdim=True)
        std = x.std(-1, keepdim=True)
        return self.gamma * (x - mean) / (std + self.eps) + self.beta


def _act(name):
    return {'tanh': nn.Tanh(), 'elu': nn.ELU(), 'relu': nn.LeakyReLU()}[name.lower()]


class Actor(nn.Module):
    def __init__(self, state_dim=7, action_dim=3, hidden=72, num_layers=3, activation='tanh'):
        super().__init__()
        layers = [nn.Linear(state_dim, hidden), _act(activation)]
        for _ in range(num_layers):
            layers.extend([nn.Linear(hidden, hidden), LayerNorm(hidden), _act(activation)])
        layers.extend([nn.Linear(hidden, action_dim), nn.Tanh()])
        self.net = nn.Sequential(*layers)
        self.dims = (state_dim, action_dim, hidden, num_layers, activation)

    def forward(self, x):
        return self.net(x)

    def select_action(self, state):
        state = torch.FloatTensor(np.asarray(state).reshape(1, -1))
        with torch.no_grad():
            return self.forward(state).cpu().data.numpy().flatten()


def num_params(state_dim, action_dim, hidden, num_layers):
    return state_dim * hidden + hidden + num_layers * (hidden * hidden + 3 * hidden) + hidden * action_dim + action_dim


def flatten(actor):
    return torch.cat([p.data.reshape(-1) for p in actor.parameters()]).numpy().astype(np.float32)


def unflatten(vec, state_dim=7, action_dim=3, hidden=72, num_layers=3, activation='tanh'):
    a = Actor(state_dim, action_dim, hidden, num_layers, activation)
    off = 0
    v = torch.as_tensor(np.asarray(vec, dtype=np.float32))
    for p in a.parameters():
        n = p.numel()
        p.data.copy_(v[off:off + n].reshape(p.shape))
        off += n
    assert off == v.numel()
    return a


def from_state_dict(sd, activation):
    """Build an oracle Actor from a reference checkpoint (keys net.{0,2,5,8,11}.{weight,bias}, net.{3,6,9}.{gamma,beta})."""
    hidden, state_dim = sd['net.0.weight'].shape
    lin = sorted({int(k.split('.')[1]) for k in sd if k.endswith('.weight')})
    action_dim = sd['net.%d.weight' % lin[-1]].shape[0]
    a = Actor(state_dim, action_dim, hidden, len(lin) - 2, activation)
    a.load_state_dict(sd)
    return a


class WideActor(nn.Module):
    """Width-list generalisation used by BASELINE config 5 ([400,300], [128,128]): Linear(S,w1), act,
    {Linear(w_i,w_{i+1}), LayerNorm, act} ..., Linear(w_n,A), Tanh — the reference's Actor is the case of equal widths."""

    def __init__(self, widths, state_dim=7, action_dim=3, activation='tanh'):
        super().__init__()
        layers = [nn.Linear(state_dim, widths[0]), _act(activation)]
        for a, b in zip(widths[:-1], widths[1:]):
            layers.extend([nn.Linear(a, b), LayerNorm(b), _act(activation)])
        layers.extend([nn.Linear(widths[-1], action_dim), nn.Tanh()])
        self.net = nn.Sequential(*layers)

    def forward(self, x):
        return self.net(x)


def num_params_wide(widths, state_dim=7, action_dim=3):
    n = state_dim * widths[0] + widths[0]
    for a, b in zip(widths[:-1], widths[1:]):
        n += a * b + 3 * b
    return n + widths[-1] * action_dim + action_dim


def unflatten_wide(vec, widths, activation='tanh'):
    a = WideActor(widths, activation=activation)
    off = 0
    v = torch.as_tensor(np.asarray(vec, dtype=np.float32))
    for p in a.parameters():
        n = p.numel()
        p.data.copy_(v[off:off + n].reshape(p.shape))
        off += n
    assert off == v.numel()
    return a
