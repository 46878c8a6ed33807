"""Mirror of the pieces of base/core/mod_utils.py the hot path and its callers use."""
import numpy as np
import torch
import torch.nn as nn

# base/core/mod_utils.py:14-18 ('relu' is LeakyReLU in the reference)
activations = {'tanh': nn.Tanh(), 'elu': nn.ELU(), 'relu': nn.LeakyReLU()}


def soft_update(target, source, tau):
    for target_param, param in zip(target.parameters(), source.parameters()):
        target_param.data.copy_(target_param.data * (1.0 - tau) + param.data * tau)


def hard_update(target, source):
    for target_param, param in zip(target.parameters(), source.parameters()):
        target_param.data.copy_(param.data)


class LayerNorm(nn.Module):
    """base/core/mod_utils.py:39-50: unbiased std, eps added to the std."""

    def __init__(self, features, eps=1e-6):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(features))
        self.beta = nn.Parameter(torch.zeros(features))
        self.eps = eps

    def forward(self, x):
        mean = x.mean(-1, keepdim=True)
        std = x.std(-1, keepdim=True)
        return self.gamma * (x - mean) / (std + self.eps) + self.beta


class GaussianNoise:
    def __init__(self, action_dimension, sd=0.1, mu=0):
        self.action_dimension, self.sd, self.mu = action_dimension, sd, mu

    def reset(self):
        pass

    def noise(self):
        return np.random.normal(self.mu, self.sd, self.action_dimension)


def is_lnorm_key(key):
    return key.startswith('lnorm')
