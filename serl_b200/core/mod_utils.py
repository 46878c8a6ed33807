"""Host-side helpers with the names the reference's callers import from base/core/mod_utils.py
(activations :14-18, soft/hard target updates :25-34, LayerNorm :39-50, Gaussian exploration noise, is_lnorm_key :130)."""
import numpy as np
import torch
from torch import nn

# name -> module; note the reference maps 'relu' to LeakyReLU (mod_utils.py:17)
activations = dict(tanh=nn.Tanh(), elu=nn.ELU(), relu=nn.LeakyReLU())


@torch.no_grad()
def soft_update(target, source, tau):
    """Polyak averaging target <- (1 - tau) * target + tau * source."""
    for t, s in zip(target.parameters(), source.parameters()):
        t.mul_(1.0 - tau).add_(s, alpha=tau)


@torch.no_grad()
def hard_update(target, source):
    for t, s in zip(target.parameters(), source.parameters()):
        t.copy_(s)


class LayerNorm(nn.Module):
    """The reference's own normalisation (not torch.nn.LayerNorm): Bessel-corrected standard deviation over the last
    axis with eps added to the *standard deviation*, then an affine map (parameters named gamma / beta so that
    checkpoints keep the keys net.{3,6,9}.{gamma,beta})."""

    def __init__(self, features, eps=1e-6):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(features))
        self.beta = nn.Parameter(torch.zeros(features))
        self.eps = eps

    def forward(self, x):
        centred = x - x.mean(dim=-1, keepdim=True)
        spread = x.std(dim=-1, keepdim=True) + self.eps
        return self.gamma * centred / spread + self.beta


class GaussianNoise:
    def __init__(self, action_dimension, sd=0.1, mu=0):
        self.action_dimension, self.sd, self.mu = action_dimension, sd, mu

    def reset(self):
        pass

    def noise(self):
        return np.random.normal(self.mu, self.sd, self.action_dimension)


def is_lnorm_key(key):
    return key.startswith('lnorm')
