"""Mirror of base/core/genetic_agent.py: Actor (:69-163) and GeneticAgent (:10-63).

The network definition is identical (same module tree -> same state_dict keys net.{0,2,5,8,11}.{weight,bias},
net.{3,6,9}.{gamma,beta}; same parameters() order = the flat genome layout of the engine).  When an actor belongs to a
`Population`, its parameters are *views* into one row of the device-resident [pop, P] genome matrix, so state_dict(),
rl_to_evo copies (agent.py:140-142) and torch.save keep working while the kernels read / write the matrix directly.
"""
import numpy as np
import torch
import torch.nn as nn

from .mod_utils import LayerNorm, activations, is_lnorm_key
from . import replay_memory


class Actor(nn.Module):
    def __init__(self, args, init=False):
        super().__init__()
        self.args = args
        h, L = args.hidden_size, args.num_layers
        activation = activations[args.activation_actor.lower()]
        layers = [nn.Linear(args.state_dim, h), activation]
        for _ in range(L):
            layers.extend([nn.Linear(h, h), LayerNorm(h), activation])
        layers.extend([nn.Linear(h, args.action_dim), nn.Tanh()])
        self.net = nn.Sequential(*layers)

    def forward(self, state: torch.Tensor) -> torch.Tensor:
        return self.net(state)

    def select_action(self, state):
        dev = next(self.parameters()).device
        state = torch.as_tensor(np.asarray(state).reshape(1, -1), dtype=torch.float32, device=dev)
        with torch.no_grad():
            return self.forward(state).cpu().data.numpy().flatten()

    def get_novelty(self, batch):
        state_batch, action_batch, _, _, _ = batch
        novelty = torch.mean(torch.sum((action_batch - self.forward(state_batch)) ** 2, dim=-1))
        self.novelty = novelty.item()
        return self.novelty

    def count_parameters(self):
        return sum(p.numel() for n, p in self.named_parameters() if not is_lnorm_key(n) and len(p.shape) == 2)

    def extract_parameters(self):
        return torch.cat([p.detach().view(-1) for n, p in self.named_parameters()
                          if not is_lnorm_key(n) and len(p.shape) == 2]).clone()

    def inject_parameters(self, pvec):
        count = 0
        for n, p in self.named_parameters():
            if is_lnorm_key(n) or len(p.shape) != 2:
                continue
            sz = p.numel()
            p.data.copy_(pvec[count:count + sz].view(p.size()))
            count += sz

    # ---- flat genome helpers (engine side) ----
    def flat(self):
        return torch.cat([p.data.reshape(-1) for p in self.parameters()])

    def bind(self, row: torch.Tensor):
        """Re-seat every parameter as a view into `row` (one row of the population genome matrix)."""
        off = 0
        for p in self.parameters():
            n = p.numel()
            p.data = row[off:off + n].view(p.shape)
            off += n
        assert off == row.numel()


class GeneticAgent:
    def __init__(self, args):
        self.args = args
        self.actor = Actor(args)
        self.buffer = replay_memory.ReplayMemory(self.args.individual_bs, args.device)
        self.critical_buffer = replay_memory.ReplayMemory(self.args.individual_bs, args.device)

    def load_from_dict(self, actor_dict: dict):
        self.actor = actor_dict
