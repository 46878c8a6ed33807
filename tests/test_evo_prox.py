"""N3: proximal / safe mutation batched over the population (serl_b200/evo_prox.py) against (a) the reference module itself
(base/core/mod_neuro_evo.py:183-252, imported in the build container) and (b) a per-actor autograd restatement."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from serl_b200 import evo, evo_prox

REF = '/root/reference/base'


def per_actor_reference(genome, states, shape, activation, mag, delta):
    """the reference's algorithm on ONE actor with plain autograd over a functional forward."""
    G = genome.clone().reshape(1, -1).requires_grad_(True)
    out = evo_prox.actor_forward_batched(G, states[None], shape, activation)[0]
    mask = evo_prox.weight_mask(shape, G.device)
    jac = []
    for i in range(3):
        (g,) = torch.autograd.grad(out[:, i].sum(), G, retain_graph=True)
        jac.append(g[0, mask])
    scaling = torch.sqrt(sum(j ** 2 for j in jac))
    scaling[scaling == 0] = 1.0
    scaling[scaling < 0.01] = 0.01
    new = genome.clone()
    new[mask] = genome[mask] + delta / scaling
    return new


def test_batched_equals_per_actor_restatement():
    torch.manual_seed(3)
    shape = (7, 3, 32, 3)
    table, P = evo.param_table(*shape)
    G = torch.randn(6, P) * 0.2
    states = torch.randn(4, 20, 7) * 0.1
    idx = [5, 0, 3, 2]
    nw = int(evo_prox.weight_mask(shape, G.device).sum())
    delta = torch.randn(4, nw) * 0.02
    G2 = G.clone()
    evo_prox.proximal_mutate_batched(G2, idx, states, shape, 'tanh', 0.02, delta=delta)
    for k, i in enumerate(idx):
        want = per_actor_reference(G[i], states[k], shape, 'tanh', 0.02, delta[k])
        assert torch.allclose(G2[i], want, rtol=1e-5, atol=1e-6)
    assert torch.equal(G2[1], G[1]) and torch.equal(G2[4], G[4])          # untouched actors
    # biases / LayerNorm parameters are not mutated (extract_parameters takes 2-D parameters only, genetic_agent.py:125-135)
    m = evo_prox.weight_mask(shape, G.device)
    assert torch.equal(G2[:, ~m], G[:, ~m])


@pytest.mark.skipif(not os.path.isdir(REF), reason='needs the reference tree (build container only)')
def test_batched_proximal_mutation_equals_the_reference_module(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == 'core' or k.startswith('core.') or k == 'parameters'}
    sys.path.insert(0, REF)
    try:
        from core import mod_neuro_evo as ref_ne, genetic_agent as ref_ga
        from parameters import Parameters as RefP
        import torch.distributions as dist
        args = RefP(types.SimpleNamespace(pop_size=4, mut_type='proximal', env='x', frames=1, seed=1, disable_cuda=True))
        args.state_dim, args.action_dim, args.device = 7, 3, torch.device('cpu')
        torch.manual_seed(0)
        genes = [ref_ga.GeneticAgent(args) for _ in range(3)]
        shape = (7, 3, args.hidden_size, args.num_layers)
        G = torch.stack([torch.cat([p.data.reshape(-1) for p in g.actor.parameters()]) for g in genes])
        states = torch.randn(3, 32, 7) * 0.1
        ssne = ref_ne.SSNE(args, None, None)

        class FakeBuf:
            def __init__(self, st):
                self.st = st

            def __len__(self):
                return 32

            def sample(self, n):
                return (self.st, None, None, None, None)
        deltas = []
        for k, g in enumerate(genes):
            g.buffer = FakeBuf(states[k])
            tot = g.actor.count_parameters()
            torch.manual_seed(100 + k)
            deltas.append(dist.Normal(torch.zeros(tot), torch.ones(tot) * args.mutation_mag).sample())
            torch.manual_seed(100 + k)
            ssne.proximal_mutate(g, mag=args.mutation_mag)
        G_ref = torch.stack([torch.cat([p.data.reshape(-1) for p in g.actor.parameters()]) for g in genes])
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == 'core' or k.startswith('core.') or k == 'parameters']:
            del sys.modules[k]
        sys.modules.update(saved)
    G2 = G.clone()
    evo_prox.proximal_mutate_batched(G2, [0, 1, 2], states, shape, args.activation_actor, args.mutation_mag, delta=torch.stack(deltas))
    assert (G_ref - G).abs().max() > 0.1
    assert (G2 - G_ref).abs().max().item() <= 1e-6
