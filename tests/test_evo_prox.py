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


@pytest.mark.skipif(not os.path.isdir(REF), reason='needs the reference tree (build container only)')
def test_batched_distillation_step_equals_the_reference_update_parameters(tmp_path, monkeypatch):
    """one Q-filtered behaviour-cloning Adam step (base/core/genetic_agent.py:22-60) for three children at once."""
    from serl_b200 import evo_distil
    monkeypatch.chdir(tmp_path)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == 'core' or k.startswith('core.') or k == 'parameters'}
    sys.path.insert(0, REF)
    try:
        from core import genetic_agent as ref_ga
        from parameters import Parameters as RefP
        args = RefP(types.SimpleNamespace(pop_size=4, mut_type='proximal', env='x', frames=1, seed=1, disable_cuda=True))
        args.state_dim, args.action_dim, args.device = 7, 3, torch.device('cpu')
        torch.manual_seed(0)
        kids = [ref_ga.GeneticAgent(args) for _ in range(3)]
        p1s = [ref_ga.GeneticAgent(args) for _ in range(3)]
        p2s = [ref_ga.GeneticAgent(args) for _ in range(3)]
        flat = lambda g: torch.cat([p.data.reshape(-1) for p in g.actor.parameters()])
        lin = torch.nn.Linear(10, 2)

        def critic(s, a):
            q = lin(torch.cat((s, a), 1))
            return q[:, :1], q[:, 1:]
        states = torch.randn(3, 40, 7) * 0.2
        shape = (7, 3, args.hidden_size, args.num_layers)
        G0 = torch.stack([flat(k) for k in kids])
        G1, G2 = torch.stack([flat(p) for p in p1s]), torch.stack([flat(p) for p in p2s])
        mse_ref = [kids[c].update_parameters((states[c], None, None, None, None), p1s[c].actor, p2s[c].actor, critic) for c in range(3)]
        G_ref = torch.stack([flat(k) for k in kids])
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == 'core' or k.startswith('core.') or k == 'parameters']:
            del sys.modules[k]
        sys.modules.update(saved)
    child = G0.clone().requires_grad_(True)
    opt = torch.optim.Adam([child], lr=1e-3)
    with torch.no_grad():
        a1 = evo_prox.actor_forward_batched(G1, states, shape, 'tanh')
        a2 = evo_prox.actor_forward_batched(G2, states, shape, 'tanh')
        fl = states.reshape(120, 7)
        q1 = torch.min(*critic(fl, a1.reshape(120, 3))).reshape(3, 40)
        q2 = torch.min(*critic(fl, a2.reshape(120, 3))).reshape(3, 40)
    opt.zero_grad()
    loss, mse = evo_distil.cloning_loss(evo_prox.actor_forward_batched(child, states, shape, 'tanh'), a1, a2, q1, q2)
    loss.backward()
    opt.step()
    assert (G_ref - G0).abs().max() > 1e-4
    assert (child.detach() - G_ref).abs().max().item() <= 2e-6
    assert np.allclose(mse.numpy(), np.asarray(mse_ref), rtol=1e-4)


def test_sort_groups_by_fitness_order():
    from serl_b200 import evo_distil
    fit = {3: -10.0, 5: -2.0, 9: -7.0}
    g = evo_distil.sort_groups_by_fitness([3, 5, 9], fit)
    assert g[0][:2] == (5, 9) and g[-1][:2] == (9, 3) and g[0][2] == -9.0
