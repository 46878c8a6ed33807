"""Proximal / safe mutation (base/core/mod_neuro_evo.py:183-252, :254-327) batched over the population.

The reference mutates one actor at a time: sample a batch of that actor's own states, build the Jacobian of the three
outputs (summed over the batch) w.r.t. every weight MATRIX with three autograd passes, scale a Gaussian perturbation
delta ~ N(0, mag) by 1 / clamp(sqrt(sum_i jac_i^2)) and add it.  Here all actors mutated in a generation go through ONE
batched forward (bmm over the flat [n, P] genome rows) and three batched backward passes on the device; the genome matrix
is updated in place.  `safe` differs only in which buffer the states come from (the critical buffer when it holds more
than one transition, :258-261)."""
import torch

from . import evo

LAM_MAX = 0.01


def _views(G, table):
    """[(kind, tensor view)] of every parameter of the [n, P] genome rows, in parameters() order."""
    out = []
    for off, rows, cols in table:
        if cols > 0:
            out.append(('W', G[:, off:off + rows * cols].reshape(G.shape[0], rows, cols)))
        else:
            out.append(('v', G[:, off:off + rows]))
    return out


def actor_forward_batched(G, states, shape, activation='tanh'):
    """G [n, P] genomes, states [n, B, S] -> actions [n, B, A]: Actor.forward (genetic_agent.py:78-109) with per-actor weights."""
    state_dim, action_dim, hidden, num_layers = shape
    table, P = evo.param_table(*shape)
    assert G.shape[1] == P
    act = {'tanh': torch.tanh, 'elu': torch.nn.functional.elu, 'relu': torch.nn.functional.leaky_relu}[activation.lower()]
    v = _views(G, table)
    x = act(torch.baddbmm(v[1][1].unsqueeze(1), states, v[0][1].transpose(1, 2)))
    i = 2
    for _ in range(num_layers):
        W, b, gamma, beta = v[i][1], v[i + 1][1], v[i + 2][1], v[i + 3][1]
        i += 4
        x = torch.baddbmm(b.unsqueeze(1), x, W.transpose(1, 2))
        mean = x.mean(-1, keepdim=True)
        std = x.std(-1, keepdim=True)
        x = act(gamma.unsqueeze(1) * (x - mean) / (std + 1e-6) + beta.unsqueeze(1))
    return torch.tanh(torch.baddbmm(v[i + 1][1].unsqueeze(1), x, v[i][1].transpose(1, 2)))


def weight_mask(shape, device):
    """1 at the positions of the 2-D parameters (what extract_parameters / extract_grad concatenate, genetic_agent.py:111-135)."""
    table, P = evo.param_table(*shape)
    m = torch.zeros(P, dtype=torch.bool, device=device)
    for off, rows, cols in table:
        if cols > 0:
            m[off:off + rows * cols] = True
    return m


def proximal_mutate_batched(genomes, actor_idx, states, shape, activation, mag, delta=None, generator=None):
    """Mutate genomes[actor_idx] in place.  states [n, B, S] (each actor's own batch).  delta: optional [n, n_weights]
    perturbation (tests); otherwise drawn N(0, mag) on the genomes' device.  Returns the scaling that was applied."""
    idx = torch.as_tensor(actor_idx, dtype=torch.int64, device=genomes.device)
    G = genomes[idx].clone().requires_grad_(True)
    out = actor_forward_batched(G, states, shape, activation)                       # [n, B, A]
    mask = weight_mask(shape, genomes.device)
    jac2 = torch.zeros((G.shape[0], int(mask.sum())), dtype=G.dtype, device=G.device)
    for i in range(out.shape[2]):                                                  # one backward pass per output (:207-213)
        (g,) = torch.autograd.grad(out[:, :, i].sum(), G, retain_graph=True)
        jac2 += g[:, mask] ** 2
    scaling = torch.sqrt(jac2)                                                     # summed-gradient sensitivity (:216)
    scaling[scaling == 0] = 1.0
    scaling[scaling < LAM_MAX] = LAM_MAX
    if delta is None:
        delta = torch.randn(scaling.shape, dtype=G.dtype, device=G.device, generator=generator) * mag
    new_w = G.detach()[:, mask] + delta / scaling
    rows = genomes[idx]
    rows[:, mask] = new_w
    genomes[idx] = rows
    return scaling


def mutation_states(pop, actor_idx, batch_size, safe=False):
    """each mutated actor's own state batch (min(batch_size, len(buffer)) states, padded by repetition to a common B)."""
    batches = []
    for i in actor_idx:
        buf = pop[int(i)].critical_buffer if safe and len(pop[int(i)].critical_buffer) > 1 else pop[int(i)].buffer
        n = len(buf)
        if n == 0:
            raise RuntimeError('proximal / safe mutation: actor %d has an empty replay buffer (the reference would fail in '
                               'buffer.sample); fly at least one generation with stored transitions first' % int(i))
        s = buf.sample(min(batch_size, n))[0]
        if s.shape[0] < batch_size:      # the Jacobian is a SUM over the batch: repeat rows with weight -> use exact tiling only
            reps = -(-batch_size // s.shape[0])
            s = s.repeat(reps, 1)[:batch_size] if s.shape[0] * reps == batch_size else s
        batches.append(s)
    B = min(b.shape[0] for b in batches)
    return torch.stack([b[:B] for b in batches])
