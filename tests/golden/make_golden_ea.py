"""Run the REFERENCE SSNE (base/core/mod_neuro_evo.py, imported from /root/reference) with the exclusive-index shim
(random.randint -> randrange at lines 51,76,79,89,92,357,358,517) and record before/after genomes.
Container-only; output tests/golden/ssne_kat.npz."""
import os, sys, random, types
import numpy as np, torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/base')
os.chdir('/tmp')
from core import mod_neuro_evo as ne, genetic_agent          # noqa: E402
from parameters import Parameters                             # noqa: E402

IDX_LINES = {51, 76, 79, 89, 92, 357, 358, 517}


class Shim:
    def __getattr__(self, k):
        return getattr(random, k)

    def randint(self, a, b):
        if sys._getframe(1).f_lineno in IDX_LINES:
            return random.randrange(a, b)
        return random.randint(a, b)


ne.random = Shim()


def make_args(pop, hidden, layers):
    cla = types.SimpleNamespace(disable_cuda=True, env='phlab_attitude_nominal', seed=7, pop_size=pop, mut_type='normal')
    args = Parameters(cla)
    args.state_dim, args.action_dim = 7, 3
    args.hidden_size, args.num_layers = hidden, layers
    args.distil_crossover = False
    return args


def run_case(pop, hidden, layers, seed, ties=False):
    args = make_args(pop, hidden, layers)
    torch.manual_seed(seed); np.random.seed(seed); random.seed(seed)
    agents = [genetic_agent.GeneticAgent(args) for _ in range(pop)]
    flat = lambda: np.stack([torch.cat([p.data.reshape(-1) for p in a.actor.parameters()]).numpy().copy() for a in agents])
    before = flat()
    fit = np.random.uniform(-3000, -50, pop)
    ev = ne.SSNE(args, None, None)
    np.random.seed(seed + 1); random.seed(seed + 2)
    elite = ev.epoch(agents, fit)
    return dict(before=before, fitness=fit, after=flat(), elite=np.int64(elite), seed=np.int64(seed),
                shape=np.array([7, 3, hidden, layers]))


if __name__ == '__main__':
    out = {}
    for k, (pop, h, L, seed) in enumerate([(10, 8, 3, 1), (10, 8, 3, 2), (16, 12, 2, 3), (50, 16, 3, 4), (7, 8, 1, 5), (6, 8, 3, 6)]):
        c = run_case(pop, h, L, seed)
        for name, v in c.items():
            out['c%d_%s' % (k, name)] = v
        print(k, pop, h, L, 'elite', c['elite'], 'changed rows', int((c['before'] != c['after']).any(1).sum()))
    np.savez_compressed(os.path.join(HERE, 'ssne_kat.npz'), **out)
