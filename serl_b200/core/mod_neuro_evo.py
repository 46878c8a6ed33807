"""Mirror of base/core/mod_neuro_evo.py SSNE (:14-543) on top of the device kernels K2-K5 (serl_b200/evo.py).

Same constructor and `epoch(pop, fitness_evals, bcs_evals=None) -> int` contract; `pop` must be the engine's
PopulationList (its genomes are mutated in place on the GPU).  Mutation operators: classic Gaussian point mutation
('normal' / 'inplace', K5) and the Jacobian-scaled 'proximal' / 'safe' mutations (:183-327) batched over the population
(serl_b200/evo_prox.py, exact against the reference module).  Crossover: the classic in-place operator (K4) or, with
`distil_crossover`, the Q-filtered distillation crossover (:131-181) batched over all children (serl_b200/evo_distil.py;
`distil_type` 'fitness'; other pairing rules raise NotImplementedError as the reference does for unknown ones, :507).
"""
import random

import torch

from .. import evo, evo_distil, evo_prox


class SSNE:
    def __init__(self, args, critic, evaluate):
        self.current_gen = 0
        self.args = args
        self.critic = critic
        self.population_size = self.args.pop_size
        self.num_elitists = max(int(self.args.elite_fraction * args.pop_size), 1)
        self.evaluate = evaluate
        self.rl_policy = None
        self.selection_stats = {'elite': 0, 'selected': 0, 'discarded': 0, 'total': 0.0000001}
        if self.args.mut_type in ('normal', 'inplace'):
            self.mutate = None       # classic mutation runs inside epoch() (K5)
        elif self.args.mut_type in ('proximal', 'safe'):
            self.mutate = self.args.mut_type      # batched over the population inside epoch() (serl_b200/evo_prox.py)
            self._mut_gen = None
        else:
            raise ValueError('Mutation type is unknown!')
        self.distil = bool(getattr(self.args, 'distil_crossover', False))
        if self.distil and 'fitness' not in str(getattr(self.args, 'distil_type', 'fitness')).lower():
            raise NotImplementedError("distillation crossover: only distil_type 'fitness' (mod_neuro_evo.py:498-499) is implemented")
        self._gen = None
        self.last_plan = None

    def _selection_bookkeeping(self, elitist_index, offsprings, unselects):
        # mod_neuro_evo.py:478-485
        if self.rl_policy is not None:
            self.selection_stats['total'] += 1.0
            if self.rl_policy in elitist_index:
                self.selection_stats['elite'] += 1.0
            elif self.rl_policy in offsprings:
                self.selection_stats['selected'] += 1.0
            elif self.rl_policy in unselects:
                self.selection_stats['discarded'] += 1.0
            self.rl_policy = None

    def epoch(self, pop, fitness_evals, bcs_evals=None):
        genomes = getattr(pop, 'genomes', None)
        if genomes is None:
            raise TypeError('SSNE.epoch needs the engine population (serl_b200.population.PopulationList)')
        classic = self.mutate is None
        if self._gen is None:
            self._gen = torch.Generator(device=genomes.device)
            self._gen.manual_seed(int(getattr(self.args, 'seed', 7)) + 2)
        fit_host = fitness_evals.detach().cpu().numpy() if torch.is_tensor(fitness_evals) else fitness_evals

        def distil(plan):
            # :497-513: one child per unselected actor, parents = the pairs of (new elitists + offsprings) ranked by
            # summed fitness; all children of the generation are trained together
            groups = evo_distil.sort_groups_by_fitness(plan.new_elitists + plan.offsprings, fit_host)
            if not plan.distil_unselects or not groups:
                return
            first, second, bufs = [], [], []
            for i, _ in enumerate(plan.distil_unselects):
                a, b, _s = groups[i % len(groups)]
                if fit_host[a] < fit_host[b]:
                    a, b = b, a
                first.append(int(a)); second.append(int(b))
                bufs.append(evo_distil.child_buffer(pop, int(a), int(b), int(self.args.individual_bs) // 2, self._gen))
            children = evo_distil.distil_children(genomes, first, second, bufs, pop.shape_tuple, self.args.activation_actor,
                                                  self.critic, generator=self._gen)
            idx = torch.as_tensor(plan.distil_unselects, dtype=torch.int64, device=genomes.device)
            genomes[idx] = children                                          # clone(offspring, pop[unselected]) :513
            for k, u in enumerate(plan.distil_unselects):                    # ... including the child's buffer (:378-382)
                pop[u].buffer.reset()
                pop[u].buffer.add_rows(bufs[k])
                pop[u].critical_buffer.reset()
        elite, plan = evo.epoch_flat(genomes, fitness_evals, pop.shape_tuple,
                                     elite_fraction=self.args.elite_fraction, mutation_prob=self.args.mutation_prob,
                                     mutation_mag=self.args.mutation_mag, selection=self._selection_bookkeeping,
                                     classic_mutation=classic, classic_crossover=not self.distil,
                                     between=distil if self.distil else None)
        self.last_plan = plan
        # clone() also copies the per-agent replay buffers (:377-382); host-side bookkeeping, in the reference's order
        for wave in plan.clone_waves:
            for src, dst in wave:
                _copy_buffers(pop[int(src)], pop[int(dst)])
        for desc, _ in plan.cross_waves:
            for g1, g2, s1, s2, _, _ in desc:
                _copy_buffers(pop[int(s1)], pop[int(g1)])
                _copy_buffers(pop[int(s2)], pop[int(g2)])
        if not classic:
            # :537-539: every non-elite rank mutates with probability mutation_prob (one random.random() each, in rank order)
            chosen = [i for i in plan.mut_candidates if random.random() < self.args.mutation_prob]
            if chosen:
                if self._mut_gen is None:
                    self._mut_gen = torch.Generator(device=genomes.device)
                    self._mut_gen.manual_seed(int(getattr(self.args, 'seed', 7)) + 1)
                states = evo_prox.mutation_states(pop, chosen, int(self.args.mutation_batch_size), safe=self.mutate == 'safe')
                evo_prox.proximal_mutate_batched(genomes, chosen, states, pop.shape_tuple, self.args.activation_actor,
                                                 float(self.args.mutation_mag), generator=self._mut_gen)
        self.current_gen += 1
        return elite


def _copy_buffers(master, replacee):
    if master is replacee:
        return
    owner = getattr(master.buffer, 'owner', None)
    if owner is not None and owner is getattr(replacee.buffer, 'owner', None):      # device-resident per-actor rings
        owner.copy_actor(master.buffer.index, replacee.buffer.index)
        master.critical_buffer.owner.copy_actor(master.critical_buffer.index, replacee.critical_buffer.index)
        return
    replacee.buffer.reset()
    replacee.buffer.add_content_of(master.buffer)
    replacee.critical_buffer.reset()
    replacee.critical_buffer.add_content_of(master.critical_buffer)
