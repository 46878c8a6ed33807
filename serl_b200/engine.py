"""Population evaluation across 1..N GPUs (one process per GPU, torch.distributed).

The (actor, env) trajectories are independent; the only coupling is the ranking in SSNE.epoch
(base/core/mod_neuro_evo.py:460).  Rank r flies the contiguous actor block shard_bounds(pop, world, r) through all
environments; one all-gather of the per-actor fitness (pop x 8 bytes, 4 KB at pop=512) makes the full fitness vector
available on every rank, which then runs the (deterministic, identically seeded) evolution step redundantly —
genomes stay bit-identical on all ranks without ever moving weights.
"""
import torch
import torch.distributed as dist

from . import rollout


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def shard_bounds(pop, world, rank):
    """contiguous block partition, first (pop % world) ranks get one extra actor."""
    base, rem = divmod(pop, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local, pop, world, rank, group=None):
    """all-gather of unequal contiguous actor blocks: local [n_local, ...] -> [pop, ...] on every rank."""
    if world == 1:
        return local
    blk = (pop + world - 1) // world
    pad = torch.zeros((blk,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world * blk,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if out.is_cuda:
        dist.all_gather_into_tensor(out, pad, group=group)
    else:
        dist.all_gather(list(out.view((world, blk) + tuple(local.shape[1:])).unbind(0)), pad, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(pop, world, r)
        parts.append(out[r * blk:r * blk + (hi - lo)])
    return torch.cat(parts)


def gather_fitness(local, pop, world, rank, group=None):
    """fitness[pop] on every rank (same dtype/device as `local`): the one collective of the path."""
    return gather_rows(local, pop, world, rank, group)


def evaluate_population(genomes, shape, ref_levels, ref_starts, env_mode, horizon=rollout.HORIZON, group=None,
                        actions=False, smooth_fitness=False):
    """fitness[pop] (f64, on device, identical on every rank), local RolloutResult, (lo, hi) of this rank's actor block.
    actions=True also records the commanded deflections and computes the per-trajectory smoothness (K6) as
    result.smoothness [pop_local, n_envs]; smooth_fitness adds it to every episode's return before the mean
    (base/core/agent.py:128-134)."""
    world, rank = world_info()
    pop = genomes.shape[0]
    lo, hi = shard_bounds(pop, world, rank)
    if hi > lo:
        r = rollout.population_rollout(genomes[lo:hi], shape, ref_levels, ref_starts, env_mode, horizon=horizon,
                                       actions=actions or smooth_fitness)
        local = r.fitness
        r.smoothness = None
        if actions or smooth_fitness:
            r.smoothness = rollout.smoothness(r.actions, r.steps)
            r.actions = None                     # 48 KB per trajectory: free it as soon as the metric exists
            if smooth_fitness:
                local = (r.returns + r.smoothness).mean(dim=1)
    else:
        r = None
        local = torch.zeros(0, dtype=torch.float64, device=genomes.device)
    return gather_fitness(local, pop, world, rank, group), r, (lo, hi)
