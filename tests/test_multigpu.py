"""the N-GPU sharded generation must equal the 1-GPU generation bit for bit (fitness after the
all-gather, selection, post-epoch genomes).  Needs >= 2 GPUs; skipped on a single-GPU box."""
import os
import random

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _generation(rank, world, port, out):
    import torch.distributed as dist
    from serl_b200 import engine, evo, refsig, rollout
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    if world > 1:
        dist.init_process_group('nccl', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world, device_id=dev)
    w = np.load(os.path.join(ROOT, 'tests', 'golden', 'actors.npz'))['serl50_pop8_h32_tanh']
    genomes = torch.as_tensor(np.tile(w, (3, 1))[:21].copy(), device=dev)          # 21 actors: uneven shards
    genomes += 1e-3 * torch.arange(21, device=dev, dtype=torch.float32)[:, None] / 21
    lv, st = refsig.make_ref_params(5, seed_base=2024)
    md = torch.tensor([rollout.mode_code(m) for m in ('nominal', 'be', 'ice', 'cg', 'sa')], dtype=torch.int32, device=dev)
    sh = rollout.actor_shape(32)
    fit, r, (lo, hi) = engine.evaluate_population(genomes, sh, torch.as_tensor(lv, device=dev), torch.as_tensor(st, device=dev), md, horizon=500)
    np.random.seed(5); random.seed(5)
    elite, plan = evo.epoch_flat(genomes, fit, (7, 3, 32, 3))
    out[(world, rank)] = (fit.cpu().numpy(), genomes.cpu().numpy(), elite, (lo, hi))
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_gpu_generation_equals_single_gpu_bitwise():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_generation, args=(1, 0, out), nprocs=1, join=True)
    mp.spawn(_generation, args=(2, 29600 + os.getpid() % 2000, out), nprocs=2, join=True)
    f1, g1, e1, _ = out[(1, 0)]
    for r in range(2):
        f2, g2, e2, (lo, hi) = out[(2, r)]
        assert hi - lo in (10, 11)
        assert np.array_equal(f1, f2)                                    # fitness after the all-gather
        assert e1 == e2
        assert np.array_equal(g1.view(np.uint32), g2.view(np.uint32))    # post-epoch genomes on every rank


def _agent_generation(rank, world, port, out):
    import types
    import torch.distributed as dist
    from serl_b200.core import agent as agent_mod
    from serl_b200.envs import config
    from serl_b200.parameters import Parameters
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    os.makedirs('/tmp/serl_test_mg%d' % rank, exist_ok=True)
    os.chdir('/tmp/serl_test_mg%d' % rank)
    args = Parameters(types.SimpleNamespace(env='PHlab_attitude_nominal', seed=7, pop_size=7, mut_type='normal', test_ea=True))
    args.state_dim, args.action_dim, args.hidden_size = 7, 3, 16
    torch.manual_seed(7); np.random.seed(7); random.seed(7)
    ag = agent_mod.Agent(args, config.select_env('PHlab_attitude_nominal'))
    stats = ag.train()
    out[rank] = (ag.num_frames, ag.num_episodes, ag.pop.genomes.cpu().numpy(), stats['best_train_fitness'], stats['elite_index'])
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_rank_agent_train_stays_in_lockstep():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_agent_generation, args=(2, 29700 + os.getpid() % 2000, out), nprocs=2, join=True)
    a, b = out[0], out[1]
    assert a[0] == b[0] and a[1] == b[1]                       # frame / episode counters
    assert a[3] == b[3] and a[4] == b[4]
    assert np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))


def _agent_rl_generations(rank, world, port, out):
    import types
    import torch.distributed as dist
    from serl_b200.core import agent as agent_mod
    from serl_b200.envs import config
    from serl_b200.parameters import Parameters
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    os.makedirs('/tmp/serl_test_rl%d' % rank, exist_ok=True)
    os.chdir('/tmp/serl_test_rl%d' % rank)
    args = Parameters(types.SimpleNamespace(env='PHlab_attitude_nominal', seed=7, pop_size=5, mut_type='normal', test_ea=False))
    args.state_dim, args.action_dim, args.hidden_size = 7, 3, 16
    args.learn_start, args.frac_frames_train, args.use_caps = 300, 0.01, False      # RL half ON: TD3 updates from the shared buffer
    torch.manual_seed(7); np.random.seed(7); random.seed(7)
    ag = agent_mod.Agent(args, config.select_env('PHlab_attitude_nominal'))
    for _ in range(2):
        stats = ag.train()
    rl = torch.cat([p.detach().reshape(-1) for p in ag.rl_agent.actor.parameters()]).cpu().numpy()
    out[rank] = (ag.num_frames, len(ag.replay_buffer), ag.pop.genomes.cpu().numpy(), rl, ag.rl_iteration, stats['TD_loss'],
                 ag.replay_buffer._chronological_rows().cpu().numpy())
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_rank_training_with_the_rl_half_stays_in_lockstep():
    """with frac_frames_train > 0 every rank must hold the SAME replay buffer (all actors' stored
    transitions are all-gathered), draw the same TD3 batches and end with bit-identical RL weights and genomes."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_agent_rl_generations, args=(2, 29800 + os.getpid() % 2000, out), nprocs=2, join=True)
    a, b = out[0], out[1]
    assert a[0] == b[0] and a[1] == b[1] and a[1] > 300 and a[4] == b[4] and a[4] > 0          # frames, buffer length, TD3 updates ran
    assert np.array_equal(a[6], b[6])                                                        # identical shared replay buffer
    assert np.array_equal(a[3].view(np.uint32), b[3].view(np.uint32))                        # RL actor weights
    assert np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))                        # genomes (incl. the injected RL actor)
    assert np.isfinite(a[5])
