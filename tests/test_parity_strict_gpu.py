"""STRICT parity of the CUDA rollout (through the C-ABI) against the oracle whose actor follows the kernel's summation
order and activation arithmetic (oracle/plant/actor_kernel_order.c) and whose plant / wrapper follow the reference
(oracle/plant/plant_oracle.c = bit-identical restatement of the reference binaries; episode.c = envs/phlabenv.py).

Bar (BASELINE.json north_star, no escape hatch): EVERY trajectory has the identical termination step and an episodic
return within 1e-4 relative.  The actor itself must agree BIT FOR BIT.  The comparison with the reference-order
(torch-like) forward pass is reported as the reference's own float32 self-sensitivity, not used as a tolerance."""
import os

import numpy as np
import pytest
import torch

from oracle import actor as A, fast, refsig

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden')
ACT = np.load(os.path.join(G, 'actors.npz'))
REL_TOL = 1e-4
MODES10 = ['nominal', 'be', 'jr', 'sa', 'se', 'ice', 'cg', 'cg-for', 'h2000-v150', 'h10000-v90']


def gpu_rollout(weights, hidden, activation, levels, starts, modes, num_layers=3, sort=False, **kw):
    from serl_b200 import rollout
    dev = torch.device('cuda:0')
    sh = rollout.actor_shape(hidden, num_layers, activation)
    w = torch.as_tensor(np.ascontiguousarray(weights, dtype=np.float32), device=dev)
    md = torch.as_tensor(np.array([rollout.mode_code(m) for m in modes], dtype=np.int32), device=dev)
    order = rollout.variant_sorted_order(md) if sort else None
    r = rollout.population_rollout(w, sh, torch.as_tensor(levels, device=dev), torch.as_tensor(starts, device=dev), md, env_order=order, **kw)
    torch.cuda.synchronize()
    r.check()
    return r


def strict_check(r, oret, ostp):
    ret, stp = r.returns.cpu().numpy(), r.steps.cpu().numpy()
    bad = np.argwhere(stp != ostp)
    assert bad.size == 0, ('termination step mismatch', bad[:5], stp[tuple(bad[0])], ostp[tuple(bad[0])])
    rel = np.abs(ret - oret) / np.abs(oret)
    assert rel.max() <= REL_TOL, ('return mismatch', np.unravel_index(rel.argmax(), rel.shape), rel.max())
    return rel.max()


def random_genomes(n, hidden, activation, seed, num_layers=3):
    torch.manual_seed(seed)
    return np.stack([A.flatten(A.Actor(hidden=hidden, num_layers=num_layers, activation=activation)) for _ in range(n)])


# ---- the actor alone: bit-exact ------------------------------------------------------------------------------------
@pytest.mark.parametrize('hidden,activation,key', [(72, 'tanh', 'serl10_elite_h72_tanh'), (32, 'tanh', 'serl50_pop8_h32_tanh'),
                                                   (96, 'relu', 'td3_h96_relu'), (64, 'elu', None), (128, 'tanh', None),
                                                   (72, 'elu', None), (50, 'tanh', None)])
def test_actor_forward_is_bit_exact_with_the_kernel_order_oracle(hidden, activation, key):
    from serl_b200 import rollout
    dev = torch.device('cuda:0')
    g = (ACT[key] if key else random_genomes(1, hidden, activation, 5)[0]).reshape(-1, rollout.num_params(rollout.actor_shape(hidden, 3, activation)))[0]
    rs = np.random.RandomState(hidden)
    obs = np.concatenate([rs.randn(4000, 7) * [0.05, 0.05, 0.01, 0.02, 0.02, 0.02, 0.05],      # flight-like
                          rs.randn(3000, 7), rs.randn(1000, 7) * 30.0, np.zeros((1, 7))]).astype(np.float32)
    if key is None:
        g = g * np.float32(3.0) if activation == 'tanh' else g         # drive the tanh into saturation too
    got = rollout.actor_forward(torch.as_tensor(g, device=dev), rollout.actor_shape(hidden, 3, activation), torch.as_tensor(obs, device=dev)).cpu().numpy()
    want = fast.actor_forward_kernel_order(g, obs, hidden, 3, activation)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), np.abs(got - want).max()
    assert np.abs(got).max() <= 1.0
    # and the kernel-order actor is the reference actor up to float32 round-off (torch forward, oracle/actor.py)
    ref = A.unflatten(g, hidden=hidden, activation=activation)(torch.as_tensor(obs)).detach().numpy()
    assert np.abs(got - ref).max() < (2e-5 if activation != 'relu' else 2e-4)


def test_activation_functions_match_libm_within_a_few_ulp():
    x = np.concatenate([np.linspace(-12, 12, 200001), 10.0 ** np.linspace(-30, 1, 20001)]).astype(np.float32)
    t = np.tanh(x.astype(np.float64))
    y = fast.tanh_kernel_order(x).astype(np.float64)
    assert (np.abs(y - t) / np.spacing(np.abs(t).astype(np.float32))).max() < 3.0 and np.abs(y).max() <= 1.0
    xn = -np.abs(x)
    e = np.expm1(xn.astype(np.float64))
    assert (np.abs(fast.expm1_neg_kernel_order(xn) - e) / np.spacing(np.abs(e).astype(np.float32))).max() < 2.0


# ---- 1000 trajectories, every fault mode and plant variant (was scripts/parity_sweep.py) ----------------------------
def test_strict_parity_sweep_1000_trajectories():
    w = np.concatenate([ACT['serl10_pop_h72_tanh'], random_genomes(40, 72, 'tanh', 123)]).astype(np.float32)
    n_envs = 20
    modes = [MODES10[i % len(MODES10)] for i in range(n_envs)]
    lv, st = refsig.make_ref_params(n_envs, seed_base=555)
    r = gpu_rollout(w, 72, 'tanh', lv, st, modes)
    oret, ostp = fast.evaluate_population(w, 72, lv, st, modes, actor_order='kernel')
    assert (ostp < 2001).sum() > 300 and (ostp == 2001).sum() > 150          # both regimes are in the sweep
    worst = strict_check(r, oret, ostp)
    # the same launch with the envs grouped by mode (env_order): identical bits
    r2 = gpu_rollout(w, 72, 'tanh', lv, st, modes, sort=True)
    assert torch.equal(r.returns, r2.returns) and torch.equal(r.steps, r2.steps)
    # reported, not asserted: how far the reference's own summation order moves the same trajectories
    iret, istp = fast.evaluate_population(w, 72, lv, st, modes, actor_order='index')
    same = istp == ostp
    print('strict sweep: max rel err %.2e; reference-order self-sensitivity: %d/%d termination steps differ, max rel %.2e'
          % (worst, (~same).sum(), same.size, (np.abs(iret - oret) / np.abs(oret))[same].max()))


# ---- BASELINE config 3 at full size: 1 % of the 65,536 (actor, env) pairs against the oracle ------------------------
def test_config3_full_size_sample_against_oracle():
    rs = np.random.RandomState(11)
    base = ACT['serl10_pop_h72_tanh']
    w = (np.tile(base, (52, 1))[:512] + rs.randn(512, base.shape[1]).astype(np.float32) * np.float32(1e-3)).astype(np.float32)
    lv, st = refsig.make_ref_params(128)
    modes = ['nominal'] * 128
    r = gpu_rollout(w, 72, 'tanh', lv, st, modes)
    actors = rs.choice(512, 41, replace=False)
    envs = rs.choice(128, 16, replace=False)
    oret, ostp = fast.evaluate_population(w[actors], 72, lv[envs], st[envs], ['nominal'] * 16, actor_order='kernel')
    sub = type('R', (), {})()
    sub.returns, sub.steps = r.returns[actors][:, envs], r.steps[actors][:, envs]
    strict_check(sub, oret, ostp)
    assert int(r.steps.sum()) > 0.9 * 512 * 128 * 2001          # the trained population flies (nearly) full episodes


def test_config4_mixed_faults_sample_against_oracle():
    rs = np.random.RandomState(12)
    base = ACT['serl10_pop_h72_tanh']
    w = (np.tile(base, (16, 1))[:150] + rs.randn(150, base.shape[1]).astype(np.float32) * np.float32(1e-3)).astype(np.float32)
    cfg4 = ['nominal', 'be', 'jr', 'sa', 'se', 'ice', 'cg']
    modes = [cfg4[i] for i in rs.randint(0, 7, size=128)]
    lv, st = refsig.make_ref_params(128, seed_base=4242)
    r = gpu_rollout(w, 72, 'tanh', lv, st, modes, sort=True)
    actors = rs.choice(150, 12, replace=False)
    envs = rs.choice(128, 28, replace=False)
    oret, ostp = fast.evaluate_population(w[actors], 72, lv[envs], st[envs], [modes[e] for e in envs], actor_order='kernel')
    sub = type('R', (), {})()
    sub.returns, sub.steps = r.returns[actors][:, envs], r.steps[actors][:, envs]
    strict_check(sub, oret, ostp)


@pytest.mark.parametrize('hidden,activation,n', [(72, 'elu', 6), (64, 'elu', 4), (96, 'relu', 3), (32, 'tanh', 8), (128, 'tanh', 3),
                                                 (50, 'tanh', 3)])
def test_strict_parity_other_shapes_and_activations(hidden, activation, n):
    """ELU / LeakyReLU / other widths (h = 50 runs the per-thread cross-check kernel, same arithmetic specification)."""
    w = random_genomes(n, hidden, activation, 77 + hidden)
    if hidden == 32 and activation == 'tanh':
        w = ACT['serl50_pop8_h32_tanh']
    if hidden == 96:
        w = np.concatenate([ACT['td3_h96_relu'][None], w])
    modes = ['nominal', 'ice', 'be', 'cg', 'sa']
    lv, st = refsig.make_ref_params(len(modes), seed_base=900 + hidden)
    r = gpu_rollout(w, hidden, activation, lv, st, modes)
    oret, ostp = fast.evaluate_population(w, hidden, lv, st, modes, activation=activation, actor_order='kernel')
    strict_check(r, oret, ostp)


def test_evaluation_mode_80s_strict():
    w = ACT['serl10_pop_h72_tanh'][:3]
    lv, st = refsig.make_ref_params(4, seed_base=80, t_max=80)
    modes = ['nominal', 'ice', 'cg', 'be']
    r = gpu_rollout(w, 72, 'tanh', lv, st, modes, horizon=8001, t_max=80.0, smooth_width=13.0)
    oret, ostp = fast.evaluate_population(w, 72, lv, st, modes, t_max=80.0, smooth_w=13.0, horizon=8001, actor_order='kernel')
    strict_check(r, oret, ostp)


# ---- the time-split persistent schedule is invisible in the results ---------------------------------------------------
def test_time_split_schedule_reproduces_whole_episodes_bitwise():
    """more tasks than CTA slots -> slots fly head / tail segments of episodes and hand trajectories over through HBM;
    the same genome must get the same bits wherever and however its episodes were scheduled."""
    from serl_b200 import rollout
    g = ACT['serl10_pop_h72_tanh'][:5]
    lv, st = refsig.make_ref_params(128, seed_base=31)
    modes = [['nominal', 'ice', 'be', 'jr'][i % 4] for i in range(128)]
    single = gpu_rollout(g, 72, 'tanh', lv, st, modes, horizon=300)           # 5 tasks: every task flown whole
    for pop in (311, 450, 700):
        w = np.tile(g, (pop // 5 + 1, 1))[:pop]
        r = gpu_rollout(w, 72, 'tanh', lv, st, modes, horizon=300)
        for a in range(pop):
            assert torch.equal(r.returns[a], single.returns[a % 5]) and torch.equal(r.steps[a], single.steps[a % 5]), (pop, a)
    # ragged env count + early terminations across a hand-over
    torch.manual_seed(3)
    wr = random_genomes(9, 72, 'tanh', 3)
    lv2, st2 = refsig.make_ref_params(70, seed_base=77)
    one = gpu_rollout(wr, 72, 'tanh', lv2, st2, ['nominal'] * 70)
    many = gpu_rollout(np.tile(wr, (40, 1)), 72, 'tanh', lv2, st2, ['nominal'] * 70)
    assert (one.steps.cpu().numpy() < 2001).any()
    for a in range(360):
        assert torch.equal(many.returns[a], one.returns[a % 9]) and torch.equal(many.steps[a], one.steps[a % 9])


def test_replay_export_equals_trace_rebuilt_transitions():
    """N2: the transitions K1 writes for the stored env (agent.py:101-112) equal the ones rebuilt from the per-step trace."""
    w = np.concatenate([ACT['serl10_pop_h72_tanh'][:2], random_genomes(2, 72, 'tanh', 9)])
    lv, st = refsig.make_ref_params(3, seed_base=17)
    modes = ['nominal', 'ice', 'nominal']
    r = gpu_rollout(w, 72, 'tanh', lv, st, modes, trace=True, replay_env=2)
    rp = r.replay.cpu().numpy()
    tr = r.trace.cpu().numpy()
    stp = r.steps.cpu().numpy()
    from oracle import plant as OP
    x_ic = np.asarray(OP.make_plant('h2000_v90', 'port').initial_state())     # reset(): obs = [0, 0, 0, IC[p, q, r, alpha]]
    for a in range(4):
        n = stp[a, 2]
        t = tr[a, 2, :n]
        next_obs = np.hstack((t[:, 19:22], t[:, [0, 1, 2, 4]])).astype(np.float32)
        obs = np.vstack((np.hstack((np.zeros(3), x_ic[[0, 1, 2, 4]]))[None].astype(np.float32), next_obs[:-1]))
        assert np.array_equal(rp[a, :n, 0:7], obs)
        assert np.array_equal(rp[a, :n, 7:10], t[:, 16:19].astype(np.float32))
        assert np.array_equal(rp[a, :n, 10:17], next_obs)
        assert np.array_equal(rp[a, :n, 17], t[:, 15].astype(np.float32))
        done = np.zeros(n, dtype=np.float32)
        done[-1] = 1.0
        assert np.array_equal(rp[a, :n, 18], done)
        cost = (np.rad2deg(np.abs(t[:, 4])) > 11.0) | (np.rad2deg(np.abs(t[:, 6])) > 0.75 * np.deg2rad(75.0)) | (t[:, 3] < 90.0 / 3)
        assert np.array_equal(rp[a, :n, 19], cost.astype(np.float32))


def test_status_flag_reports_non_finite_trajectories():
    from serl_b200 import rollout, _native
    w = ACT['serl10_pop_h72_tanh'][:2].copy()
    w[1, 100] = np.nan
    lv, st = refsig.make_ref_params(2, seed_base=1)
    dev = torch.device('cuda:0')
    md = torch.tensor([0, 0], dtype=torch.int32, device=dev)
    r = rollout.population_rollout(torch.as_tensor(w, device=dev), rollout.actor_shape(72), torch.as_tensor(lv, device=dev),
                                   torch.as_tensor(st, device=dev), md, horizon=50)
    with pytest.raises(_native.NativeError):
        r.check()
    good = rollout.population_rollout(torch.as_tensor(w[:1], device=dev), rollout.actor_shape(72), torch.as_tensor(lv, device=dev),
                                      torch.as_tensor(st, device=dev), md, horizon=50)
    good.check()
