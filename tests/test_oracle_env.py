"""Pin the oracle's env wrapper and actor against the reference's logged artefacts."""
import os

import numpy as np

from oracle import actor as A, phlab, plant as P, refsig

G = os.path.join(os.path.dirname(__file__), 'golden')
TRAJ = np.load(os.path.join(G, 'plant_traj_kat.npz'))
ACT = np.load(os.path.join(G, 'actors.npz'))


def test_actor_forward_matches_logged_td3_episode():
    """u_{i+1} = 10deg * actor([deg2rad(ref_i) - (theta,phi,beta)_i, p,q,r,alpha_i])  (SURVEY 4.3)."""
    a = TRAJ['l_TD3_rl_statehistory_episode575']
    act = A.unflatten(ACT['td3_h96_relu'], hidden=96, activation='relu')
    env = phlab.CitationEnv('nominal', 'port')
    worst = []
    for i in range(0, 40):
        ref = np.deg2rad(a[i, 0:3])
        x = a[i, 6:18]
        obs = np.hstack((ref - np.array([x[7], x[6], x[5]]), x[[0, 1, 2, 4]]))
        u = env.scale_action(act.select_action(obs))
        worst.append(np.abs(u - a[i + 1, 3:6]).max())
    # the logged ref columns are re-sampled on linspace(0, 20.01, 2001) (utils.py:30-32): exact only while the
    # reference is flat (first 4 s), so the first rows pin the forward pass to fp32 eps
    assert max(worst[:40]) < 5e-7, worst


def test_early_termination_penalty_and_rule():
    a = TRAJ['ERL10_rl_statehistory_episode209']          # 780 rows: the episode died early
    n = a.shape[0]
    x = a[:, 6:18]
    dead = (np.abs(x[:, 7]) > np.deg2rad(60)) | (np.abs(x[:, 6]) > np.deg2rad(75)) | (x[:, 9] < 50)
    assert dead[-1] and not dead[:-1].any()
    t = 0.0
    for _ in range(n - 1):
        t += 0.01
    penalty = -1 / 0.01 * (20 - t) * 2
    assert abs(a[-1, 18] - penalty) < 1.0 + 1e-9          # base reward is in [-1, 0]
    assert a[-1, 18] <= penalty


def test_full_episode_is_2001_steps_and_reward_bounds():
    env = phlab.CitationEnv('nominal', 'port')
    act = A.unflatten(ACT['serl10_elite_h72_tanh'], hidden=72)
    lv, st = refsig.make_ref_params(1)
    r = phlab.run_episode(env, act, lv[0], st[0])
    assert r['steps'] == 2001 and abs(r['t'] - 20.01) < 1e-9
    assert -400 < r['fitness'] < 0


def test_fault_shims():
    c = P.apply_fault('be', np.array([0.1, 0.2, 0.3] + [0.] * 7))
    assert c[0] == 0.1 * 0.3
    c = P.apply_fault('jr', np.array([0.1, 0.2, 0.3] + [0.] * 7))
    assert c[2] == 15 * 3.14159 / 180
    c = P.apply_fault('sa', np.array([0.1, 0.2, 0.3] + [0.] * 7))
    assert c[1] == np.deg2rad(1)
    c = P.apply_fault('se', np.array([-0.1, 0.2, 0.3] + [0.] * 7))
    assert c[0] == -np.deg2rad(2.5)


def test_numpy_constants_used_by_kernel():
    assert np.deg2rad(1.0) == 0.017453292519943295
    assert np.rad2deg(1.0) == 57.29577951308232
    assert np.deg2rad(60.) == 60.0 * 0.017453292519943295
    assert np.deg2rad(75.) == 75.0 * 0.017453292519943295
    assert np.deg2rad(10) == 10.0 * 0.017453292519943295
    assert np.deg2rad(2.5) == 2.5 * 0.017453292519943295
    assert -1 / 0.01 == -100.0
