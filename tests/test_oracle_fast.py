"""The C whole-episode port (oracle/fast.py) against the pinned Python oracle (reference execution model)."""
import os

import numpy as np
import torch

from oracle import actor as A, fast, phlab, refsig

ACT = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'actors.npz'))


def test_c_episode_port_matches_python_oracle():
    torch.manual_seed(7)
    w = np.concatenate([ACT['serl10_pop_h72_tanh'][:2], np.stack([A.flatten(A.Actor(hidden=72)) for _ in range(2)])])
    modes = ['nominal', 'ice', 'se']
    lv, st = refsig.make_ref_params(3, seed_base=21)
    ret, stp = fast.evaluate_population(w, 72, lv, st, modes, threads=4)
    envs = {m: phlab.CitationEnv(m, 'port') for m in modes}
    early = 0
    for a in range(4):
        act = A.unflatten(w[a], hidden=72)
        for e, m in enumerate(modes):
            o = phlab.run_episode(envs[m], act, lv[e], st[e])
            assert stp[a, e] == o['steps']
            assert abs(ret[a, e] - o['fitness']) <= 1e-5 * abs(o['fitness'])
            early += o['steps'] < 2001
    assert early > 0
