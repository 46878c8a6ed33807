"""K2-K5 through the C-ABI vs the reference-recorded genomes and vs the oracle: selection / elite indices and
post-epoch genomes bit-exact (BASELINE.json parity bar for the EA)."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import ssne as OS

pytestmark = pytest.mark.gpu
KAT = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ssne_kat.npz'))
CASES = sorted({k.split('_')[0] for k in KAT.files})


@pytest.mark.parametrize('case', CASES)
def test_device_epoch_equals_reference_module(case):
    from serl_b200 import evo
    before, fit, after = KAT[case + '_before'], KAT[case + '_fitness'], KAT[case + '_after']
    seed = int(KAT[case + '_seed'])
    shape = tuple(int(x) for x in KAT[case + '_shape'])
    W = torch.as_tensor(before.copy(), device='cuda:0')
    np.random.seed(seed + 1)
    random.seed(seed + 2)
    elite, plan = evo.epoch_flat(W, fit, shape)
    assert elite == int(KAT[case + '_elite'])
    assert np.array_equal(W.cpu().numpy().view(np.uint32), after.view(np.uint32))


@pytest.mark.parametrize('pop,hidden,ties', [(64, 32, False), (50, 72, True), (200, 16, True), (5, 8, True)])
def test_device_epoch_equals_oracle(pop, hidden, ties):
    from serl_b200 import evo
    shape = (7, 3, hidden, 3)
    rng = np.random.RandomState(pop)
    P = OS.param_table(*shape)[1]
    before = rng.normal(0, 0.2, (pop, P)).astype(np.float32)
    fit = rng.uniform(-3000, -50, pop)
    if ties:
        fit = np.round(fit, -2)
        fit[:2] = np.nan if pop > 50 else fit[:2]
    for gen in range(2):
        np.random.seed(11 + gen); random.seed(13 + gen)
        Wo = before.copy()
        elite_o = OS.SSNE(pop, shape).epoch(Wo, fit)
        np.random.seed(11 + gen); random.seed(13 + gen)
        W = torch.as_tensor(before.copy(), device='cuda:0')
        elite, plan = evo.epoch_flat(W, torch.as_tensor(fit, device='cuda:0'), shape)
        assert elite == elite_o
        assert np.array_equal(W.cpu().numpy().view(np.uint32), Wo.view(np.uint32))
        before = Wo


def test_select_kernel_rank_rule():
    from serl_b200 import evo
    fit = np.array([1.0, 3.0, 3.0, -2.0, np.nan, 0.5, 3.0])
    np.random.seed(0)
    rank, offs = evo.select_device(torch.as_tensor(fit, device='cuda:0'), 1)
    assert list(rank) == list(np.argsort(fit, kind='stable')[::-1])
    np.random.seed(0)
    draws = np.stack([np.random.randint(7, size=3) for _ in range(6)])
    assert list(offs) == list(rank[draws.min(1)])
