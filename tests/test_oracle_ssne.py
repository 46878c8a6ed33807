"""Pin the oracle's neuro-evolution restatement against before/after genomes recorded from the REFERENCE module
(tests/golden/make_golden_ea.py: base/core/mod_neuro_evo.py + the exclusive-index shim)."""
import os
import random

import numpy as np
import pytest

from oracle import ssne as OS

KAT = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ssne_kat.npz'))
CASES = sorted({k.split('_')[0] for k in KAT.files})


@pytest.mark.parametrize('case', CASES)
def test_epoch_bit_exact_vs_reference_module(case):
    before, fit, after = KAT[case + '_before'], KAT[case + '_fitness'], KAT[case + '_after']
    seed = int(KAT[case + '_seed'])
    shape = tuple(int(x) for x in KAT[case + '_shape'])
    ev = OS.SSNE(before.shape[0], shape)
    assert ev.P == before.shape[1]
    W = before.copy()
    np.random.seed(seed + 1)
    random.seed(seed + 2)
    elite = ev.epoch(W, fit)
    assert elite == int(KAT[case + '_elite'])
    assert np.array_equal(W.view(np.uint32), after.view(np.uint32))      # bit-exact fp32 genomes
