"""The C-ABI library loads and exports every symbol include/serl_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, 'include', 'serl_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(serl_[a-z0-9_]+)\s*\(', txt)))


def test_library_exports_all_declared_symbols():
    from serl_b200 import build, _native
    build.build()
    lib = ctypes.CDLL(_native.LIB_PATH)
    syms = declared_symbols()
    assert 'serl_rollout' in syms
    for s in syms:
        assert hasattr(lib, s), s


def test_num_params_matches_reference_formula():
    from serl_b200 import rollout
    for h, p in ((32, 3715), (72, 16995), (96, 29571), (128, 51715)):   # P = 3h^2 + 20h + 3
        assert rollout.num_params(rollout.actor_shape(h)) == p


def test_compute_fails_loudly_without_cuda():
    import torch
    from serl_b200 import rollout, _native
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    sh = rollout.actor_shape(32)
    w = torch.zeros((1, rollout.num_params(sh)))
    with pytest.raises(_native.NativeError):
        rollout.population_rollout(w, sh, torch.zeros((1, 2, 6), dtype=torch.float64), torch.zeros((1, 2, 6), dtype=torch.float64),
                                   torch.zeros(1, dtype=torch.int32))
