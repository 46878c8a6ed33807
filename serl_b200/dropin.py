"""Expose the mirror modules under the names base/train.py imports (`core`, `envs`, `parameters`) — see INTEGRATION.md §1.

    python -c "import serl_b200.dropin as d; d.install(); import runpy; runpy.run_path('base/train.py', run_name='__main__')" ...
"""
import importlib
import sys


def install():
    import serl_b200.core, serl_b200.envs, serl_b200.parameters
    sys.modules['core'] = serl_b200.core
    sys.modules['envs'] = serl_b200.envs
    sys.modules['parameters'] = serl_b200.parameters
    for m in ('agent', 'genetic_agent', 'mod_neuro_evo', 'mod_utils', 'replay_memory', 'td3', 'utils'):
        sys.modules['core.' + m] = importlib.import_module('serl_b200.core.' + m)
        setattr(serl_b200.core, m, sys.modules['core.' + m])
    for m in ('config', 'phlabenv'):
        sys.modules['envs.' + m] = importlib.import_module('serl_b200.envs.' + m)
        setattr(serl_b200.envs, m, sys.modules['envs.' + m])
