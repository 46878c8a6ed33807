"""Mirror of envs/config.py:9-41 — environment registry (`select_env(name) -> env`)."""
from .phlabenv import CitationEnv


def select_env(environemnt_name: str):
    _name = environemnt_name
    if 'lunar' in _name.lower():
        raise ValueError('LunarLanderContinuous-v2 needs gym + Box2D, which this B200 engine does not ship '
                         '(BASELINE config 1 is a CPU plumbing case; see DESIGN.md)')
    elif 'ph' in _name.lower():
        tokens = _name.lower().split('_')
        phlab_mode = 'nominal'
        if len(tokens) == 3:
            _, phlab_config, phlab_mode = tokens
        else:
            phlab_config = tokens[-1]
            phlab_mode = ''
        return CitationEnv(configuration=phlab_config, mode=phlab_mode)
    else:
        raise ValueError(f'{_name} is an unknown environment type')
