"""Environment registry with the reference's entry point `select_env(name)` (envs/config.py:9-41).

Names are '<family>_<configuration>[_<mode>]', e.g. 'PHlab_attitude_nominal', 'phlab_attitude_ice'; anything containing
'ph' is the PH-LAB Citation task, anything containing 'lunar' would be the gym LunarLander plumbing case."""
from .phlabenv import CitationEnv


def select_env(environemnt_name: str):
    name = environemnt_name.lower()
    if 'lunar' in name:
        raise ValueError('LunarLanderContinuous-v2 needs gym + Box2D, which this B200 engine does not ship '
                         '(BASELINE config 1 is a CPU plumbing case; see DESIGN.md)')
    if 'ph' not in name:
        raise ValueError(f'{environemnt_name} is an unknown environment type')
    parts = name.split('_')
    if len(parts) == 3:
        configuration, mode = parts[1], parts[2]
    else:
        configuration, mode = parts[-1], ''
    return CitationEnv(configuration=configuration, mode=mode)
