"""Where does the host block while the next generation's front is launched?  (python scripts/front_profile.py)"""
import os, sys, time, types, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from serl_b200.core import agent as agent_mod
from serl_b200 import rollout

acc = {}
def wrap(obj, name, label=None):
    f = getattr(obj, name)
    label = label or name
    def g(*a, **k):
        t = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            acc.setdefault(label, []).append(1e3 * (time.perf_counter() - t))
    setattr(obj, name, g)

wrap(agent_mod, '_to_device')
wrap(rollout, 'population_rollout')
wrap(rollout, 'smoothness')
wrap(agent_mod.Agent, '_fly')
wrap(agent_mod.Agent, '_launch_population')
wrap(agent_mod.Agent, '_launch_front')
wrap(torch, 'full', 'torch.full')
wrap(agent_mod.Agent, 'rl_to_evo')
wrap(agent_mod.Agent, '_collect')
wrap(agent_mod.Agent, '_finish_population')
dev = torch.device('cuda:0')
for g in range(5):
    pass
ms, stats, ag = bench.agent_train_timing(dev, 512, 128, generations=4)
print('generation ms', ms)
print({k: round(v, 1) for k, v in ag.last_timing.items()})
for k, v in acc.items():
    print(f'{k:22s} n={len(v):3d} total={sum(v):8.1f}  last8={[round(x, 1) for x in v[-8:]]}')
