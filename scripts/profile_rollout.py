"""Launch the BASELINE config-3 population rollout a few times (for `ncu -k regex:rollout_kernel_persist`).
usage: python scripts/profile_rollout.py [pop] [n_envs] [horizon] [launches] [mode: nominal|mixed|random]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from serl_b200 import rollout, refsig

pop = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n_envs = int(sys.argv[2]) if len(sys.argv) > 2 else 128
horizon = int(sys.argv[3]) if len(sys.argv) > 3 else 2001
n = int(sys.argv[4]) if len(sys.argv) > 4 else 2
mode = sys.argv[5] if len(sys.argv) > 5 else 'nominal'
dev = torch.device('cuda:0')
sh = rollout.actor_shape(72, 3, 'tanh')
w = torch.from_numpy(bench.population(pop)).to(dev)
lv, st = refsig.make_ref_params(n_envs)
if mode == 'mixed':
    rs = np.random.RandomState(7)
    modes = [['nominal', 'be', 'jr', 'sa', 'se', 'ice', 'cg'][i] for i in rs.randint(0, 7, n_envs)]
else:
    modes = ['nominal'] * n_envs
md = torch.tensor([rollout.mode_code(m) for m in modes], dtype=torch.int32, device=dev)
lv, st = torch.as_tensor(lv, device=dev), torch.as_tensor(st, device=dev)
for i in range(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = rollout.population_rollout(w, sh, lv, st, md, horizon=horizon)
    e1.record()
    torch.cuda.synchronize()
    steps = int(r.steps.sum().item())
    print('launch %d: %.2f ms, %d env-steps, %.4g env-steps/s' % (i, e0.elapsed_time(e1), steps, steps / (e0.elapsed_time(e1) * 1e-3)))
