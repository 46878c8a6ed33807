"""Small driver for ncu: launches the rollout kernel a few times on a reduced problem (same per-thread work)."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from serl_b200 import rollout, refsig
import bench

ap = argparse.ArgumentParser()
ap.add_argument('--pop', type=int, default=296)
ap.add_argument('--envs', type=int, default=128)
ap.add_argument('--horizon', type=int, default=200)
ap.add_argument('--iters', type=int, default=3)
ap.add_argument('--mixed', action='store_true')
ap.add_argument('--sorted', action='store_true')
ap.add_argument('--hidden', type=int, default=72)
a = ap.parse_args()
dev = torch.device('cuda:0')
sh = rollout.actor_shape(a.hidden)
if a.hidden == 72:
    w = torch.from_numpy(bench.population(a.pop)).to(dev)
else:
    base = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'actors.npz'))['serl50_pop8_h32_tanh']
    w = torch.from_numpy(np.ascontiguousarray(base[np.arange(a.pop) % 8])).to(dev)
lv, st = refsig.make_ref_params(a.envs)
modes = ['nominal'] * a.envs
if a.mixed:
    modes = [['nominal', 'be', 'jr', 'sa', 'se', 'ice', 'cg'][i % 7] for i in range(a.envs)]
if a.sorted:
    modes = sorted(modes, key=rollout.mode_code)
md = torch.tensor([rollout.mode_code(m) for m in modes], dtype=torch.int32, device=dev)
lv, st = torch.from_numpy(lv).to(dev), torch.from_numpy(st).to(dev)
r = None
for i in range(a.iters):
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    r = rollout.population_rollout(w, sh, lv, st, md, horizon=a.horizon, out=r)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    steps = int(r.steps.sum().item())
    print('iter %d: %.2f ms, %d steps, %.3e env-steps/s' % (i, ms, steps, steps / ms * 1e3))
