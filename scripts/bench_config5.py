"""BASELINE config 5: actor hidden = [400,300] vs [128,128], pop = 1024 x 256 envs (8 GPUs: 128 actors per GPU).
Times the tensor-core kernel (csrc/rollout_tc.cu) for both shapes and the warp-GEMV kernel K1 for [128,128] (which is the
reference's uniform form with hidden = 128, num_layers = 1) -> the tensor-core vs warp-GEMV crossover.  Prints one JSON object.
usage: python scripts/bench_config5.py [pop=128] [n_envs=256] [horizon=2001]"""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from serl_b200 import rollout, refsig
from oracle import actor as A

pop = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n_envs = int(sys.argv[2]) if len(sys.argv) > 2 else 256
horizon = int(sys.argv[3]) if len(sys.argv) > 3 else 2001
dev = torch.device('cuda:0')


def genomes(widths, n, seed=7, out_gain=0.2):
    torch.manual_seed(seed)
    base = []
    for _ in range(min(n, 16)):
        m = A.WideActor(widths)
        with torch.no_grad():
            m.net[-2].weight.mul_(out_gain); m.net[-2].bias.mul_(out_gain)
        base.append(A.flatten(m))
    base = np.stack(base)
    rs = np.random.RandomState(seed)
    w = base[np.arange(n) % base.shape[0]] + rs.normal(0, 1e-3, size=(n, base.shape[1])).astype(np.float32)
    return torch.from_numpy(np.ascontiguousarray(w.astype(np.float32))).to(dev)


lv, st = refsig.make_ref_params(n_envs)
lv, st = torch.as_tensor(lv, device=dev), torch.as_tensor(st, device=dev)
md = torch.zeros(n_envs, dtype=torch.int32, device=dev)
out = {'pop': pop, 'n_envs': n_envs, 'horizon': horizon, 'population': 'random-init wide actors, output gain 0.2, tiled + N(0,1e-3) noise'}


def timed(name, w, shape, widths):
    res = None
    for i in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = rollout.population_rollout(w, shape, lv, st, md, horizon=horizon, widths=widths)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    res.check()
    steps = int(res.steps.sum().item())
    macs = (7 * 128 + 128 * 128 + 128 * 3) if (widths is None or widths == [128, 128]) else (7 * widths[0] + widths[0] * widths[1] + widths[1] * 3)
    out[name] = {'ms': ms, 'executed_env_steps': steps, 'env_steps_per_sec': steps / (ms * 1e-3), 'mean_episode_steps': steps / (pop * n_envs),
                 'actor_gflops_algorithmic': 2 * macs * steps / (ms * 1e-3) / 1e9}
    return res


w128 = genomes([128, 128], pop)
r_tc = timed('tc_128_128', w128, rollout.actor_shape(72), [128, 128])
r_k1 = timed('k1_warp_gemv_128_128', w128, rollout.actor_shape(128, 1, 'tanh'), None)
out['tc_vs_k1_128_128'] = {'same_termination_steps': bool(torch.equal(r_tc.steps, r_k1.steps)),
                           'max_rel_return_diff': float(((r_tc.returns - r_k1.returns).abs() / r_k1.returns.abs()).max().item())}
w400 = genomes([400, 300], pop)
timed('tc_400_300', w400, rollout.actor_shape(72), [400, 300])
print(json.dumps(out))
