"""One-off wide parity sweep (bigger than the test-suite cases): random-init + trained actors x envs x all modes,
CUDA rollout vs the oracle (reference plant binary when oracle/_ref is present).  Prints a JSON summary."""
import json, os, sys, time
import multiprocessing as mp
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
MODES = ['nominal', 'be', 'jr', 'sa', 'se', 'ice', 'cg', 'cg-for', 'h2000-v150', 'h10000-v90']


def worker(args):
    import torch
    torch.set_num_threads(1)
    from oracle import actor as A, phlab
    w, jobs, lv, st, modes = args
    envs = {}
    out = []
    for a, e in jobs:
        m = modes[e]
        if m not in envs:
            envs[m] = phlab.CitationEnv(m, 'auto')
        o = phlab.run_episode(envs[m], A.unflatten(w[a], hidden=72), lv[e], st[e])
        out.append((a, e, o['steps'], o['fitness']))
    return out


if __name__ == '__main__':
    import torch
    from serl_b200 import rollout, refsig
    from oracle import actor as A
    n_rand, n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 40, 20
    acts = np.load(os.path.join(ROOT, 'tests', 'golden', 'actors.npz'))
    torch.manual_seed(123)
    w = np.concatenate([acts['serl10_pop_h72_tanh'], np.stack([A.flatten(A.Actor(hidden=72)) for _ in range(n_rand)])]).astype(np.float32)
    modes = [MODES[i % len(MODES)] for i in range(n_envs)]
    lv, st = refsig.make_ref_params(n_envs, seed_base=555)
    dev = torch.device('cuda:0')
    md = torch.tensor([rollout.mode_code(m) for m in modes], dtype=torch.int32, device=dev)
    r = rollout.population_rollout(torch.as_tensor(w, device=dev), rollout.actor_shape(72), torch.as_tensor(lv, device=dev),
                                   torch.as_tensor(st, device=dev), md)
    torch.cuda.synchronize()
    ret, stp = r.returns.cpu().numpy(), r.steps.cpu().numpy()
    jobs = [(a, e) for a in range(w.shape[0]) for e in range(n_envs)]
    import bench
    cores = bench.host_cores()
    chunks = [jobs[i::cores] for i in range(cores)]
    t0 = time.time()
    with mp.get_context('spawn').Pool(cores) as pool:
        res = [x for part in pool.map(worker, [(w, c, lv, st, modes) for c in chunks]) for x in part]
    mism, worst, early = [], 0.0, 0
    for a, e, s, f in res:
        early += s < 2001
        if stp[a, e] != s:
            mism.append((a, e, modes[e], int(stp[a, e]), s))
        else:
            worst = max(worst, abs(ret[a, e] - f) / abs(f))
    print(json.dumps({'trajectories': len(res), 'early_terminations': int(early), 'termination_step_mismatches': mism[:10],
                      'n_mismatch': len(mism), 'max_rel_return_error_when_steps_match': worst, 'oracle_seconds': time.time() - t0,
                      'modes': MODES}))
