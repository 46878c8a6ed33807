import os, sys, json, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
import test_rollout_gpu as T
from oracle import refsig
lv, st = refsig.make_ref_params(2, seed_base=5)
out = {}
for name, h, act, sl in (('serl50_pop8_h32_tanh', 32, 'tanh', slice(0, 3)), ('td3_h96_relu', 96, 'relu', None), ('serl10_pop_h72_tanh', 72, 'tanh', slice(0, 3))):
    w = T.ACT[name]
    w = w[sl] if sl is not None else w[None]
    r = T.gpu_rollout(w, h, act, lv, st, ['nominal', 'nominal'])
    out[name] = [r.returns.cpu().tolist(), r.steps.cpu().tolist()]
print(json.dumps(out))
''' % (ROOT, ROOT)
res = {}
for impl in ('warp', 'simple'):
    p = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=dict(os.environ, SERL_ROLLOUT_IMPL=impl))
    if p.returncode: print(p.stderr[-2000:])
    res[impl] = json.loads(p.stdout.strip().splitlines()[-1])
for k in res['warp']:
    print(k); print('  warp  ', res['warp'][k]); print('  simple', res['simple'][k])
