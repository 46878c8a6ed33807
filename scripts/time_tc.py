import os, sys, json, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from serl_b200 import rollout, refsig
from oracle import actor as A
dev = torch.device('cuda:0')
out = {}
for widths in ([400, 300], [128, 128]):
    torch.manual_seed(7)
    base = []
    for _ in range(8):
        m = A.WideActor(widths)
        with torch.no_grad():
            m.net[-2].weight.mul_(0.2); m.net[-2].bias.mul_(0.2)
        base.append(A.flatten(m))
    w = torch.from_numpy(np.tile(np.stack(base), (16, 1)).astype(np.float32)).to(dev)      # 128 actors
    lv, st = refsig.make_ref_params(256)
    md = torch.zeros(256, dtype=torch.int32, device=dev)
    for i in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = rollout.population_rollout(w, rollout.actor_shape(72), torch.as_tensor(lv, device=dev), torch.as_tensor(st, device=dev), md, horizon=600, widths=widths)
        e1.record(); torch.cuda.synchronize()
    out[str(widths)] = [round(e0.elapsed_time(e1), 2), float(r.returns.sum())]
print(os.environ.get('SERL_B200_LIB', 'default').split('_')[-1], json.dumps(out))
