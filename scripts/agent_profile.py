import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
pop = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ms, stats, ag = bench.agent_train_timing(torch.device('cuda:0'), pop, 128, generations=4)
print(json.dumps({'generation_ms': ms, 'phases': ag.last_timing, 'spec': [ag.spec_hits, ag.spec_tries], 'test_score': stats['test_score'], 'avg_ep_len': stats['avg_ep_len']}))
