"""Golden vectors for the reference-signal shape: columns 0-2 (theta_ref, phi_ref, beta_ref in degrees) of the
15 episodes the reference logged under logs/wandb/*/files/*statehistory*.txt.  Episode.get_history
(base/core/utils.py:24-36) samples the episode's reference callables on tt = linspace(0, length, n_rows), length =
info['t'] (n_rows * 0.01 accumulated).  Run in the build container only (needs /root/reference); output:
tests/golden/refsig_logged.npz  {names, <name>: float64 [n_rows, 3]}."""
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == '__main__':
    out = {}
    for f in sorted(glob.glob('/root/reference/logs/wandb/*/files/*statehistory*.txt')):
        run = f.split('/')[-3].split('_')[-1]
        out[run + '_' + os.path.basename(f)[:-4]] = np.loadtxt(f)[:, 0:3]
    np.savez_compressed(os.path.join(HERE, 'refsig_logged.npz'), **out)
    print(len(out), 'episodes', sum(v.nbytes for v in out.values()), 'bytes raw')
