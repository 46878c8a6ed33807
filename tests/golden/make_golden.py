"""Generate the committed golden fixtures from the reference tree (runs only where /root/reference exists).

  plant_rhs_kat.npz   per plant variant: random states/commands and the right-hand side the reference binary
                      reports for them (rtM_.odeF[0] after step(); SURVEY.md 2.3 "Solver")
  plant_traj_kat.npz  logged reference episodes (logs/wandb/*/files/*statehistory*.txt): commanded actions,
                      12-state trajectory, reward column
  actors.npz          flattened genomes of shipped checkpoints (SERL10 elite h=72 tanh; SERL50 first 8 actors
                      h=32; TD3 actor h=96 relu) + the first rows of TD3 episode 575 (obs -> action KAT)
"""
import ctypes, glob, os, subprocess, sys, shutil, tempfile
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = '/root/reference'
D = ctypes.c_double
VARIANTS = ['h2000_v90', 'ice', 'cg', 'cg_for', 'h2000_v150', 'h10000_v90']


def load(variant):
    src = '%s/envs/%s/_citation.cpython-38-x86_64-linux-gnu.so' % (REF, variant)
    fd, path = tempfile.mkstemp(suffix='.so'); os.close(fd); shutil.copy(src, path)
    lib = ctypes.CDLL(path)
    lib.step.argtypes = [ctypes.POINTER(D)] * 2
    nm = subprocess.run(['nm', path], capture_output=True, text=True).stdout
    sym = {l.split()[2]: int(l.split()[0], 16) for l in nm.splitlines() if len(l.split()) == 3}
    rtX = (D * 19).in_dll(lib, 'rtX')
    odeF = (D * 114).from_address(ctypes.addressof(rtX) - sym['rtX'] + sym['rtM_'] + 46592)
    return lib, rtX, odeF


def rhs_kat(n=200):
    out = {}
    rng = np.random.RandomState(20220924)
    for v in VARIANTS:
        lib, rtX, odeF = load(v)
        lib.initialize()
        ic = np.array(rtX[:])
        Xs, Us, Fs = [], [], []
        cmd, o = (D * 10)(), (D * 12)()
        while len(Xs) < n:
            sc = rng.choice([0.0, 0.01, 0.1, 0.5, 1.0])
            X = ic.copy()
            span = np.array([1, 1, 1, 40, .3, .3, 1.2, 1.0, 3, 1500])
            X[:10] += sc * rng.uniform(-1, 1, 10) * span
            X[9] = max(X[9], 60.0)
            X[12] += sc * rng.uniform(-1, 1)
            X[15:19] += sc * rng.uniform(-10, 30, 4)
            U = sc * rng.uniform(-.3, .3, 3)
            lib.initialize()
            rtX[:] = list(X)
            cmd[:] = list(U) + [0.0] * 7
            lib.step(cmd, o)
            f = np.array(odeF[0:19])
            if np.all(np.isfinite(f)):
                Xs.append(X); Us.append(U); Fs.append(f)
        out[v + '_X'] = np.array(Xs); out[v + '_U'] = np.array(Us); out[v + '_F'] = np.array(Fs); out[v + '_ic'] = ic
    np.savez_compressed(os.path.join(HERE, 'plant_rhs_kat.npz'), **out)


def traj_kat():
    out = {}
    files = sorted(glob.glob(REF + '/logs/wandb/run-*/files/*statehistory*.txt'))
    keep = [f for f in files if 'episode209' in f or 'rl_statehistory_episode575' in f or 'statehistory_episode612' in f]
    for f in keep:
        a = np.loadtxt(f)
        key = os.path.basename(os.path.dirname(os.path.dirname(f)))[-5:] + '_' + os.path.basename(f)[:-4]
        out[key] = a.astype(np.float64)
    np.savez_compressed(os.path.join(HERE, 'plant_traj_kat.npz'), **out)
    print('trajectories', {k: v.shape for k, v in out.items()})


def actors():
    from oracle import actor as A
    out = {}
    sd = torch.load(REF + '/logs/wandb/run-20220913_165505-12zowviu_SERL10/files/elite_net.pkl', weights_only=False)
    out['serl10_elite_h72_tanh'] = A.flatten(A.from_state_dict(sd, 'tanh'))
    pop = torch.load(REF + '/logs/wandb/run-20220913_165505-12zowviu_SERL10/files/evo_nets.pkl', weights_only=False)
    out['serl10_pop_h72_tanh'] = np.stack([A.flatten(A.from_state_dict(pop['actor_%d' % i], 'tanh')) for i in range(10)])
    pop = torch.load(REF + '/logs/wandb/run-20220924_144643-1xzaqiba_SERL50/files/evo_nets.pkl', weights_only=False)
    out['serl50_pop8_h32_tanh'] = np.stack([A.flatten(A.from_state_dict(pop['actor_%d' % i], 'tanh')) for i in range(8)])
    sd = torch.load(REF + '/logs/wandb/run-20221102_144601-1dixcrrl_TD3/files/rl_net.pkl', weights_only=False)
    out['td3_h96_relu'] = A.flatten(A.from_state_dict(sd, 'relu'))
    np.savez_compressed(os.path.join(HERE, 'actors.npz'), **out)
    print('actors', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    rhs_kat(); traj_kat(); actors()
