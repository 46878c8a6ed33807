"""C whole-episode port of the hot path (oracle/plant/episode.c), OpenMP over trajectories — checker / CPU-baseline tool.
The pinned oracle remains oracle/phlab.py + oracle/actor.py (the reference's execution model); this port agrees with it to
fp32 round-off of the actor's forward pass and is ~4-5x faster per core because it has no Python / torch dispatch."""
import ctypes
import os

import numpy as np

from . import build as _build, plant as P

_ACT = {'tanh': 0, 'elu': 1, 'relu': 2}


def mode_code(mode):
    v, f = P.MODES[mode]
    return P.VARIANTS.index(v) | (P.FAULTS.index(f) << 8)


def evaluate_population(genomes, hidden, levels, starts, modes, num_layers=3, activation='tanh', t_max=20.0, smooth_w=3.0,
                        horizon=2001, threads=None, actor_order='index'):
    """genomes [pop,P] f32; levels/starts [n_envs,2,6]; modes list of env mode strings -> (returns [pop,n_envs], steps).
    actor_order: 'index' (plain index-order sums + libm tanh: a stand-in for the reference's torch forward pass) or
    'kernel' (the device kernel's summation order and activation arithmetic, oracle/plant/actor_kernel_order.c)."""
    if threads:
        os.environ['OMP_NUM_THREADS'] = str(int(threads))
    lib = ctypes.CDLL(_build.build())
    lib.oracle_population.restype = ctypes.c_long
    g = np.ascontiguousarray(genomes, dtype=np.float32)
    lv = np.ascontiguousarray(levels, dtype=np.float64)
    st = np.ascontiguousarray(starts, dtype=np.float64)
    md = np.asarray([mode_code(m) for m in modes], dtype=np.int32)
    pop, n_envs = g.shape[0], md.shape[0]
    ret = np.zeros((pop, n_envs))
    stp = np.zeros((pop, n_envs), dtype=np.int32)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.oracle_population(vp(g), pop, g.shape[1], 7, 3, int(hidden), int(num_layers), _ACT[activation], vp(md), vp(lv), vp(st), n_envs,
                          ctypes.c_double(t_max), ctypes.c_double(smooth_w), int(horizon), vp(ret), vp(stp),
                          {'index': 0, 'kernel': 1}[actor_order])
    return ret, stp


def actor_forward_kernel_order(genome, obs, hidden, num_layers=3, activation='tanh'):
    """the kernel-order actor on a batch of observations [n,7] -> actions [n,3] (float32, bit-exact with the GPU)."""
    lib = ctypes.CDLL(_build.build())
    g = np.ascontiguousarray(genome, dtype=np.float32)
    o = np.ascontiguousarray(obs, dtype=np.float32)
    out = np.empty((o.shape[0], 3), dtype=np.float32)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.ko_actor_forward_batch(vp(g), 7, 3, int(hidden), int(num_layers), _ACT[activation], vp(o), o.shape[0], vp(out))
    return out


def tanh_kernel_order(x):
    lib = ctypes.CDLL(_build.build())
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    lib.ko_tanhf_batch(x.ctypes.data_as(ctypes.c_void_p), x.size, y.ctypes.data_as(ctypes.c_void_p))
    return y


def expm1_neg_kernel_order(x):
    lib = ctypes.CDLL(_build.build())
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    lib.ko_expm1f_neg_batch(x.ctypes.data_as(ctypes.c_void_p), x.size, y.ctypes.data_as(ctypes.c_void_p))
    return y


def evaluate_population_wide(genomes, widths, levels, starts, modes, activation='tanh', t_max=20.0, smooth_w=3.0, horizon=2001, threads=None):
    """width-list actors ([w1, w2, ...]) through the C episode port (index-order float32 forward pass)."""
    if threads:
        os.environ['OMP_NUM_THREADS'] = str(int(threads))
    lib = ctypes.CDLL(_build.build())
    lib.oracle_population_wide.restype = ctypes.c_long
    g = np.ascontiguousarray(genomes, dtype=np.float32)
    lv = np.ascontiguousarray(levels, dtype=np.float64)
    st = np.ascontiguousarray(starts, dtype=np.float64)
    md = np.asarray([mode_code(m) for m in modes], dtype=np.int32)
    w = np.asarray(widths, dtype=np.int32)
    pop, n_envs = g.shape[0], md.shape[0]
    ret = np.zeros((pop, n_envs))
    stp = np.zeros((pop, n_envs), dtype=np.int32)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.oracle_population_wide(vp(g), pop, g.shape[1], vp(w), len(w), _ACT[activation], vp(md), vp(lv), vp(st), n_envs,
                               ctypes.c_double(t_max), ctypes.c_double(smooth_w), int(horizon), vp(ret), vp(stp))
    return ret, stp
