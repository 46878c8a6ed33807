"""Host side of K1: population rollout on one GPU (torch tensors in, torch tensors out).

Mirrors the population-evaluation loop of base/core/agent.py:229-245: every actor of the population is flown
through every environment; fitness[a] = mean over envs of the episodic return.
"""
import ctypes

import torch

from . import _native

HORIZON = 2001          # envs/phlabenv.py:82,181,392: t_max = 20 s, dt = 0.01, done checked before t += dt
PLANT_VARIANTS = ['h2000_v90', 'ice', 'cg', 'cg_for', 'h2000_v150', 'h10000_v90']
FAULTS = ['none', 'be', 'jr', 'sa', 'se']
# env mode string (envs/phlabenv.py:99-172) -> (plant variant, command fault)
MODES = {
    'nominal': ('h2000_v90', 'none'), 'be': ('h2000_v90', 'be'), 'jr': ('h2000_v90', 'jr'),
    'sa': ('h2000_v90', 'sa'), 'se': ('h2000_v90', 'se'), 'ice': ('ice', 'none'), 'cg': ('cg', 'none'),
    'cg-for': ('cg_for', 'none'), 'h2000-v150': ('h2000_v150', 'none'), 'h10000-v90': ('h10000_v90', 'none'),
}


def mode_code(mode):
    v, f = MODES[mode]
    return PLANT_VARIANTS.index(v) | (FAULTS.index(f) << 8)


def actor_shape(hidden, num_layers=3, activation='tanh', state_dim=7, action_dim=3):
    return _native.ActorShape(state_dim, action_dim, hidden, num_layers, _native.ACTIVATIONS[activation.lower()])


def num_params(shape):
    return int(_native.lib().serl_actor_num_params(ctypes.byref(shape)))


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class RolloutResult:
    __slots__ = ('returns', 'steps', 'fitness', 'trace', 'actions', 'smoothness')

    # views into the trace record (include/serl_b200.h: SERL_TRACE_COLS)
    trace_x = property(lambda s: s.trace[..., 0:12])
    trace_u = property(lambda s: s.trace[..., 12:15])
    trace_r = property(lambda s: s.trace[..., 15])
    trace_a = property(lambda s: s.trace[..., 16:19])
    trace_err = property(lambda s: s.trace[..., 19:22])


TRACE_COLS = 22


def population_rollout(weights, shape, ref_levels, ref_starts, env_mode, horizon=HORIZON, trace=False, out=None, action_noise=None,
                       actions=False, t_max=None, smooth_width=None):
    """weights [pop,P] fp32 cuda; ref_levels/ref_starts [n_envs,2,6] f64 cuda; env_mode [n_envs] int32 cuda."""
    if not weights.is_cuda:
        raise _native.NativeError('population_rollout needs CUDA tensors (no CPU fallback)')
    L = _native.lib()
    pop, P = weights.shape
    assert weights.dtype == torch.float32 and weights.is_contiguous()
    assert P == num_params(shape), (P, num_params(shape))
    n_envs = env_mode.shape[0]
    assert ref_levels.shape == (n_envs, 2, 6) and ref_levels.dtype == torch.float64 and ref_levels.is_contiguous()
    assert ref_starts.shape == (n_envs, 2, 6) and ref_starts.dtype == torch.float64 and ref_starts.is_contiguous()
    assert env_mode.dtype == torch.int32
    if action_noise is not None:
        assert action_noise.shape == (pop, n_envs, horizon, 3) and action_noise.dtype == torch.float32 and action_noise.is_contiguous()
    dev = weights.device
    r = out if out is not None else RolloutResult()
    if out is None:
        r.returns = torch.empty((pop, n_envs), dtype=torch.float64, device=dev)
        r.steps = torch.empty((pop, n_envs), dtype=torch.int32, device=dev)
        r.fitness = torch.empty((pop,), dtype=torch.float64, device=dev)
        r.trace = None
        r.actions = torch.empty((pop, n_envs, horizon, 3), dtype=torch.float32, device=dev) if actions else None
        if trace:
            r.trace = torch.full((pop, n_envs, horizon, TRACE_COLS), float('nan'), dtype=torch.float64, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    args = (_ptr(weights), pop, ctypes.byref(shape), _ptr(ref_levels), _ptr(ref_starts), _ptr(env_mode),
            n_envs, horizon, _ptr(action_noise), _ptr(r.returns), _ptr(r.steps), _ptr(r.fitness),
            _ptr(r.trace), _ptr(getattr(r, 'actions', None)))
    if t_max is None:
        rc = L.serl_rollout(*args, stream)
    else:      # evaluation mode (envs/phlabenv.py:295-301): longer episodes, wider reference transitions
        rc = L.serl_rollout_eval(*args, ctypes.c_double(t_max), ctypes.c_double(smooth_width if smooth_width is not None else float(t_max // 6)), stream)
    _native.check(rc, 'serl_rollout')
    return r


def smoothness(actions, steps, dt=0.01):
    """K6: per-trajectory action smoothness (core/utils.py calc_smoothness) of `actions` [..., horizon, 3] fp32 (cuda) over the
    first `steps` [...] executed steps. Returns f64 tensor shaped like `steps`."""
    L = _native.lib()
    horizon = actions.shape[-2]
    n = steps.numel()
    assert actions.is_cuda and actions.dtype == torch.float32 and actions.is_contiguous() and actions.numel() == n * horizon * 3
    st = steps.contiguous().to(torch.int32)
    out = torch.empty(st.shape, dtype=torch.float64, device=actions.device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(actions.device).cuda_stream)
    _native.check(L.serl_smoothness(_ptr(actions), _ptr(st), n, horizon, dt, _ptr(out), stream), 'serl_smoothness')
    return out
