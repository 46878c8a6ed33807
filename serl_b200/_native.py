"""ctypes binding of the C-ABI library (include/serl_b200.h).  The product path has no CPU fallback:
importing succeeds without a GPU (so host logic is testable), but the library must exist and every
compute call fails loudly when CUDA is unavailable."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SERL_B200_LIB') or os.path.join(HERE, 'libserl_b200.so')   # SERL_B200_LIB: e.g. the --exact validation build


class ActorShape(ctypes.Structure):
    _fields_ = [('state_dim', ctypes.c_int32), ('action_dim', ctypes.c_int32), ('hidden', ctypes.c_int32),
                ('num_layers', ctypes.c_int32), ('activation', ctypes.c_int32)]


class RolloutDesc(ctypes.Structure):
    """serl_rollout_desc (include/serl_b200.h)"""
    _fields_ = [('d_weights', ctypes.c_void_p), ('pop', ctypes.c_int32), ('shape', ActorShape),
                ('d_ref_levels', ctypes.c_void_p), ('d_ref_starts', ctypes.c_void_p), ('d_env_mode', ctypes.c_void_p),
                ('n_envs', ctypes.c_int32), ('horizon', ctypes.c_int32), ('d_action_noise', ctypes.c_void_p),
                ('d_returns', ctypes.c_void_p), ('d_steps', ctypes.c_void_p), ('d_fitness', ctypes.c_void_p),
                ('d_trace', ctypes.c_void_p), ('d_actions', ctypes.c_void_p),
                ('t_max', ctypes.c_double), ('smooth_width', ctypes.c_double),
                ('d_env_order', ctypes.c_void_p), ('d_replay', ctypes.c_void_p), ('replay_env', ctypes.c_int32),
                ('d_status', ctypes.c_void_p), ('sm_limit', ctypes.c_int32),
                ('widths', ctypes.c_void_p), ('n_widths', ctypes.c_int32), ('d_sensor_noise', ctypes.c_void_p),
                ('flags', ctypes.c_int32)]


REPLAY_COLS = 20
STATUS_NONFINITE = 1
STATUS_GUST_FLAG = 2
ROLLOUT_GUST = 1
ACTIVATIONS = {'tanh': 0, 'elu': 1, 'relu': 2}
_lib = None


class NativeError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError('serl_b200: %s is missing — build it with `python -m serl_b200.build` '
                              '(there is no CPU fallback)' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
        L.serl_actor_num_params.restype = i64
        L.serl_actor_num_params.argtypes = [ctypes.POINTER(ActorShape)]
        L.serl_rollout.restype = ctypes.c_int
        L.serl_rollout.argtypes = [vp, i32, ctypes.POINTER(ActorShape), vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp]
        L.serl_rollout_eval.restype = ctypes.c_int
        L.serl_rollout_eval.argtypes = L.serl_rollout.argtypes[:-1] + [ctypes.c_double, ctypes.c_double, vp]
        L.serl_rollout_run.restype = ctypes.c_int
        L.serl_rollout_run.argtypes = [ctypes.POINTER(RolloutDesc), vp]
        L.serl_actor_forward.restype = ctypes.c_int
        L.serl_actor_forward.argtypes = [vp, ctypes.POINTER(ActorShape), vp, i32, vp, vp]
        L.serl_actor_num_params_wide.restype = i64
        L.serl_actor_num_params_wide.argtypes = [vp, i32]
        L.serl_actor_forward_wide.restype = ctypes.c_int
        L.serl_actor_forward_wide.argtypes = [vp, vp, i32, i32, vp, i32, vp, vp]
        L.serl_plant_step_timed.restype = ctypes.c_int
        L.serl_plant_step_timed.argtypes = [vp, vp, vp, vp, i32, vp]
        L.serl_smoothness.restype = ctypes.c_int
        L.serl_smoothness.argtypes = [vp, vp, i32, i32, ctypes.c_double, vp, vp]
        L.serl_launch_count.restype = i64
        L.serl_last_error.restype = ctypes.c_char_p
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        raise NativeError('%s failed (%d): %s' % (what, rc, lib().serl_last_error().decode()))
