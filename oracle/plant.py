"""Plant back-ends of the oracle (checker only).

RefPlant  : the reference's own native model (oracle/_ref/citation_<variant>.so) driven through ctypes the
            way envs/<variant>/citation.py:65-72 drives it (initialize / step(cmd[10], out[12])).  The model
            state is process-global, but it is a pure function of rtX[19], so any number of
            environments are multiplexed on one handle by saving / restoring the 152 bytes of rtX.
PortPlant : the C restatement oracle/plant/plant_oracle.c (bit-identical to RefPlant, see tests).
"""
import ctypes
import os
import shutil
import tempfile

import numpy as np

from . import build as _build

D = ctypes.c_double
VARIANTS = _build.VARIANTS

# env "mode" (envs/phlabenv.py:99-172) -> (plant variant, command fault)
MODES = {
    'nominal': ('h2000_v90', 'none'), 'be': ('h2000_v90', 'be'), 'jr': ('h2000_v90', 'jr'),
    'sa': ('h2000_v90', 'sa'), 'se': ('h2000_v90', 'se'), 'ice': ('ice', 'none'), 'cg': ('cg', 'none'),
    'cg-for': ('cg_for', 'none'), 'h2000-v150': ('h2000_v150', 'none'), 'h10000-v90': ('h10000_v90', 'none'),
    'cg-timed': ('cg_timed', 'none'),       # time-triggered build: reference binary only, one episode at a time (own clock)
    'gust': ('gust', 'none'),               # likewise (envs/gust: vertical gust for 20 s <= t <= 23 s + the sensor-noise shim)
    'test': ('test', 'none'),               # envs/test: the same pulse with the opposite sign
}
FAULTS = ['none', 'be', 'jr', 'sa', 'se']


def apply_fault(fault, cmd):
    """envs/{be,jr,sa,se}/citation.py:71-79 — in-place transform of the padded command."""
    if fault == 'be':
        cmd[0] *= 0.3
    elif fault == 'jr':
        cmd[2] = 15 * 3.14159 / 180
    elif fault == 'sa':
        b = np.deg2rad(1)
        cmd[1] = np.clip(cmd[1], -b, b)
    elif fault == 'se':
        b = np.deg2rad(2.5)
        cmd[0] = np.clip(cmd[0], -b, b)
    return cmd


class RefPlant:
    """One handle on the reference binary of `variant`. step(X, cmd) -> (out12 = state before the step, X')."""

    def __init__(self, variant):
        src = os.path.join(_build.HERE, '_ref', 'citation_%s.so' % variant)
        if not os.path.exists(src):
            raise FileNotFoundError(src + ' (run oracle/build.py where /root/reference exists)')
        # private copy: dlopen of the same path twice would alias the process-global model state
        fd, self._path = tempfile.mkstemp(suffix='_%s.so' % variant)
        os.close(fd)
        shutil.copy(src, self._path)
        self.lib = ctypes.CDLL(self._path)
        self.lib.step.argtypes = [ctypes.POINTER(D), ctypes.POINTER(D)]
        self.rtX = (D * 19).in_dll(self.lib, 'rtX')
        self.cmd = (D * 10)()
        self.out = (D * 12)()
        self.lib.initialize()
        self.ic = np.array(self.rtX[:], dtype=np.float64)

    def initial_state(self):
        self.lib.initialize()          # also resets the model clock (time-triggered builds)
        return self.ic.copy()

    def step(self, X, cmd10):
        self.rtX[:] = list(X)
        self.cmd[:] = list(cmd10)
        self.lib.step(self.cmd, self.out)
        return np.array(self.out[:]), np.array(self.rtX[:])

    def __del__(self):
        try:
            os.unlink(self._path)
        except Exception:
            pass


class PortPlant:
    def __init__(self, variant):
        self.lib = ctypes.CDLL(_build.build())
        self.v = VARIANTS.index(variant)
        self.lib.plant_step.argtypes = [ctypes.c_int, ctypes.POINTER(D), ctypes.POINTER(D)]
        self.lib.plant_rhs.argtypes = [ctypes.c_int, ctypes.POINTER(D), ctypes.POINTER(D), ctypes.POINTER(D)]
        self.lib.plant_get_ic.argtypes = [ctypes.c_int, ctypes.POINTER(D)]
        ic = (D * 19)()
        self.lib.plant_get_ic(self.v, ic)
        self.ic = np.array(ic[:])

    def initial_state(self):
        return self.ic.copy()

    def rhs(self, X, U):
        xd = (D * 19)()
        self.lib.plant_rhs(self.v, (D * 19)(*X), (D * 3)(*U[:3]), xd)
        return np.array(xd[:])

    def step(self, X, cmd10):
        x = (D * 19)(*X)
        out = np.array(X[:12], dtype=np.float64)
        self.lib.plant_step(self.v, x, (D * 3)(*cmd10[:3]))
        return out, np.array(x[:])


def make_plant(variant, backend='auto'):
    if variant in _build.REF_ONLY:
        backend = 'ref'
    if backend == 'auto':
        backend = 'ref' if _build.have_ref() else 'port'
    return RefPlant(variant) if backend == 'ref' else PortPlant(variant)
