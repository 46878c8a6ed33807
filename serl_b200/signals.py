"""The few classes of the third-party `signals` package (signals==0.0.1, absent here) that the reference's evaluation code
constructs (base/evaluate.py:169-180, base/evaluation_utils.py:40-56, envs/phlabenv.py:305-344), restated on the shape
recovered from the logged episodes (tests/test_refsig_pin.py): a sequence of set-points joined by raised-cosine
transitions of width `smooth_width` that START at the set-point times.  An instance is a callable t -> value [deg] and
carries the block parameters (levels, starts) the rollout kernel evaluates on the device."""
import numpy as np

from . import refsig


class SmoothedStepSequence:
    def __init__(self, times, amplitudes, smooth_width=3.0):
        n = refsig.N_BLOCKS
        t, a = list(np.asarray(times, dtype=np.float64)), list(np.asarray(amplitudes, dtype=np.float64))
        if len(t) != len(a) or len(t) > n:
            raise ValueError('SmoothedStepSequence: need equally many times and amplitudes, at most %d' % n)
        while len(t) < n:                      # pad with set-points that are never reached
            t.append(1e30); a.append(a[-1])
        self.starts, self.levels = np.asarray(t), np.asarray(a)
        self.smooth_width = float(smooth_width)
        self.offset, self.t_end = 0.0, None

    def __call__(self, t):
        return refsig.ref_value_deg(self.levels, self.starts, t, self.offset, self.smooth_width, self.t_end)

    def __add__(self, other):
        """step + Const(t0, t1, value): the constant exists on [t0, t1] only (see refsig.ref_value_deg)."""
        if isinstance(other, Const):
            s = SmoothedStepSequence(self.starts, self.levels, self.smooth_width)
            s.offset, s.t_end = self.offset + other.value, other.t1
            return s
        return NotImplemented


class RandomizedCosineStepSequence(SmoothedStepSequence):
    """envs/phlabenv.py:321-335: n blocks of `block_width`, first level 0, levels on linspace(-A, A, 10), start times jittered by
    +-vary_timings; consumes the global np.random stream (the third-party generator's own draw order is unpinned)."""

    def __init__(self, t_max=20, ampl_max=30, block_width=4, smooth_width=3, n_levels=10, vary_timings=0.04):
        n = refsig.N_BLOCKS
        grid = np.linspace(-ampl_max, ampl_max, refsig.N_LEVELS)
        lv = grid[np.random.randint(0, refsig.N_LEVELS, size=n)]
        lv[0] = 0.0
        st = block_width * np.arange(n) + np.random.uniform(-vary_timings, vary_timings, size=n)
        st[0] = 0.0
        super().__init__(st, lv, smooth_width)


class Const:
    def __init__(self, t0, t1, value):
        self.t0, self.t1, self.value = float(t0), float(t1), float(value)

    def __call__(self, t):
        return self.value if self.t0 <= t <= self.t1 else 0.0
