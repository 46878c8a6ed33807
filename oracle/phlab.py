"""Oracle restatement of the PH-LAB gym wrapper (checker only): envs/phlabenv.py:76-486, gym-free.

Differences from the reference, all documented in DESIGN.md:
  * reference signals come from oracle/refsig.py (the `signals` package is absent; parity unpinned);
  * each (actor, env) trajectory is a fresh env: the stale `self.error` in reset() (phlabenv.py:422) is 0.
"""
import numpy as np

from . import plant as P
from . import refsig

DT = 0.01


class CitationEnv:
    n_actions = 3
    obs_idx = [0, 1, 2, 4]

    def __init__(self, mode='nominal', backend='auto', plant=None, t_max=20):
        self.t_max = t_max
        self.smooth_w = refsig.widths(t_max)[1]
        self.variant, self.fault = P.MODES[mode]
        self.plant = plant if plant is not None else P.make_plant(self.variant, backend)
        self.bound = np.deg2rad(10)
        self.max_theta = np.deg2rad(60.)
        self.max_phi = np.deg2rad(75.)
        self.error = np.zeros(3)
        self.error_scaler = 6 / np.pi * np.array([1., 1., 4.])
        self.max_bound = np.ones(3)
        self.low = -self.bound * np.ones(3)
        self.high = self.bound * np.ones(3)
        self.dt = DT

    # phlabenv.py:62-73
    def scale_action(self, a):
        return self.low + 0.5 * (a + 1.0) * (self.high - self.low)

    def _native_step(self, u):
        cmd = np.pad(u, (0, 7), 'constant', constant_values=(0.))
        cmd = P.apply_fault(self.fault, cmd)
        out, self.X = self.plant.step(self.X, cmd)
        z = getattr(self, 'noise_z', None)
        if z is not None:                      # sensor-noise shim envs/noise/citation.py:72-82 with GIVEN standard-normal draws
            zz = z[self._call]
            out = np.array(out, dtype=np.float64)
            out[:3] += 3.0 * 10**(-5) + 6.3 * 10**(-4) * zz[0:3]
            out[4] += 4.0 * 10**(-10) * zz[3]
            out[5] += 1.8 * 10**(-3) + 2.7 * 10**(-4) * zz[4]
            out[6:8] += 4.0 * 10**(-3) + 3.2 * 10**(-5) * zz[5:7]
        self._call = getattr(self, '_call', 0) + 1
        return out

    # phlabenv.py:401-428
    def reset(self, levels=None, starts=None):
        self.t = 0.
        self._call = 0
        self.X = self.plant.initial_state()
        self.last_u = np.zeros(3)
        self.x = self._native_step(self.last_u)
        self.V0 = self.x[3]
        self.theta_trim = np.rad2deg(self.x[7])
        self.levels, self.starts = levels, starts
        self.error = np.zeros(3)           # fresh env object (see module docstring)
        self.obs = np.hstack((self.error.flatten(), self.x[self.obs_idx]))
        return self.obs

    def ref_deg(self):
        return np.array([refsig.ref_value_deg(self.levels[0], self.starts[0], self.t, self.theta_trim, self.smooth_w, self.t_max),
                         refsig.ref_value_deg(self.levels[1], self.starts[1], self.t, 0.0, self.smooth_w), 0.0])

    # phlabenv.py:430-482
    def step(self, action):
        u = self.scale_action(action)
        self.x = self._native_step(u)
        self.ref_values = np.deg2rad(self.ref_deg())
        self.error[:] = self.ref_values - np.asarray([self.x[7], self.x[6], self.x[5]])
        reward_vec = np.abs(np.clip(self.error_scaler * self.error, -self.max_bound, self.max_bound))
        reward = -reward_vec.sum() / 3
        self.obs = np.hstack((self.error.flatten(), self.x[self.obs_idx]))
        self.last_u = u
        done, penalty = False, 0.
        if self.t >= self.t_max or np.abs(self.x[7]) > self.max_theta or np.abs(self.x[6]) > self.max_phi or self.x[9] < 50:
            penalty = -1 / self.dt * (self.t_max - self.t) * 2
            done = True
        reward += penalty
        self.t += self.dt
        return self.obs, reward, done, {'ref': self.ref_values, 'x': self.x, 't': self.t}


def run_episode(env, actor, levels, starts, record=False):
    """base/core/agent.py:63-138 without noise / replay: returns dict(fitness, steps, t, rewards[, states, actions])."""
    obs = env.reset(levels, starts)
    done = False
    rewards, states, actions = [], [], []
    while not done:
        action = actor.select_action(obs)
        obs, reward, done, info = env.step(action.flatten())
        rewards.append(reward)
        if record:
            states.append(env.x.copy())
            actions.append(env.last_u.copy())
    out = {'fitness': float(np.sum(rewards)), 'steps': len(rewards), 't': info['t'], 'rewards': np.asarray(rewards)}
    if record:
        out['states'] = np.asarray(states)
        out['actions'] = np.asarray(actions)
    return out
