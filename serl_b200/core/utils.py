"""Mirror of base/core/utils.py: the Episode record returned by Agent.evaluate, the action-smoothness metric, config loading."""
import os
from dataclasses import dataclass
from pathlib import Path
from typing import List

import numpy as np


@dataclass
class Episode:
    """base/core/utils.py:12-36"""
    fitness: np.float64
    smoothness: np.float64
    length: np.float64
    state_history: List
    ref_signals: List
    actions: List
    reward_lst: List

    def get_history(self) -> np.ndarray:
        """[refs, actions, states, reward] per step, shape (ep_len, 19) (utils.py:24-36)."""
        tt = np.linspace(0, self.length, len(self.state_history))
        ref_values = np.array([[ref(t_i) for t_i in tt] for ref in self.ref_signals]).transpose()
        reward_lst = np.asarray(self.reward_lst).reshape((len(self.state_history), 1))
        return np.concatenate((ref_values, self.actions, self.state_history, reward_lst), axis=1)


def calc_smoothness(y: np.ndarray, dt: float = 0.01, **kwargs) -> float:
    """base/core/utils.py:82-120 (host-side; the device version is SURVEY.md 8(f) N1)."""
    N, A = y.shape[0], y.shape[1]
    T = N * dt
    freq = np.linspace(dt, 1 / (2 * dt), N // 2 - 1)
    Syy = np.zeros((N // 2 - 1, A))
    for i in range(A):
        Y = np.fft.fft(y[:, i], N)
        Syy_disc = Y[1:N // 2] * np.conjugate(Y[1:N // 2])
        Syy[:, i] = np.abs(Syy_disc) * dt
    signal_roughness = np.einsum('ij,i -> j', Syy, freq) * 2 / N
    roughness = np.sqrt(np.sum(signal_roughness, axis=-1)) * 100 * (80 / T)
    return -roughness


def load_config(model_path: str, verbose: bool = False) -> dict:
    import yaml
    model_path = model_path / Path('files/')
    conf_raw = yaml.safe_load(Path(os.path.join(model_path, 'config.yaml')).read_text(encoding='utf-8'))
    return {k: (v['value'] if isinstance(v, dict) else v) for k, v in conf_raw.items()}
