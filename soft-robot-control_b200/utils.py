"""State-layout helpers and containers of sofacontrol/utils.py that the hot path uses (utils.py:8-16, 129-159).
Pure data plumbing -- no numerics live here."""
import os
import pickle

import numpy as np


class QuadraticCost:
    """utils.py:8-16."""

    def __init__(self, Q=None, R=None, Qf=None):
        self.Qf = Qf
        self.Q = Q
        self.R = R


def qv2x(q, v):
    """Reduced/full state is x = [v; q]  (utils.py:129-130); extends to stacked points."""
    return np.concatenate((v, q), axis=-1)


def x2qv(x):
    """Returns (q, v) from x = [v; q]  (utils.py:133-142)."""
    half = x.shape[-1] // 2
    if x.ndim == 1:
        return x[half:], x[:half]
    if x.ndim == 2:
        return x[:, half:], x[:, :half]
    raise IndexError('Unable to process x.ndim > 2')


def vq2qv(x):
    """utils.py:144-146."""
    q, v = x2qv(x)
    return np.hstack((q, v))


def save_data(filename, data):
    """utils.py:148-153."""
    folder = os.path.split(filename)[0]
    if folder and not os.path.isdir(folder):
        os.mkdir(folder)
    with open(filename, 'wb') as file:
        pickle.dump(data, file, protocol=pickle.HIGHEST_PROTOCOL)


def load_data(filename):
    """utils.py:156-159."""
    with open(filename, 'rb') as file:
        return pickle.load(file)
