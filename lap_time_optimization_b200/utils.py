"""Small host helpers the hot-path callers use (reference src/utils.py:5-22)."""
import numpy as np


def idx_modulo(a, b, n):
    """Indices a..b-1 on a ring of n elements (utils.py:5-14)."""
    i, j = a % n, b % n
    if i < j:
        return np.arange(i, j, dtype=int)
    return np.append(np.arange(i, n, dtype=int), np.arange(0, j, dtype=int))


def is_closed(left, right):
    """A track is a closed loop when both boundaries end where they start (utils.py:17-22)."""
    return bool(all(left[:, 0] == left[:, -1]) and all(right[:, 0] == right[:, -1]))
