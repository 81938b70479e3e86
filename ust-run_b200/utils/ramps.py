"""Consistency ramp (reference: utils/ramps.py:19-26, used at train.py:82-84,819-820).

Host-side scalar math; kept in double precision like the reference's numpy version."""
import math

import numpy as np


def sigmoid_rampup(current, rampup_length):
    """exp(-5 (1 - t)^2) with t = clip(current, 0, L) / L; 1.0 when L == 0."""
    if rampup_length == 0:
        return 1.0
    phase = 1.0 - float(np.clip(current, 0.0, rampup_length)) / rampup_length
    return float(np.exp(-5.0 * phase * phase))      # numpy exp: bit-identical to the reference


def linear_rampup(current, rampup_length):
    assert current >= 0 and rampup_length >= 0
    return 1.0 if current >= rampup_length else current / rampup_length


def cosine_rampdown(current, rampdown_length):
    assert 0 <= current <= rampdown_length
    return float(0.5 * (math.cos(math.pi * current / rampdown_length) + 1))
