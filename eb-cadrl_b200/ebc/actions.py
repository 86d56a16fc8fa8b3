"""Discrete action table, built on the host exactly like rl/policy/cadrl.py:91-116
(float64 numpy; rotation-major, speed-minor; index 0 = stop)."""
import itertools

import numpy as np


def build_action_space(v_pref, kinematics="holonomic", speed_samples=5, rotation_samples=16):
    holonomic = kinematics == "holonomic"
    speeds = [(np.exp((i + 1) / speed_samples) - 1) / (np.e - 1) * v_pref for i in range(speed_samples)]
    if holonomic:
        rotations = np.linspace(0, 2 * np.pi, rotation_samples, endpoint=False)
    else:
        rotations = np.linspace(-np.pi / 4, np.pi / 4, rotation_samples)
    table = [(0.0, 0.0)]
    for rotation, speed in itertools.product(rotations, speeds):
        if holonomic:
            table.append((speed * np.cos(rotation), speed * np.sin(rotation)))   # ActionXY(vx, vy)
        else:
            table.append((speed, rotation))                                      # ActionRot(v, r)
    return np.asarray(table, dtype=np.float64)
