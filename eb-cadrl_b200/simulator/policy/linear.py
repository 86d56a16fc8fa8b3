"""Go-to-goal policy (simulator/policy/linear.py:6-23): full preferred speed straight at the goal,
never stops.  Three float64 operations on the host; as a HUMAN policy (`policy = linear` in an env
INI) it runs inside K1 on the device."""
import numpy as np

from simulator.policy.policy import Policy
from simulator.utils.action import ActionXY


class Linear(Policy):
    def __init__(self):
        super().__init__()
        self.name = "Linear"
        self.trainable = False
        self.kinematics = "holonomic"
        self.multiagent_training = True

    def configure(self, config):
        return

    def predict(self, state, env=None):
        s = state.self_state
        theta = np.arctan2(s.gy - s.py, s.gx - s.px)
        return ActionXY(np.cos(theta) * s.v_pref, np.sin(theta) * s.v_pref)
