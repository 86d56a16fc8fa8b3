"""Policy plugin base class (simulator/policy/policy.py:5-54): same attributes and methods."""
import numpy as np


class Policy(object):
    def __init__(self):
        self.trainable = False
        self.phase = None
        self.model = None
        self.device = None
        self.last_state = None
        self.time_step = None

    def configure(self, config):
        return

    def set_phase(self, phase):
        self.phase = phase

    def set_device(self, device):
        self.device = device

    def get_model(self):
        return self.model

    def predict(self, state, env=None):
        raise NotImplementedError

    @staticmethod
    def reach_destination(state):
        s = state.self_state
        return bool(np.linalg.norm((s.py - s.gy, s.px - s.gx)) < s.radius)
