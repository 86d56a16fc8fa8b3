"""ORCA policy (simulator/policy/orca.py:10-157).  The reference keeps one rvo2 simulator per agent
and calls doStep; here `predict` hands the joint state to the device (ebc_robot_orca: the ego agent as
RVO2 agent 0, every observed agent as a neighbour, same parameters) through a one-episode BatchedSim.
Inside env.step the humans' ORCA actions are computed for the whole batch by ebc_orca instead."""
import numpy as np

from ebc.config import SimConfig
from simulator.policy.policy import Policy
from simulator.utils.action import ActionXY

MAX_SPEED = 200


class ORCA(Policy):
    def __init__(self):
        super().__init__()
        self.name = "ORCA"
        self.trainable = False
        self.multiagent_training = None
        self.kinematics = "holonomic"
        self.safety_space = 0
        self.neighbor_dist = 10
        self.max_neighbors = 10
        self.time_horizon = 5
        self.time_horizon_obst = 5
        self.radius = 0.3
        self.max_speed = MAX_SPEED
        self.sim = None

    def configure(self, config):
        return

    def set_phase(self, phase):
        return

    def predict(self, state, env=None):
        from ebc.engine import BatchedSim
        s = state.self_state
        others = state.agent_states
        n = max(len(others), 1)
        if self.sim is None or self.sim.Hmax != n:
            cfg = SimConfig()
            cfg.time_step = self.time_step
            cfg.orca_neighbor_dist, cfg.orca_max_neighbors = self.neighbor_dist, self.max_neighbors
            cfg.orca_time_horizon = self.time_horizon
            self.sim = BatchedSim(cfg, 1, n, 0, 0, 1)
        pv = np.zeros((1, n, 4), np.float32)
        gr = np.zeros((1, n, 4), np.float32)
        for i, o in enumerate(others):
            pv[0, i] = (o.px, o.py, o.vx, o.vy)
            gr[0, i, 3] = o.radius
        self.sim.load_episodes(0, pv, gr, np.zeros((1, n), np.uint8), np.array([len(others)], np.int32), None,
                               np.zeros(1, np.int32), None, np.zeros(1, np.int32),
                               np.array([[s.px, s.py, s.vx, s.vy]], np.float32),
                               np.array([[s.gx, s.gy, s.v_pref, s.radius]], np.float32),
                               np.array([s.theta or 0.0], np.float32), np.zeros(1))
        a = self.sim.robot_orca(float(self.safety_space)).cpu().numpy()[0]
        self.last_state = state
        return ActionXY(float(a[0]), float(a[1]))
