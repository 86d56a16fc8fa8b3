"""Reward configuration (simulator/utils/reward.py:18-75).  The reward / done / event classification
itself (`Reward.compute`, :80-181) runs on the device inside ebc_lookahead / ebc_step; this class only
carries the parsed [reward] block so that code reading `env.reward.<attr>` keeps working."""
from ebc.config import SimConfig


class Reward(object):
    def __init__(self, config):
        c = SimConfig.from_ini(config)
        self.new_reward = c.new_reward
        self.time_max = c.time_max if config.has_option("reward", "time_max") else None
        self.max_goal_distance = c.max_goal_distance
        self.time_good = c.time_good
        self.success_reward = c.success_reward
        for t in ("adult", "bicycle", "obstacle", "child"):
            setattr(self, "collision_penalty_" + t, getattr(c, "collision_penalty_" + t))
        self.discomfort_dist = c.discomfort_dist
        self.discomfort_penalty_factor = config.getfloat("reward", "discomfort_penalty_factor")
        for t in ("adult", "bicycle", "child"):
            setattr(self, "discomfort_dist_" + t, getattr(c, "discomfort_dist_" + t))
            setattr(self, "discomfort_penalty_factor_" + t, getattr(c, "discomfort_penalty_factor_" + t))
        self.rotation_penalty_factor = c.rotation_penalty_factor
        self.time_step = c.time_step
        self.time_limit = int(c.time_limit)
        self.robot = None

    def set_robot(self, robot):
        self.robot = robot
