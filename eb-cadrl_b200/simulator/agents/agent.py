"""Agent base class (simulator/agents/agent.py:11-234): physical attributes, state accessors and the
kinematics used on the HOST side of the API (single objects).  Inside the batched env the same
kinematics run on the device (K2 / K3); after every env.step the env copies the device state back
into these objects so that code reading `agent.px` etc. keeps working."""
import numpy as np
from numpy.linalg import norm

from simulator.policy.policy_factory import policy_factory
from simulator.utils.action import ActionRot, ActionXY, ActionXYRot
from simulator.utils.state import FullState, ObservableState
from simulator.utils.utils import AgentType


class Agent(object):
    def __init__(self, config, section):
        self.visible = config.getboolean(section, "visible")
        self.v_pref = config.getfloat(section, "v_pref", fallback=None)
        self.radius = config.getfloat(section, "radius", fallback=None)
        self.policy = policy_factory[config.get(section, "policy")]()
        self.sensor = config.get(section, "sensor")
        self.kinematics = self.policy.kinematics if self.policy is not None else None
        self.px = self.py = self.gx = self.gy = self.vx = self.vy = self.theta = None
        self.time_step = None
        self.agent_type = None
        self.v_pref_min = config.getfloat(section, "v_pref_min", fallback=None)
        self.v_pref_max = config.getfloat(section, "v_pref_max", fallback=None)
        self.radius_min = config.getfloat(section, "radius_min", fallback=None)
        self.radius_max = config.getfloat(section, "radius_max", fallback=None)

    def set_policy(self, policy):
        self.policy = policy
        self.kinematics = policy.kinematics

    def sample_random_attributes(self):
        self.v_pref = np.random.uniform(self.v_pref_min, self.v_pref_max)
        self.radius = np.random.uniform(self.radius_min, self.radius_max)
        assert 0 < self.v_pref < 20
        assert 0 < self.radius < 20

    def set(self, px, py, gx, gy, vx, vy, theta, radius=None, v_pref=None, agent_type=None):
        self.px, self.py, self.gx, self.gy, self.vx, self.vy, self.theta = px, py, gx, gy, vx, vy, theta
        if radius is not None:
            self.radius = radius
        if v_pref is not None:
            self.v_pref = v_pref
        if agent_type is not None:
            self.agent_type = agent_type

    def get_observable_state(self):
        return ObservableState(self.px, self.py, self.vx, self.vy, self.radius, self.agent_type)

    def get_next_observable_state(self, action):
        self.check_validity(action)
        next_px, next_py = self.compute_position(action, self.time_step)
        if self.kinematics == "holonomic":
            next_vx, next_vy = action.vx, action.vy
        else:
            next_theta = self.theta + action.r
            next_vx, next_vy = action.v * np.cos(next_theta), action.v * np.sin(next_theta)
        return ObservableState(next_px, next_py, next_vx, next_vy, self.radius, self.agent_type)

    def get_full_state(self):
        return FullState(self.px, self.py, self.vx, self.vy, self.radius, self.gx, self.gy, self.v_pref,
                         self.theta, self.agent_type)

    def get_state_dict(self):
        return {"pos": (self.px, self.py), "vel": (self.vx, self.vy), "radius": self.radius,
                "goal": (self.gx, self.gy), "v_pref": self.v_pref, "theta": self.theta,
                "agent_type": self.agent_type}

    def set_from_state_dict(self, state):
        self.px, self.py = state["pos"]
        self.vx, self.vy = state["vel"]
        self.radius = state["radius"]
        self.gx, self.gy = state["goal"]
        self.v_pref = state["v_pref"]
        self.theta = state["theta"]
        if state.get("agent_type") is not None:
            self.agent_type = AgentType(state["agent_type"])

    def get_position(self):
        return self.px, self.py

    def set_position(self, position):
        self.px, self.py = position

    def get_goal_position(self):
        return self.gx, self.gy

    def get_velocity(self):
        return self.vx, self.vy

    def set_velocity(self, velocity):
        self.vx, self.vy = velocity

    def act(self, ob):
        raise NotImplementedError

    def check_validity(self, action):
        if self.kinematics == "holonomic":
            assert isinstance(action, ActionXY)
        else:
            assert isinstance(action, (ActionRot, ActionXYRot))

    def compute_position(self, action, delta_t):
        self.check_validity(action)
        if self.kinematics == "holonomic":
            return self.px + action.vx * delta_t, self.py + action.vy * delta_t
        theta = self.theta + action.r
        if isinstance(action, ActionRot):
            return (self.px + np.cos(theta) * action.v * delta_t, self.py + np.sin(theta) * action.v * delta_t)
        if isinstance(action, ActionXYRot):
            return (self.px + np.cos(theta) * action.vx * delta_t - np.sin(theta) * action.vy * delta_t,
                    self.py + np.sin(theta) * action.vx * delta_t + np.cos(theta) * action.vy * delta_t)
        raise Exception("Wrong action type")

    def compute_velocity(self, action):
        self.check_validity(action)
        theta = self.theta + action.r
        if isinstance(action, ActionRot):
            return action.v * np.cos(theta), action.v * np.sin(theta)
        return (action.vx * np.cos(theta) - action.vy * np.sin(theta),
                action.vx * np.sin(theta) + action.vy * np.cos(theta))

    def step(self, action):
        self.check_validity(action)
        self.px, self.py = self.compute_position(action, self.time_step)
        if self.kinematics == "holonomic":
            self.vx, self.vy = action.vx, action.vy
        else:
            self.theta = (self.theta + action.r) % (2 * np.pi)
            if isinstance(action, ActionRot):
                self.vx, self.vy = action.v * np.cos(self.theta), action.v * np.sin(self.theta)
            elif isinstance(action, ActionXYRot):
                self.vx = action.vx * np.cos(self.theta) - action.vy * np.sin(self.theta)
                self.vy = action.vx * np.sin(self.theta) + action.vy * np.cos(self.theta)
            else:
                raise Exception("Wrong action type")

    def reached_destination(self):
        return norm(np.array(self.get_position()) - np.array(self.get_goal_position())) < self.radius
