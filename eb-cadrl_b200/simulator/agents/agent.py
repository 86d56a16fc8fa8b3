"""Agent base class — the plugin type of simulator/agents/agent.py:11-234 (same attributes, same methods).

Inside the batched env the kinematics run on the device (K2 / K3); after every env.step the env writes the
device state back into these objects, so code that reads `agent.px` etc. keeps working.  The host-side
kinematics below serve single objects (scene generation, tests, user code) and are organised around two
frame helpers instead of one branch per method; every float64 operation happens in the reference's order
(pinned bit for bit by tests/golden/agent_kinematics.npz)."""
import logging

import numpy as np

from simulator.policy.policy_factory import policy_factory
from simulator.utils.action import ActionRot, ActionXY, ActionXYRot
from simulator.utils.state import FullState, ObservableState
from simulator.utils.utils import AgentType

_POSE = ("px", "py", "gx", "gy", "vx", "vy", "theta")
_SAMPLED = ("v_pref_min", "v_pref_max", "radius_min", "radius_max")
_OPTIONAL = ("radius", "v_pref", "agent_type")


def _body_frame(action):
    """(forward, lateral) speed of a non-holonomic action; lateral is None for ActionRot."""
    if isinstance(action, ActionRot):
        return action.v, None
    if isinstance(action, ActionXYRot):
        return action.vx, action.vy
    raise Exception("Wrong action type")


def _to_world(heading, forward, lateral):
    """Body-frame velocity -> world frame (agent.py:190-200)."""
    c, s = np.cos(heading), np.sin(heading)
    if lateral is None:
        return forward * c, forward * s
    return forward * c - lateral * s, forward * s + lateral * c


def _displace(x, y, heading, forward, lateral, dt):
    """Position after dt along a body-frame velocity (agent.py:172-186: cos/sin first, then speed, then dt)."""
    c, s = np.cos(heading), np.sin(heading)
    x, y = x + c * forward * dt, y + s * forward * dt
    if lateral is not None:
        x, y = x - s * lateral * dt, y + c * lateral * dt
    return x, y


class Agent(object):
    def __init__(self, config, section):
        self.visible = config.getboolean(section, "visible")
        for key in ("v_pref", "radius") + _SAMPLED:
            setattr(self, key, config.getfloat(section, key, fallback=None))
        self.sensor = config.get(section, "sensor")
        self.policy = policy_factory[config.get(section, "policy")]()
        self.kinematics = None if self.policy is None else self.policy.kinematics
        for key in _POSE + ("time_step", "agent_type"):
            setattr(self, key, None)

    # ---- configuration --------------------------------------------------------------------------------
    def print_info(self):
        """The log line rl/train.py:201 and rl/test.py:120 ask for (agent.py:37-42)."""
        logging.info("Agent is %s and has %s kinematic constraint", "visible" if self.visible else "invisible",
                     self.kinematics)

    def set_policy(self, policy):
        self.policy, self.kinematics = policy, policy.kinematics

    def sample_random_attributes(self):
        """v_pref, then radius, from numpy's global stream (agent.py:49-56: the draw order is part of the
        scene generator's RNG replay)."""
        self.v_pref = np.random.uniform(self.v_pref_min, self.v_pref_max)
        self.radius = np.random.uniform(self.radius_min, self.radius_max)
        assert 0 < self.v_pref < 20 and 0 < self.radius < 20

    def set(self, px, py, gx, gy, vx, vy, theta, radius=None, v_pref=None, agent_type=None):
        for key, value in zip(_POSE, (px, py, gx, gy, vx, vy, theta)):
            setattr(self, key, value)
        for key, value in zip(_OPTIONAL, (radius, v_pref, agent_type)):
            if value is not None:
                setattr(self, key, value)

    # ---- state views ----------------------------------------------------------------------------------
    def get_position(self):
        return self.px, self.py

    def set_position(self, position):
        self.px, self.py = position[0], position[1]

    def get_goal_position(self):
        return self.gx, self.gy

    def get_velocity(self):
        return self.vx, self.vy

    def set_velocity(self, velocity):
        self.vx, self.vy = velocity[0], velocity[1]

    def get_observable_state(self):
        return ObservableState(*self.get_position(), *self.get_velocity(), self.radius, self.agent_type)

    def get_full_state(self):
        return FullState(*self.get_position(), *self.get_velocity(), self.radius, *self.get_goal_position(),
                         self.v_pref, self.theta, self.agent_type)

    def get_state_dict(self):
        """Wire format of the scene files (scene_generator.py:865-886)."""
        return dict(pos=self.get_position(), vel=self.get_velocity(), radius=self.radius,
                    goal=self.get_goal_position(), v_pref=self.v_pref, theta=self.theta, agent_type=self.agent_type)

    def set_from_state_dict(self, state):
        self.set_position(state["pos"])
        self.set_velocity(state["vel"])
        self.gx, self.gy = state["goal"][0], state["goal"][1]
        for key in ("radius", "v_pref", "theta"):
            setattr(self, key, state[key])
        if state.get("agent_type") is not None:
            self.agent_type = AgentType(state["agent_type"])

    def reached_destination(self):
        gap = np.array(self.get_position()) - np.array(self.get_goal_position())
        return np.linalg.norm(gap) < self.radius

    # ---- kinematics -----------------------------------------------------------------------------------
    def act(self, ob):
        raise NotImplementedError

    def check_validity(self, action):
        wanted = ActionXY if self.kinematics == "holonomic" else (ActionRot, ActionXYRot)
        assert isinstance(action, wanted)

    def compute_position(self, action, delta_t):
        self.check_validity(action)
        if self.kinematics == "holonomic":
            return self.px + action.vx * delta_t, self.py + action.vy * delta_t
        return _displace(self.px, self.py, self.theta + action.r, *_body_frame(action), delta_t)

    def compute_velocity(self, action):
        self.check_validity(action)
        return _to_world(self.theta + action.r, *_body_frame(action))

    def get_next_observable_state(self, action):
        """The state another agent would observe after `action`, without committing it (agent.py:80-93); the
        heading is not wrapped here."""
        where = self.compute_position(action, self.time_step)
        if self.kinematics == "holonomic":
            velocity = (action.vx, action.vy)
        else:
            velocity = _to_world(self.theta + action.r, action.v, None)
        return ObservableState(*where, *velocity, self.radius, self.agent_type)

    def step(self, action):
        """Commit `action`: position from the unwrapped heading, then the heading wrapped into [0, 2 pi) and the
        velocity taken along the wrapped one (agent.py:202-228)."""
        self.px, self.py = self.compute_position(action, self.time_step)
        if self.kinematics == "holonomic":
            self.vx, self.vy = action.vx, action.vy
            return
        self.theta = (self.theta + action.r) % (2 * np.pi)
        self.vx, self.vy = _to_world(self.theta, *_body_frame(action))
