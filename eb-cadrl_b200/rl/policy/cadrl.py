"""CADRL base policy (rl/policy/cadrl.py): configuration, the discrete action space, robot `propagate`
and the agent-centric `rotate`.  The hot loops (lookahead + rotate + value net) run on the device; the
host versions kept here are the API surface other reference code calls (`propagate`, `rotate` on a tensor).
CADRL's own single-human `predict` (min over humans) is out of scope."""
import logging

import numpy as np
import torch
import torch.nn as nn
from torch.nn.functional import one_hot

from ebc.actions import build_action_space as _build_table
from simulator.policy.policy import Policy
from simulator.utils.action import ActionRot, ActionXY
from simulator.utils.state import FullState, ObservableState


def mlp(input_dim, mlp_dims, last_relu=False):
    """nn.Sequential of Linear/ReLU with the reference's module indices (0, 2, 4, ...) so that shipped
    state_dicts load unchanged (rl/policy/cadrl.py:13-21)."""
    dims = [input_dim] + list(mlp_dims)
    layers = []
    for i in range(len(dims) - 1):
        layers.append(nn.Linear(dims[i], dims[i + 1]))
        if i != len(dims) - 2 or last_relu:
            layers.append(nn.ReLU())
    return nn.Sequential(*layers)


class CADRL(Policy):
    def __init__(self):
        super().__init__()
        self.name = "CADRL"
        self.trainable = True
        self.multiagent_training = None
        self.kinematics = None
        self.epsilon = None
        self.gamma = None
        self.sampling = None
        self.speed_samples = None
        self.rotation_samples = None
        self.query_env = None
        self.action_space = None
        self.speeds = None
        self.rotations = None
        self.action_values = None
        self.with_om = None
        self.with_agent_type = False
        self.cell_num = self.cell_size = self.om_channel_size = None
        self.self_state_dim = 6
        self.agent_state_dim = 7
        self.agent_type_state_dim = 0
        self.joint_state_dim = self.self_state_dim + self.agent_state_dim
        self.weights_version = 0      # bumped by the trainer whenever the model's parameters change

    def set_common_parameters(self, config):
        self.gamma = config.getfloat("rl", "gamma")
        self.kinematics = config.get("action_space", "kinematics")
        self.sampling = config.get("action_space", "sampling")
        self.speed_samples = config.getint("action_space", "speed_samples")
        self.rotation_samples = config.getint("action_space", "rotation_samples")
        self.query_env = config.getboolean("action_space", "query_env")
        self.cell_num = config.getint("om", "cell_num")
        self.cell_size = config.getfloat("om", "cell_size")
        self.om_channel_size = config.getint("om", "om_channel_size")

    @property
    def n_actions(self):
        return (self.speed_samples or 5) * (self.rotation_samples or 16) + 1

    def set_device(self, device):
        self.device = device
        if self.model is not None:
            self.model.to(device)

    def set_epsilon(self, epsilon):
        self.epsilon = epsilon

    def build_action_space(self, v_pref):
        """81 actions, rotation-major (rl/policy/cadrl.py:91-116)."""
        table = _build_table(v_pref, self.kinematics, self.speed_samples, self.rotation_samples)
        holonomic = self.kinematics == "holonomic"
        self.action_space = [ActionXY(a, b) if holonomic else ActionRot(a, b) for a, b in table]
        self.action_table = table
        return self.action_space

    def propagate(self, state, action):
        """Next state of one agent under `action` (rl/policy/cadrl.py:118-165)."""
        dt = self.time_step
        if isinstance(state, ObservableState):
            return ObservableState(state.px + action.vx * dt, state.py + action.vy * dt, action.vx, action.vy,
                                   state.radius, state.obj_type)
        if isinstance(state, FullState):
            if self.kinematics == "holonomic":
                return FullState(state.px + action.vx * dt, state.py + action.vy * dt, action.vx, action.vy,
                                 state.radius, state.gx, state.gy, state.v_pref, state.theta, state.obj_type)
            th = state.theta + action.r
            vx, vy = action.v * np.cos(th), action.v * np.sin(th)
            return FullState(state.px + vx * dt, state.py + vy * dt, vx, vy, state.radius, state.gx, state.gy,
                             state.v_pref, th, state.obj_type)
        raise ValueError("Type error")

    def rotate(self, state):
        """Agent-centric transform of (batch, 15) joint rows (rl/policy/cadrl.py:236-337), torch fp32.
        Device twin: rotate_row in csrc/ebc_sim.cu (used by ebc_lookahead / ebc_transform)."""
        px, py, vx, vy, r, gx, gy, vp, th = (state[:, i] for i in range(9))
        px1, py1, vx1, vy1, r1 = (state[:, i] for i in range(9, 14))
        rot = torch.atan2(gy - py, gx - px)
        c, s = torch.cos(rot), torch.sin(rot)
        dg = torch.norm(torch.stack([gx - px, gy - py], 1), 2, dim=1)
        theta = (th - rot) if self.kinematics == "unicycle" else torch.zeros_like(vp)
        da = torch.norm(torch.stack([px - px1, py - py1], 1), 2, dim=1)
        cols = [dg, vp, theta, r, vx * c + vy * s, vy * c - vx * s, (px1 - px) * c + (py1 - py) * s,
                (py1 - py) * c - (px1 - px) * s, vx1 * c + vy1 * s, vy1 * c - vx1 * s, r1, da, r + r1]
        out = torch.stack(cols, 1)
        if self.with_agent_type:
            out = torch.cat([out, one_hot(state[:, 14].long(), num_classes=self.agent_type_state_dim).to(out.dtype)], 1)
        return out
