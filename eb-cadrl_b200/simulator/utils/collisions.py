"""Robot / agent swept-disc test (simulator/utils/collisions.py:4-57) — evaluated by the device's
committed-step kernel (`evaluate_action`, the same code the batched env runs) on a one-agent episode, so
that the reference's own known-answer tests (tests/test_collisions.py) exercise the C ABI."""
import numpy as np
import torch

from ebc import abi
from ebc.config import SimConfig
from ebc.engine import BatchedSim
from simulator.utils.action import ActionXY

_sims = {}


def _sim(time_step, kinematics):
    key = (float(time_step), kinematics)
    if key not in _sims:
        cfg = SimConfig()
        cfg.time_step = float(time_step)
        cfg.robot_kinematics = kinematics
        cfg.collision_penalty_adult = cfg.collision_penalty_bicycle = cfg.collision_penalty_child = -1.0
        cfg.collision_penalty_obstacle = -1.0
        cfg.discomfort_dist_adult = cfg.discomfort_dist_bicycle = cfg.discomfort_dist_child = 0.0
        _sims[key] = BatchedSim(cfg, 1, 1, 0, 0, 1)
    return _sims[key]


def point_to_segment_dist(x1, y1, x2, y2, x3, y3):
    """Closest distance from (x3, y3) to the segment (x1, y1)-(x2, y2): a robot of radius 0 at (x3, y3)
    against an agent of radius 0 sweeping the segment in one unit time step."""
    robot = type("R", (), dict(px=x3, py=y3, theta=0.0, radius=0.0, kinematics="holonomic"))()
    agent = type("A", (), dict(px=x1, py=y1, vx=x2 - x1, vy=y2 - y1, radius=0.0))()
    dmin, collision = compute_collision_agent_with_robot(agent, robot, ActionXY(0.0, 0.0), float("inf"), 1.0)
    return 0.0 if collision else dmin


def compute_collision_agent_with_robot(agent, robot, action, dmin, time_step):
    kin = "holonomic" if robot.kinematics == "holonomic" else "unicycle"
    sim = _sim(time_step, kin)
    sim.load_episodes(0, np.array([[[agent.px, agent.py, agent.vx, agent.vy]]], np.float32),
                      np.array([[[0.0, 0.0, 1.0, agent.radius]]], np.float32), np.zeros((1, 1), np.uint8),
                      np.array([1], np.int32), None, np.zeros(1, np.int32), None, np.zeros(1, np.int32),
                      np.array([[robot.px, robot.py, 0.0, 0.0]], np.float32),
                      np.array([[1e6, 1e6, 1.0, robot.radius]], np.float32),
                      np.array([robot.theta], np.float32), np.zeros(1))
    sim.hum_nv.zero_()
    a = (action.vx, action.vy) if kin == "holonomic" else (action.v, action.r)
    sim.step(action=torch.tensor([a], dtype=torch.float64, device=sim.device))
    collision = int(sim.event[0].item()) == abi.EV_COLLISION_ADULT
    closest = float(sim.dmin[0, 0].item())
    if not collision and closest < dmin:
        dmin = closest
    return dmin, collision
