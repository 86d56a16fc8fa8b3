"""configure_env_policy_robot — the reference's canonical way to stand up env + policy + robot
(simulator/utils/test_utils.py:8-36), same signature."""
import configparser

import numpy as np
import torch

import simulator
from rl.policy.policy_factory import policy_factory
from simulator.agents.robot import Robot


def configure_env_policy_robot(env_config_path, policy_config_path, model_path=None, phase="test", device="cpu",
                               policy="sarl", env_name="EntityBasedCollisionAvoidance-v0"):
    env_config = configparser.RawConfigParser()
    env_config.read(env_config_path)
    env = simulator.make(env_name)
    env.configure(env_config)
    robot = Robot(env_config, "robot")
    env.set_robot(robot)
    policy = policy_factory[policy]()
    policy_config = configparser.RawConfigParser()
    policy_config.read(policy_config_path)
    policy.configure(policy_config)
    if model_path is not None:
        if str(model_path).endswith(".npz"):
            z = np.load(model_path)
            state = {k: torch.as_tensor(z[k]) for k in z.files}
        else:
            state = torch.load(model_path, map_location="cpu")
        policy.get_model().load_state_dict(state)
    robot.set_policy(policy)
    policy.set_phase(phase)
    policy.set_device(device)
    return env, policy, robot
