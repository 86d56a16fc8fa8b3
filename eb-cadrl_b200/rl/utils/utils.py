"""configure_policy / configure_environment_and_robot (rl/utils/utils.py:9-32), batched."""
import configparser

import torch

from ebc.batched_env import BatchedEnv
from rl.policy.policy_factory import policy_factory


def configure_policy(args):
    policy = policy_factory[args.policy]()
    if not policy.trainable:
        raise ValueError("Policy has to be trainable")
    if args.policy_config is None:
        raise ValueError("Policy config has to be specified for a trainable network")
    policy_config = configparser.RawConfigParser()
    policy_config.read(args.policy_config)
    policy.configure(policy_config)
    return policy


def configure_environment_and_robot(env_config_path, policy, n_episodes, device):
    env_config = configparser.RawConfigParser()
    env_config.read(env_config_path)
    return BatchedEnv(env_config, policy, n_episodes, device)
