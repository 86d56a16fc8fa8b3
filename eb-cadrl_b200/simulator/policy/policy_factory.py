"""String -> class registry (simulator/policy/policy_factory.py:10-14).  `orca_obstacles` is dead code
in the reference and is not provided."""
from simulator.policy.linear import Linear
from simulator.policy.orca import ORCA


def none_policy():
    return None


policy_factory = dict()
policy_factory["linear"] = Linear
policy_factory["orca"] = ORCA
policy_factory["none"] = none_policy
