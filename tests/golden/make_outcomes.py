#!/usr/bin/env python
"""Golden episode outcomes of the UNMODIFIED reference on its own test seeds (rate parity, north_star item 3).

Runs `rl/test.py`-style episodes (SARL baseline weights, configs/test_configs/test_env_configs/env_adults_5.config,
phase "test", seeds 1000..1000+N-1) in a process pool like rl/test_parallel.py:165-173 and writes
tests/golden/outcomes_cfg1.json: per seed the final Info class, global_time and step count.
Build-container only (needs /root/reference); `rvo2` = oracle/shims/rvo2.py.
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("EBC_REFERENCE", "/root/reference")

if os.environ.get("EBC_GOLDEN_CHILD") != "1":
    env = dict(os.environ)
    env["EBC_GOLDEN_CHILD"] = "1"
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(REPO, "oracle", "shims"), REF, os.path.join(REF, "tests")])
    sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], cwd=REF, env=env))

import multiprocessing as mp  # noqa: E402


def run(seed):
    import torch
    torch.set_num_threads(1)
    from simulator.utils.test_utils import configure_env_policy_robot
    env, policy, robot = configure_env_policy_robot(
        "configs/test_configs/test_env_configs/env_adults_5.config",
        "configs/test_configs/test_policy_configs/policy.config", "model_weights/sarl_model_baseline.pth")
    env.get_local_map_angular = lambda *a, **k: None
    ob, lm = env.reset("test", scene_number=seed)
    done, steps = False, 0
    while not done:
        ob, _, reward, done, info = env.step(robot.act(ob, local_map=lm, env=env))
        steps += 1
    return {"seed": seed, "info": type(info).__name__, "time": env.global_time, "steps": steps}


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    with mp.Pool(os.cpu_count()) as pool:
        rows = pool.map(run, range(1000, 1000 + n))
    json.dump(rows, open(os.path.join(HERE, "outcomes_cfg1.json"), "w"), indent=0)
    ok = sum(r["info"] == "ReachGoal" for r in rows)
    print("wrote %d outcomes: success %d, collisions %d, timeouts %d" % (
        len(rows), ok, sum(r["info"].startswith("Collision") for r in rows), sum(r["info"] == "Timeout" for r in rows)))
