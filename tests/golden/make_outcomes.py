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


CONFIGS = {
    # tag -> (env config, policy config, weights, first seed)
    "cfg1": ("configs/test_configs/test_env_configs/env_adults_5.config",
             "configs/test_configs/test_policy_configs/policy.config", "model_weights/sarl_model_baseline.pth", 1000),
    # the shipped EB-CADRL experiment: 24 typed humans + 3 walls, entity-typed rows (the network the bench runs)
    "ebcadrl": ("data/eb-cadrl/adults_8_bikes_8_child_8_static_3_35_sec_new_reward_fix_static.config",
                "data/eb-cadrl/policy_x2_agent_type.config", "data/eb-cadrl/rl_model_val.pth", 1000),
}
TAG = "cfg1"


def run(seed):
    import torch
    torch.set_num_threads(1)
    from simulator.utils.test_utils import configure_env_policy_robot
    env_cfg, pol_cfg, weights, _ = CONFIGS[TAG]
    env, policy, robot = configure_env_policy_robot(env_cfg, pol_cfg, weights)
    env.get_local_map_angular = lambda *a, **k: None
    ob, lm = env.reset("test", scene_number=seed)
    done, steps = False, 0
    while not done:
        ob, _, reward, done, info = env.step(robot.act(ob, local_map=lm, env=env))
        steps += 1
    return {"seed": seed, "info": type(info).__name__, "time": env.global_time, "steps": steps}


def _init(tag):
    global TAG
    TAG = tag


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    tag = sys.argv[2] if len(sys.argv) > 2 else "cfg1"
    first = CONFIGS[tag][3]
    with mp.Pool(os.cpu_count(), initializer=_init, initargs=(tag,)) as pool:
        rows = pool.map(run, range(first, first + n), chunksize=1)
    json.dump(rows, open(os.path.join(HERE, "outcomes_%s.json" % tag), "w"), indent=0)
    ok = sum(r["info"] == "ReachGoal" for r in rows)
    print("wrote %d outcomes: success %d, collisions %d, timeouts %d" % (
        len(rows), ok, sum(r["info"].startswith("Collision") for r in rows), sum(r["info"] == "Timeout" for r in rows)))
